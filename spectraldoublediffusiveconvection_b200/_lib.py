"""ctypes binding of libsddc_b200.so (C ABI in include/sddc_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the product path.
Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDDC_B200_LIB", os.path.join(_HERE, "libsddc_b200.so"))  # override: A/B builds

c_double_p = C.POINTER(C.c_double)


class SddcConfig(C.Structure):
    _fields_ = [("N_fm", C.c_int), ("N_r", C.c_int), ("symmetric", C.c_int), ("max_batch", C.c_int),
                ("device", C.c_int), ("dt", C.c_double), ("Pr", C.c_double), ("Tau", C.c_double), ("d", C.c_double),
                ("flags", C.c_int)]


FLAG_DENSE_TRANSFORMS = 1


class SddcOperators(C.Structure):
    _fields_ = [(name, c_double_p) for name in
                ("Dr", "Dsq", "D2r", "D2", "r2", "ir2", "ir4", "a4_ir2", "a4_ir4", "dT0", "gbuoy", "ir", "nu_in",
                 "nu_out", "r")] + [("R_in", C.c_double), ("R_out", C.c_double)] + \
               [(name, c_double_p) for name in ("Linv_A4", "Linv_T", "Linv_S")]


# every symbol include/sddc_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll, _dp = C.c_void_p, C.c_int, C.c_longlong, C.c_void_p  # device/host data pointers pass as void*
SYMBOLS = {
    "sddc_version": (_i, []),
    "sddc_device_count": (_i, []),
    "sddc_plan_create": (_i, [C.POINTER(_vp), C.POINTER(SddcConfig), C.POINTER(SddcOperators)]),
    "sddc_plan_destroy": (None, [_vp]),
    "sddc_last_error": (C.c_char_p, [_vp]),
    "sddc_launch_count": (_ll, [_vp]),
    "sddc_plan_info": (_i, [_vp, _i]),
    "sddc_plan_set_linv": (_i, [_vp, _i, _dp, C.c_double]),
    "sddc_plan_set_a4_aux": (_i, [_vp, _dp, _dp, _dp]),
    "sddc_nlin_fx": (_i, [_vp, _dp, _dp, _i, _vp]),
    "sddc_nlin_dfx": (_i, [_vp, _dp, _dp, _dp, _i, _vp]),
    "sddc_linear_op": (_i, [_vp, _i, _dp, _dp, _i, _vp]),
    "sddc_solve_a4": (_i, [_vp, _dp, _dp, _i, _vp]),
    "sddc_solve_nab2": (_i, [_vp, _i, _dp, _dp, _i, _vp]),
    "sddc_step": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _i, _i, _vp]),
    "sddc_residual": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _vp]),
    "sddc_jvp": (_i, [_vp, _dp, _dp, _dp, _dp, _dp, _i, _vp]),
    "sddc_jvp_set_base": (_i, [_vp, _dp, _i, _vp]),
    "sddc_jvp_apply": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _vp]),
    "sddc_jvp_apply_plus": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _vp]),
    "sddc_dF_dRa": (_i, [_vp, _dp, _dp, _i, _vp]),
    "sddc_diagnostics": (_i, [_vp, _dp, _dp, _i, _vp]),
    "sddc_transform": (_i, [_i, _dp, _dp, _i, _i, _i, _vp]),
    "sddc_plan_set_ckpt_phase": (_i, [_vp, _i]),
    "sddc_interp_radial": (_i, [_dp, _dp, _dp, _ll, _i, _i, _vp]),
    "sddc_interp_thetas": (_i, [_dp, _dp, _i, _i, _i, _i, _vp]),
    "sddc_gs_chunks": (_i, [_i]),
    "sddc_gs_dots": (_i, [_dp, _ll, _i, _i, _dp, _dp, _i, _i, _vp]),
    "sddc_gs_update": (_i, [_dp, _ll, _i, _i, _dp, _dp, _dp, _dp, _i, _i, _vp, _i, _vp]),
    "sddc_gmres_column": (_i, [_dp, _i, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sddc_profile_begin": (_i, [_vp]),
    "sddc_profile_end": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "sddc_time_step": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _i, _i, _i, _dp, _vp]),
    "sddc_step_host": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _i, _i, _dp]),
    "sddc_time_step_host": (_i, [_vp, _dp, _dp, _dp, _dp, _i, _i, _i, _i, _dp, _i, _dp]),
    "sddc_jvp_host": (_i, [_vp, _dp, _dp, _dp, _dp, _dp, _i]),
}

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsddc_b200.so not found at %s -- the CUDA extension is required (no CPU fallback); "
            "build it with __graft_entry__.build()" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
