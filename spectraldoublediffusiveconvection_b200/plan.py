"""EnsemblePlan: the batched, device-resident API over the C ABI.

One plan holds the operators of one (N_fm, N_r, d, dt, Pr, Tau, symmetric) on one GPU and applies the hot path
to B independent members at once.  States are torch.float64 CUDA tensors of shape [B, 3*nr*N_fm] in the
reference's flat layout; Ra / Ra_s are per-member tensors [B].
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .operators import RadialOperators

_PINNED_KEEPALIVE = {}

OP_J_THETA, OP_DT0_THETA, OP_A2_SINE, OP_A2_SINE_R2, OP_KGR, OP_R2 = range(6)
T_IDCT, T_IDST, T_DCT, T_DST = range(4)


def _ptr(a):
    return a.ctypes.data_as(_lib.c_double_p)


class SddcError(RuntimeError):
    pass


class EnsemblePlan:
    def __init__(self, N_fm, N_r, d, dt, Pr, Tau, symmetric=False, max_batch=1, device=None, operators=None,
                 dense_transforms=False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise SddcError("no CUDA device: the B200 path has no CPU fallback")
        if int(N_fm) % 2 != 0:
            # same check and message as the reference (Matrix_Operators.py:758-759)
            raise ValueError("The number of Fourier modes is not even %d" % N_fm)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.ops = operators if operators is not None else RadialOperators(N_fm, N_r, d, dt, Pr, Tau)
        self.N_fm, self.N_r, self.nr = int(N_fm), int(N_r), int(N_r) - 1
        self.N = self.nr * self.N_fm
        self.symmetric = bool(symmetric)
        self.max_batch = int(max_batch)
        self.dt, self.Pr, self.Tau, self.d = float(dt), float(Pr), float(Tau), float(d)
        cfg = _lib.SddcConfig(self.N_fm, self.N_r, int(self.symmetric), self.max_batch, self.device.index,
                              self.dt, self.Pr, self.Tau, self.d,
                              _lib.FLAG_DENSE_TRANSFORMS if dense_transforms else 0)
        o = self.ops
        cops = _lib.SddcOperators(
            Dr=_ptr(o.Dr), Dsq=_ptr(o.Dsq), D2r=_ptr(o.D2r), D2=_ptr(o.D2), r2=_ptr(o.r2), ir2=_ptr(o.ir2),
            ir4=_ptr(o.ir4), a4_ir2=_ptr(o.a4_ir2), a4_ir4=_ptr(o.a4_ir4), dT0=_ptr(o.dT0), gbuoy=_ptr(o.gbuoy),
            ir=_ptr(o.ir), nu_in=_ptr(o.nu_in), nu_out=_ptr(o.nu_out), r=_ptr(o.r), R_in=float(o.R[0]),
            R_out=float(o.R[-1]), Linv_A4=_ptr(o.L_inv_A4), Linv_T=_ptr(o.L_inv_T), Linv_S=_ptr(o.L_inv_S))
        handle = C.c_void_p()
        rc = self.lib.sddc_plan_create(C.byref(handle), C.byref(cfg), C.byref(cops))
        if rc != 0:
            msg = self.lib.sddc_last_error(None).decode()
            if rc == -1:
                raise ValueError(msg)
            raise SddcError("sddc_plan_create failed (%d): %s" % (rc, msg))
        self._h = handle

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.sddc_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.sddc_last_error(self._h).decode()
            if rc == -1:
                raise ValueError(msg)
            raise SddcError("libsddc_b200 error %d: %s" % (rc, msg))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _in(self, t, width):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
            raise TypeError("expected a float64 CUDA tensor")
        if t.device != self.device:
            raise ValueError("tensor on %s, plan on %s" % (t.device, self.device))
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.shape[-1] != width:
            raise ValueError("last dimension %d, expected %d" % (t.shape[-1], width))
        return t.contiguous()

    def _param(self, v, B):
        """Per-member parameter [B] (a scalar or 1-element tensor is broadcast); the C ABI trusts the pointer, so any
        other length is rejected here."""
        if isinstance(v, (torch.Tensor, np.ndarray, list, tuple)):
            v = torch.as_tensor(v).to(device=self.device, dtype=torch.float64).reshape(-1)
            if v.numel() == 1:
                v = v.expand(B)
            elif v.numel() != B:
                raise ValueError("per-member parameter has %d entries, the batch has %d members" % (v.numel(), B))
            return v.contiguous()
        return torch.full((B,), float(v), dtype=torch.float64, device=self.device)

    def _out(self, out, shape):
        """Result tensor: allocated here, or the caller's - which must be exactly what the kernels write into."""
        shape = tuple(int(x) for x in shape)
        if out is None:
            return torch.empty(shape, dtype=torch.float64, device=self.device)
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float64):
            raise TypeError("out must be a float64 CUDA tensor")
        if out.device != self.device:
            raise ValueError("out on %s, plan on %s" % (out.device, self.device))
        if tuple(out.shape) != shape and not (shape[0] == 1 and tuple(out.shape) == shape[1:]):
            raise ValueError("out has shape %s, expected %s" % (tuple(out.shape), shape))
        if not out.is_contiguous():
            raise ValueError("out must be contiguous")
        return out

    def info(self):
        """Kernel-selection facts: quarter-wave split active, synthesis variant, JVP availability, padded n, grid size
        of the FFT formulation of the nonlinear term (0 = dense DMMA transforms), FFT formulation used for JVPs."""
        names = ("quarter_wave", "synth_variant", "jvp_two_state", "n8", "fft_M", "fft_jvp", "ke_fft_M", "direct_rows",
                 "solve_gather")
        return {nm: int(self.lib.sddc_plan_info(self._h, i)) for i, nm in enumerate(names)}

    @property
    def launch_count(self):
        return int(self.lib.sddc_launch_count(self._h))

    STAGES = ("scan", "prep", "synth", "analysis", "solve", "ke_prep", "ke_synth", "diag")

    def profile_begin(self):
        self._check(self.lib.sddc_profile_begin(self._h))

    def profile_end(self):
        """{stage: (total_ms, launches)} measured with CUDA events around each kernel since profile_begin()."""
        ms = (C.c_double * 8)()
        cnt = (C.c_int * 8)()
        self._check(self.lib.sddc_profile_end(self._h, ms, cnt))
        return {s: (ms[i], cnt[i]) for i, s in enumerate(self.STAGES)}

    def set_linv(self, which, stack, dt_eff):
        """Upload another pre-inverted operator stack [N_fm, nr, nr] (0: A4, 1: NAB2-T, 2: NAB2-S) and the
        effective dt its back-substitution uses."""
        stack = np.ascontiguousarray(stack, dtype=np.float64)
        if stack.shape != (self.N_fm, self.nr, self.nr):
            raise ValueError("operator stack has shape %s, expected %s" % (stack.shape, (self.N_fm, self.nr, self.nr)))
        self._check(self.lib.sddc_plan_set_linv(self._h, int(which), stack.ctypes.data, float(dt_eff)))

    def set_a4_aux(self, D2, ir2_diag, ir4_diag):
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        a = np.ascontiguousarray(ir2_diag, dtype=np.float64)
        b = np.ascontiguousarray(ir4_diag, dtype=np.float64)
        if D2.shape != (self.nr, self.nr) or a.shape != (self.nr,) or b.shape != (self.nr,):
            raise ValueError("A4 auxiliary arrays have the wrong shape")
        self._check(self.lib.sddc_plan_set_a4_aux(self._h, D2.ctypes.data, a.ctypes.data, b.ctypes.data))

    def new_state(self, B):
        return torch.empty((B, 3 * self.N), dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ hot path
    def nlin_fx(self, X, out=None):
        X = self._in(X, 3 * self.N)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_nlin_fx(self._h, X.data_ptr(), out.data_ptr(), X.shape[0], self._stream()))
        return out

    def nlin_dfx(self, dv, X, out=None):
        dv, X = self._in(dv, 3 * self.N), self._in(X, 3 * self.N)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_nlin_dfx(self._h, dv.data_ptr(), X.data_ptr(), out.data_ptr(), X.shape[0], self._stream()))
        return out

    def linear_op(self, op, f, out=None):
        f = self._in(f, self.N)
        out = self._out(out, f.shape)
        self._check(self.lib.sddc_linear_op(self._h, int(op), f.data_ptr(), out.data_ptr(), f.shape[0], self._stream()))
        return out

    def solve_a4(self, g, out=None):
        g = self._in(g, self.N)
        out = self._out(out, g.shape)
        self._check(self.lib.sddc_solve_a4(self._h, g.data_ptr(), out.data_ptr(), g.shape[0], self._stream()))
        return out

    def solve_nab2(self, g, which, out=None):
        g = self._in(g, self.N)
        out = self._out(out, g.shape)
        self._check(self.lib.sddc_solve_nab2(self._h, int(which), g.data_ptr(), out.data_ptr(), g.shape[0], self._stream()))
        return out

    def step(self, X, Ra, Ra_s, nsteps=1, linear=False, out=None):
        X = self._in(X, 3 * self.N)
        B = X.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_step(self._h, X.data_ptr(), out.data_ptr(), Ra.data_ptr(), Ra_s.data_ptr(), B,
                                       int(nsteps), int(bool(linear)), self._stream()))
        return out

    def time_step(self, X, Ra, Ra_s, nsteps, diag_every=1, linear=False, out=None):
        """The loop of Main._Time_Step (Main.py:286-329), device resident: nsteps member-steps and the diagnostics of
        every diag_every-th step.  Returns (X_new, history [nsteps // diag_every, B, 6]) as device tensors; one
        stream-ordered C-ABI call (the kinetic energy of step s rides on the prep stage of step s+1)."""
        X = self._in(X, 3 * self.N)
        B = X.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        out = self._out(out, X.shape)
        nrec = int(nsteps) // int(diag_every) if diag_every else 0
        hist = torch.empty((nrec, B, 6), dtype=torch.float64, device=self.device)
        self._check(self.lib.sddc_time_step(self._h, X.data_ptr(), out.data_ptr(), Ra.data_ptr(), Ra_s.data_ptr(), B,
                                            int(nsteps), int(bool(linear)), int(diag_every),
                                            hist.data_ptr() if nrec else None, self._stream()))
        return out, hist

    def residual(self, X, Ra, Ra_s, out=None):
        X = self._in(X, 3 * self.N)
        B = X.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_residual(self._h, X.data_ptr(), out.data_ptr(), Ra.data_ptr(), Ra_s.data_ptr(), B, self._stream()))
        return out

    def jvp(self, dv, X, Ra, Ra_s, out=None):
        dv, X = self._in(dv, 3 * self.N), self._in(X, 3 * self.N)
        B = X.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_jvp(self._h, dv.data_ptr(), X.data_ptr(), out.data_ptr(), Ra.data_ptr(),
                                      Ra_s.data_ptr(), B, self._stream()))
        return out

    def jvp_set_base(self, X):
        """Cache the base state X of the following jvp_apply calls (one synthesis of X instead of one per product)."""
        X = self._in(X, 3 * self.N)
        self._check(self.lib.sddc_jvp_set_base(self._h, X.data_ptr(), X.shape[0], self._stream()))
        self._base_B = X.shape[0]

    def jvp_apply(self, dv, Ra, Ra_s, out=None, plus_identity=False):
        """PDFX(dv, X_base) for the state given to jvp_set_base; plus_identity=True returns PDFX(dv) + dv (the
        linearised step itself: no subtrahend in the back-substitution), for Krylov solvers that handle the shift."""
        dv = self._in(dv, 3 * self.N)
        B = dv.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        out = self._out(out, dv.shape)
        fn = self.lib.sddc_jvp_apply_plus if plus_identity else self.lib.sddc_jvp_apply
        self._check(fn(self._h, dv.data_ptr(), out.data_ptr(), Ra.data_ptr(), Ra_s.data_ptr(), B, self._stream()))
        return out

    @property
    def has_jvp_plus(self):
        """True when jvp_apply(plus_identity=True) is available (every shape with the row pipeline for two-state products
        or the dense persistent synthesis)."""
        i = self.info()
        return bool(i["fft_jvp"] or i["synth_variant"])

    def dF_dRa(self, X, out=None):
        X = self._in(X, 3 * self.N)
        out = self._out(out, X.shape)
        self._check(self.lib.sddc_dF_dRa(self._h, X.data_ptr(), out.data_ptr(), X.shape[0], self._stream()))
        return out

    def diagnostics(self, X, out=None):
        """[B, 6]: ||X||_2, KE, Nu_T, Nu_S, Nu_T(outer wall), Nu_S(outer wall)."""
        X = self._in(X, 3 * self.N)
        out = self._out(out, (X.shape[0], 6))
        self._check(self.lib.sddc_diagnostics(self._h, X.data_ptr(), out.data_ptr(), X.shape[0], self._stream()))
        return out

    # ------------------------------------------------------------------ host-buffer entry points (NumPy in / out)
    @staticmethod
    def pinned(shape):
        """Page-locked float64 host array (NumPy view of a pinned torch tensor) for the *_host entry points."""
        t = torch.empty(tuple(shape), dtype=torch.float64).pin_memory()
        a = t.numpy()
        _PINNED_KEEPALIVE[id(a)] = t
        return a

    def step_host(self, X, Ra, Ra_s, nsteps=1, linear=False, want_diag=False, out=None, diag_out=None):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3 * self.N)
        B = X.shape[0]
        Ra = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra, dtype=np.float64), (B,)))
        Ra_s = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra_s, dtype=np.float64), (B,)))
        if out is None:
            out = np.empty_like(X)
        elif out.shape != X.shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % (X.shape,))
        diag = (np.empty((B, 6)) if diag_out is None else diag_out) if want_diag else None
        self._check(self.lib.sddc_step_host(self._h, X.ctypes.data, out.ctypes.data, Ra.ctypes.data, Ra_s.ctypes.data,
                                            B, int(nsteps), int(bool(linear)), diag.ctypes.data if want_diag else None))
        return (out, diag) if want_diag else out

    def time_step_host(self, X, Ra, Ra_s, nsteps, diag_every=1, ckpt_every=0, linear=False, out=None, diag_hist=None,
                       ckpt=None, ckpt_first=None):
        """Ensemble analogue of Main._Time_Step from NumPy buffers: returns (X_final, diag_hist[, checkpoints]) with
        diag_hist [nsteps // diag_every, B, 6] and checkpoints after the steps ckpt_first, ckpt_first + ckpt_every, ...
        ckpt_first defaults to ckpt_every ([nsteps // ckpt_every, B, 3N] records); ckpt_first=1 gives the reference's own
        checkpoints, which Main._Time_Step takes after the steps 1, 1 + N_save, ... (Main.py:301-303)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3 * self.N)
        B = X.shape[0]
        Ra = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra, dtype=np.float64), (B,)))
        Ra_s = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra_s, dtype=np.float64), (B,)))
        if out is None:
            out = np.empty_like(X)
        elif out.shape != X.shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % (X.shape,))
        nrec = nsteps // diag_every if diag_every else 0
        first = int(ckpt_every if ckpt_first is None else ckpt_first)
        if ckpt_every and not (1 <= first):
            raise ValueError("ckpt_first must be >= 1")
        nck = ((nsteps - first) // ckpt_every + 1 if nsteps >= first else 0) if ckpt_every else 0
        self._check(self.lib.sddc_plan_set_ckpt_phase(self._h, first if ckpt_every else 0))
        if diag_hist is None:
            diag_hist = np.empty((nrec, B, 6))
        if ckpt is None and nck:
            ckpt = np.empty((nck, B, 3 * self.N))
        if diag_hist.shape != (nrec, B, 6) or (nck and ckpt.shape != (nck, B, 3 * self.N)):
            raise ValueError("history / checkpoint buffers have the wrong shape")
        for buf in (diag_hist, ckpt):
            if buf is not None and (buf.dtype != np.float64 or not buf.flags.c_contiguous):
                raise ValueError("history / checkpoint buffers must be C-contiguous float64 arrays")
        self._check(self.lib.sddc_time_step_host(
            self._h, X.ctypes.data, out.ctypes.data, Ra.ctypes.data, Ra_s.ctypes.data, B, int(nsteps),
            int(bool(linear)), int(diag_every), diag_hist.ctypes.data if nrec else None, int(ckpt_every),
            ckpt.ctypes.data if nck else None))
        return (out, diag_hist, ckpt) if nck else (out, diag_hist)

    def jvp_host(self, dv, X, Ra, Ra_s):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3 * self.N)
        dv = np.ascontiguousarray(dv, dtype=np.float64).reshape(-1, 3 * self.N)
        B = X.shape[0]
        Ra = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra, dtype=np.float64), (B,)))
        Ra_s = np.ascontiguousarray(np.broadcast_to(np.asarray(Ra_s, dtype=np.float64), (B,)))
        out = np.empty_like(X)
        self._check(self.lib.sddc_jvp_host(self._h, dv.ctypes.data, X.ctypes.data, out.ctypes.data, Ra.ctypes.data,
                                           Ra_s.ctypes.data, B))
        return out


def transform(kind, x, n=None):
    """Batched Transforms.IDCT/IDST/DCT/DST on the last axis of a float64 CUDA tensor."""
    lib = _lib.load()
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64):
        raise TypeError("expected a float64 CUDA tensor")
    n_in = x.shape[-1]
    n_out = n_in if n is None else int(n)
    if kind in (T_DCT, T_DST):
        n_out = min(n_out, n_in)
    xc = x.contiguous().reshape(-1, n_in)
    out = torch.empty((xc.shape[0], n_out), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.sddc_transform(int(kind), xc.data_ptr(), out.data_ptr(), xc.shape[0], n_in, n_out,
                                C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != 0:
        raise SddcError("sddc_transform failed (%d)" % rc)
    return out.reshape(x.shape[:-1] + (n_out,))
