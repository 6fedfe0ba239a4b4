"""CPU check of the FFT formulation used by nlin_fft_kernel: the per-thread phase functions of csrc/fft_core.h are
compiled for the host (tests/fft_emul.cpp) and run thread by thread against the oracle's dense-sum transforms
(Transforms.py:73-129, Matrix_Operators.py:776-799 / 884-887)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import sddc_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fft_emul") / "libfft_emul.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                    os.path.join(HERE, "fft_emul.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.fft_emul_butterfly_error.restype = ctypes.c_double
    dp = ctypes.POINTER(ctypes.c_double)
    lib.fft_emul_rows.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_int]
    lib.fft_emul_ke_rows.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int]
    lib.fft_emul_rows_staged.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int]
    return lib


def _rows(X3, op, symmetric):
    """[n][7][K] stored rows (JT, omega, DT, Dpsi, DS, -kT, -kS) and the nine grid-ready arrays of the oracle."""
    c, s = orc._spectral_fields(X3, op, symmetric)
    K, n = op.K, op.n
    if symmetric:
        X3 = X3 * orc.sym_mask(K, n)
    kk = -np.arange(K, dtype=np.float64)[None, :]
    rows = np.stack([c["JT"], s["om"], c["DT"], s["Dpsi"], c["DS"], kk * X3[1].T, kk * X3[2].T], axis=1)
    return np.ascontiguousarray(rows)


def _expected(g, h=None):
    """[n][4][K] analysed products (Matrix_Operators.py:791-799; bilinear form 884-887 when h is given)."""
    if h is None:
        P1 = g["JT"] * g["om"]
        P2 = g["kDpsi"] * g["om"] + g["Dpsi"] * g["kom"]
        NT = g["JT"] * g["DT"] - g["Dpsi"] * g["kT"]
        NS = g["JT"] * g["DS"] - g["Dpsi"] * g["kS"]
    else:
        P1 = g["JT"] * h["om"] + h["JT"] * g["om"]
        P2 = (g["kDpsi"] * h["om"] + g["Dpsi"] * h["kom"]) + (h["kDpsi"] * g["om"] + h["Dpsi"] * g["kom"])
        NT = (h["JT"] * g["DT"] - h["Dpsi"] * g["kT"]) + (g["JT"] * h["DT"] - g["Dpsi"] * h["kT"])
        NS = (h["JT"] * g["DS"] - h["Dpsi"] * g["kS"]) + (g["JT"] * h["DS"] - g["Dpsi"] * h["kS"])
    K = (2 * P1.shape[1]) // 3
    return np.stack([orc.DST(P1)[:, :K], orc.DST(P2)[:, :K], orc.DCT(NT)[:, :K], orc.DCT(NS)[:, :K]], axis=1)


def _call(lib, M, rows0, rows1=None, cached=False):
    n, _, K = rows0.shape
    out = np.full((n, 4, K), np.nan)
    dp = ctypes.POINTER(ctypes.c_double)
    r1 = rows1 if rows1 is not None else rows0
    rc = lib.fft_emul_rows(M, 2 if cached else int(rows1 is not None), rows0.ctypes.data_as(dp), r1.ctypes.data_as(dp),
                           out.ctypes.data_as(dp), n)
    assert rc == 0
    return _natural(out)


def _natural(out):
    """The kernels store every [K] row parity-split (even sinusoid indices first; fft_core.h: spec_pos)."""
    K = out.shape[-1]
    k = np.arange(K)
    return np.ascontiguousarray(out[..., (k & 1) * (K // 2) + (k >> 1)])


def test_register_butterflies(emul):
    assert emul.fft_emul_butterfly_error() < 5e-14


@pytest.mark.parametrize("K,N_r,symmetric", [(128, 8, False), (256, 6, False), (256, 6, True), (512, 5, False)])
def test_fft_rows_match_dense_transforms(emul, K, N_r, symmetric):
    op = orc.Operators(K, N_r, 0.4, 1e-2, 1.0, 0.5)
    rng = np.random.default_rng(K + N_r)
    X3 = rng.standard_normal((3, K, op.n))
    Y3 = rng.standard_normal((3, K, op.n))
    M = 3 * K // 2
    rows = _rows(X3, op, symmetric)
    g = orc._grid_fields(X3, op, symmetric)
    exp = _expected(g)
    got = _call(emul, M, rows)
    assert np.isfinite(got).all()
    scale = np.abs(exp).max(axis=(0, 2), keepdims=True)
    assert (np.abs(got - exp) / scale).max() < 1e-12  # white spectra times k: the dense sums themselves round at this level
    if K == 256:
        # the staged schedule of the headline kernel (per-warp ownership, coefficient rows staged in the dead planes)
        # performs the same arithmetic: bit-identical
        got_s = np.full(got.shape, np.nan)
        dp = ctypes.POINTER(ctypes.c_double)
        assert emul.fft_emul_rows_staged(M, rows.ctypes.data_as(dp), got_s.ctypes.data_as(dp), rows.shape[0]) == 0
        assert np.array_equal(_natural(got_s), got)
    # two-state (Jacobian-vector product) variant
    rows1 = _rows(Y3, op, symmetric)
    h = orc._grid_fields(Y3, op, symmetric)
    exp2 = _expected(g, h)
    got2 = _call(emul, M, rows, rows1)
    scale2 = np.abs(exp2).max(axis=(0, 2), keepdims=True)
    assert (np.abs(got2 - exp2) / scale2).max() < 1e-12
    # the same products with the base state's grid fields synthesised once and cached (sddc_jvp_set_base / _apply)
    got3 = _call(emul, M, rows, rows1, cached=True)
    assert (np.abs(got3 - exp2) / scale2).max() < 1e-12


@pytest.mark.parametrize("K", [128, 256, 512])
def test_fft_kinetic_energy_rows(emul, K):
    """Weighted sum of squares of the two Kinetic_Energy fields on the 3K grid (Main.py:104-130)."""
    rng = np.random.default_rng(K)
    n, M3 = 7, 3 * K
    a = rng.standard_normal((n, K))
    b = rng.standard_normal((n, K))
    b[:, 0] = 0.0
    rows = np.ascontiguousarray(np.stack([a, b], axis=1))
    out = np.full(n, np.nan)
    dp = ctypes.POINTER(ctypes.c_double)
    assert emul.fft_emul_ke_rows(M3, rows.ctypes.data_as(dp), out.ctypes.data_as(dp), n) == 0
    th = orc.grid(M3)
    wt = np.zeros(M3)
    dth = np.diff(th)
    wt[:-1] += 0.5 * dth
    wt[1:] += 0.5 * dth
    ref = ((orc.IDCT(a, n=M3) ** 2 + orc.IDST(b, n=M3) ** 2) * (wt * np.sin(th))[None, :]).sum(axis=1)
    assert np.abs(out / ref - 1).max() < 1e-13


def _at(j, c):
    """fft_core.h at(): element c of block j of a plane."""
    return ((j ^ ((j >> 3) & 1)) << 3) | (c ^ (j & 7))


def _wavefronts(addrs):
    """Shared-memory wavefronts of one 64-bit warp access (two half-warps; bank = address mod 16 in doubles)."""
    tot = 0
    for h in range(2):
        lanes = {a for a in addrs[16 * h:16 * h + 16] if a is not None}
        if lanes:
            banks = {}
            for a in lanes:
                banks.setdefault(a % 16, set()).add(a)
            tot += max(len(v) for v in banks.values())
    return tot


@pytest.mark.parametrize("M", [384, 768])
def test_shared_memory_layout_is_conflict_free(M):
    """The XOR-swizzled block layout of fft_core.h keeps the radix-8, radix-RD and packing passes free of bank
    conflicts and the radix-6 pass within 4/3 of the ideal wavefront count."""
    NBLK, L, RD, NT = M // 8, M // 6, M // 48, (128 if M == 768 else 64)
    for warp in range(NT // 32):
        lanes = [warp * 32 + l for l in range(32)]
        # radix-8 pass: lane <-> block (five transforms back to back), one element index at a time
        for r in range(0, 5 * NBLK, NT):
            us = [t + r if t + r < 5 * NBLK else None for t in lanes]
            for c in range(8):
                ad = [None if u is None else 2 * (u // NBLK) * M + _at(u % NBLK, c) for u in us]
                assert _wavefronts(ad) == sum(1 for h in range(2) if any(a is not None for a in ad[16 * h:16 * h + 16]))
        # radix-RD pass: lane <-> (k2, element a), blocks 6 d + k2
        for r in range(0, 5 * 48, NT):
            us = [t + r if t + r < 5 * 48 else None for t in lanes]
            for d in range(RD):
                ad = [None if u is None else 2 * (u // 48) * M + _at(6 * d + ((u % 48) >> 3), (u % 48) & 7) for u in us]
                assert _wavefronts(ad) == sum(1 for h in range(2) if any(a is not None for a in ad[16 * h:16 * h + 16]))
        # packing: consecutive wavenumbers k -> consecutive blocks
        for r in range(0, M // 2 + 1, NT):
            ad = [_at((t + r) % NBLK, (t + r) // NBLK) if t + r <= M // 2 else None for t in lanes]
            assert _wavefronts(ad) == sum(1 for h in range(2) if any(a is not None for a in ad[16 * h:16 * h + 16]))
        # radix-6 pass: lane <-> column n1, blocks 6 (n1 >> 3) + m: at most 2-way, 4/3 of ideal overall
        got = ideal = 0
        for r in range(0, L, NT):
            ns = [t + r if t + r < L else None for t in lanes]
            for m in range(6):
                ad = [None if n is None else _at(6 * (n >> 3) + m, n & 7) for n in ns]
                got += _wavefronts(ad)
                ideal += sum(1 for h in range(2) if any(a is not None for a in ad[16 * h:16 * h + 16]))
        assert got <= ideal * 4 / 3 + 1e-9
