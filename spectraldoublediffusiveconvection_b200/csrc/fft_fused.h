// Second generation of the FFT formulation of NLIN_FX / NLIN_DFX (Matrix_Operators.py:743-898): the phase functions of
// nlin_fft_kernel (k_nlin_fft.cuh).  Like fft_core.h they compile for the device and for the host
// (tests/fft_emul.cpp runs them thread by thread against the test-only oracle).
//
// What changed against the first generation (fft_core.h: build | radix-8 | radix-RD | radix-6 ... | post):
//
//  * Two syntheses fewer.  k Dpsi and k omega are the theta-derivatives of the sine series Dpsi and omega, so the second
//    stream-function product of Matrix_Operators.py:791 is a derivative,
//          kDpsi * omega + Dpsi * komega = d/dtheta (Dpsi * omega),
//    and its sine analysis is  DST(.)_k = -k DCT(Dpsi * omega)_k  (exact on the 3/2-padded grid for k < K: the product
//    has degree <= 2K-2 and aliases fold onto wavenumbers >= K+2).  Seven real fields per state are left
//          cosine type: JT, DT, DS          sine type: omega, Dpsi, -kT, -kS
//    = four complex transforms (JT|omega) (DT|Dpsi) (DS|-kT) (0|-kS) instead of five; a pair of states (JVP) needs
//    seven instead of nine: 3 + 3 + (-kS | -kS').
//  * The packing phase feeds the radix-8 pass from registers.  A thread builds the Hermitian-packed spectrum of a block
//    pair (j, NB - j) -- the spectral mirror k -> M - k maps element c of block j to element 7 - c of block NB - j --
//    and transforms both blocks before anything is stored: one shared-memory round trip per transform less.  Mirror
//    partners beyond the truncated spectrum (k <= M/3) are compile-time zeros: five of the eight pairs of a unit take
//    the 6-flop form instead of the 12-flop one.
//  * The same on the way back: the forward radix-8 pass hands its two blocks to the separation / scaling code in
//    registers and the results go straight to HBM.
//
// Shared-memory traffic per radial row at M = 384: 144 KB (252 KB before), fp64 instructions about 18 % fewer.
#pragma once
#include "fft_core.h"

namespace sddc {
namespace fftp {

SDDC_HD void store_block(double* __restrict__ re, double* __restrict__ im, int j, const C (&y)[8]) {
    const int base = (j ^ ((j >> 3) & 1)) << 3, sw = j & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        re[base + (c ^ sw)] = y[c].r;
        im[base + (c ^ sw)] = y[c].i;
    }
}
SDDC_HD void load_block(const double* __restrict__ re, const double* __restrict__ im, int j, C (&x)[8]) {
    const int base = (j ^ ((j >> 3) & 1)) << 3, sw = j & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = C{re[base + (c ^ sw)], im[base + (c ^ sw)]};
}

// Stored rows of one radial point (k_prep.cuh, FFTL): 0 JT, 1 omega, 2 DT, 3 Dpsi, 4 DS, 5 -k T, 6 -k S (the factor -k of
// the theta-derivative of a cosine series, Matrix_Operators.py:711-716, is applied by the prep stage).
// Transform q of a state packs (cosine-type row | sine-type row): 0 (JT | omega)  1 (DT | Dpsi)  2 (DS | -kT)  3 (- | -kS):
// rows (2q, 2q + 1), so that what a warp needs for its transforms is contiguous (TMA staging in nlin_fft_staged_kernel).
SDDC_HD constexpr int row_a(int q) { return q < 3 ? 2 * q : -1; }
SDDC_HD constexpr int row_b(int q) { return q < 3 ? 2 * q + 1 : 6; }

// TYPE 0: (cosine | sine).  TYPE 1: (sine | sine).  HASA 0: no field a, 1: field a present (both compile time),
// 2: decided per thread (a == nullptr: none).
struct PackSrc {
    const double* a;
    const double* b;
};
template <int K>
SDDC_HD PackSrc pack_src(int q, const double* __restrict__ rows) {
    return PackSrc{q < 3 ? rows + row_a(q) * K : nullptr, rows + row_b(q) * K};
}
template <int HASA>
SDDC_HD double ld_a(const PackSrc& s, int k) {
    if (HASA == 0) return 0.0;
    if (HASA == 1) return s.a[k];
    return s.a ? s.a[k] : 0.0;
}

// Packed spectrum at kappa and M - kappa (1 <= kappa <= M/2) from the raw coefficients at kappa (a, s) and, when the mirror
// index lies inside the truncated spectrum (FULL), at M - kappa (ap, sp).  DCT-III inputs: cosine type X_k = c_k, sine
// type X_k = s_{M-k}; Z_k = w_k [(X^a_k - i X^a_{M-k}) + i (X^b_k - i X^b_{M-k})] / 2 (fft_core.h, build()).
template <int TYPE, bool FULL, bool HAS_A>
SDDC_HD void pack_pair(double a, double s, double ap, double sp, double wc, double ws, C& zk, C& zkp) {
    if (!HAS_A) {
        // the sine-type field alone (either slot: the other one is identically zero)
        if (FULL) {
            zk = C{wc * s - ws * sp, wc * sp + ws * s};
            zkp = C{ws * sp - wc * s, ws * s + wc * sp};
        } else {
            zk = C{wc * s, ws * s};
            zkp = C{-(wc * s), ws * s};
        }
    } else if (TYPE == 0) {
        if (FULL) {
            const double P = a + s, Q = sp - ap, P2 = ap + sp, Q2 = s - a;
            zk = C{wc * P - ws * Q, wc * Q + ws * P};
            zkp = C{ws * P2 - wc * Q2, ws * Q2 + wc * P2};
        } else {
            const double P = a + s, Q2 = s - a;
            zk = C{wc * P, ws * P};
            zkp = C{-(wc * Q2), ws * Q2};
        }
    } else {
        // both sine type: X^a_k = a_{M-k}, X^a_{M-k} = a_k (a, ap are the raw sine coefficients at kappa, M - kappa)
        if (FULL) {
            const double P = ap + s, Q = sp - a, P2 = a + sp, Q2 = s - ap;
            zk = C{wc * P - ws * Q, wc * Q + ws * P};
            zkp = C{ws * P2 - wc * Q2, ws * Q2 + wc * P2};
        } else {
            zk = C{wc * s + ws * a, ws * s - wc * a};
            zkp = C{ws * a - wc * s, ws * s + wc * a};
        }
    }
}

// ---- bc: packing + inverse radix-8 pass of the block pair (j, NB - j), 1 <= j <= NB/2 ----------------------------------
template <int M, int TYPE, int HASA>
SDDC_HD void bc_unit(int j, const PackSrc& s, double* __restrict__ re, double* __restrict__ im, const Tables& tb) {
    constexpr int NB = Cfg<M>::NBLK, K = Cfg<M>::K;
    constexpr bool HA = HASA != 0;
    const int jb = NB - j;
    C zA[8], zB[8];
    // pairs whose lower index lies in block j: kappa = j + NB c <= M/2, mirror = element 7 - c of block NB - j
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = j + NB * c, kp = M - k;
        const double a = ld_a<HASA>(s, k), b = s.b[k];
        const double wc = tb.wkc[k], ws = tb.wks[k];
        if (c == 3) {   // kappa > M/3: the mirror index is inside the truncated spectrum
            pack_pair<TYPE, true, HA>(a, b, ld_a<HASA>(s, kp), s.b[kp], wc, ws, zA[c], zB[7 - c]);
        } else {
            pack_pair<TYPE, false, HA>(a, b, 0.0, 0.0, wc, ws, zA[c], zB[7 - c]);
        }
    }
    // pairs whose lower index lies in block NB - j
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = jb + NB * c, kp = M - k;
        const double a = ld_a<HASA>(s, k), b = s.b[k];
        const double wc = tb.wkc[k], ws = tb.wks[k];
        if (c == 3) {
            pack_pair<TYPE, true, HA>(a, b, ld_a<HASA>(s, kp), s.b[kp], wc, ws, zB[c], zA[7 - c]);
        } else if (c == 2) {
            // kappa = NB - j + 2 NB > M/3 only for j < NB/3: the mirror loads are predicated per thread
            const bool in = kp < K;
            const double ap = in ? ld_a<HASA>(s, kp) : 0.0, bp = in ? s.b[kp] : 0.0;
            pack_pair<TYPE, true, HA>(a, b, ap, bp, wc, ws, zB[c], zA[7 - c]);
        } else {
            pack_pair<TYPE, false, HA>(a, b, 0.0, 0.0, wc, ws, zB[c], zA[7 - c]);
        }
    }
    C y[8];
    Dft<8, +1>::run(zA, y);
    store_block(re, im, j, y);
    Dft<8, +1>::run(zB, y);
    store_block(re, im, jb, y);   // j == NB/2: the same block and the same values a second time
}

// block 0 (k = NB c): mirror of element c is element 8 - c; k = 0 and k = M/2 are their own mirrors
template <int M, int TYPE, int HASA>
SDDC_HD void bc_block0(const PackSrc& s, double* __restrict__ re, double* __restrict__ im, const Tables& tb) {
    constexpr int NB = Cfg<M>::NBLK;
    constexpr bool HA = HASA != 0;
    C z[8];
    // V_0 = X_0; index 0 of a sine-type row is ignored (Transforms.py:41-54)
    z[0] = C{TYPE == 0 ? ld_a<HASA>(s, 0) : 0.0, 0.0};
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        const int k = NB * c, kp = M - k;
        const double a = ld_a<HASA>(s, k), b = s.b[k];
        const double wc = tb.wkc[k], ws = tb.wks[k];
        if (c == 3) pack_pair<TYPE, true, HA>(a, b, ld_a<HASA>(s, kp), s.b[kp], wc, ws, z[c], z[8 - c]);
        else pack_pair<TYPE, false, HA>(a, b, 0.0, 0.0, wc, ws, z[c], z[8 - c]);
    }
    {
        const int k = M / 2;
        const double a = ld_a<HASA>(s, k), b = s.b[k];
        C dummy;
        pack_pair<TYPE, true, HA>(a, b, a, b, tb.wkc[k], tb.wks[k], z[4], dummy);
    }
    C y[8];
    Dft<8, +1>::run(z, y);
    store_block(re, im, 0, y);
}

// thread that takes the block-0 unit number `i` (< nspec) of a phase with `nreg` regular units: the first lanes of the warp
// that is idle (or least loaded) in the last round of the regular loop
template <int NTH>
SDDC_HD constexpr int spec_base(int nreg) { return (((nreg % NTH) + 31) / 32 * 32) % NTH; }

// Distribution of the NTR transforms of a phase over the NTH / 32 warps of a worker with the transform index uniform
// per warp (a compile-time constant inside the branch): no per-lane row selection, no selects.  With fewer warps than
// transforms warp w takes q = w, w + NWARP, ..; otherwise WPT warps share the block pairs of one transform.
template <int NTH, int NTR>
struct WarpMap {
    static constexpr int NWARP = NTH / 32;
    static constexpr int WPT = NWARP >= NTR ? NWARP / NTR : 1;
    static SDDC_HD bool mine(int q, int warp) { return NWARP >= NTR ? (warp / WPT == q) : (q % NWARP == warp); }
    static SDDC_HD int first(int t) { return (NWARP >= NTR ? ((t >> 5) % WPT) * 32 : 0) + (t & 31); }
    static constexpr int STEP = 32 * WPT;
};

// one whole transform by the lanes `id`, id + STEP, ..: block pairs 1..NP, then block 0 by the first lane that would
// otherwise idle in the last round
template <int M, int TYPE, int HASA, int STEP>
SDDC_HD void bc_transform(int id, const PackSrc& s, double* __restrict__ re, double* __restrict__ im, const Tables& tb) {
    constexpr int NP = Cfg<M>::NBLK / 2;
    for (int j = id + 1; j <= NP; j += STEP) bc_unit<M, TYPE, HASA>(j, s, re, im, tb);
    if (id == NP % STEP) bc_block0<M, TYPE, HASA>(s, re, im, tb);
}

// One-state packing: transforms 0..3 of `rows` into plane pairs 0..3 of buf
template <int M, int NTH>
SDDC_HD void bc_inv_fx(int t, const double* __restrict__ rows, double* __restrict__ buf, const Tables& tb) {
    constexpr int NP = Cfg<M>::NBLK / 2, PL = Cfg<M>::PL, K = Cfg<M>::K, NQ = 4;
    if (M >= 384) {
        using WM = WarpMap<NTH, NQ>;
        const int warp = t >> 5, id = WM::first(t);
#define SDDC_BC_FX(Q, HASA)                                                                                           \
    if (WM::mine(Q, warp))                                                                                            \
        bc_transform<M, 0, HASA, WM::STEP>(id, PackSrc{rows + (HASA ? row_a(Q) : 0) * K, rows + row_b(Q) * K},        \
                                           buf + pair_off<M>(Q), buf + pair_off<M>(Q) + PL, tb);
        SDDC_BC_FX(0, 1) SDDC_BC_FX(1, 1) SDDC_BC_FX(2, 1) SDDC_BC_FX(3, 0)
#undef SDDC_BC_FX
        return;
    }
    // small grids: lanes of a warp work on different transforms (row selection per thread)
    for (int u = t; u < NQ * NP; u += NTH) {
        const int q = u / NP, j = u - q * NP + 1;
        bc_unit<M, 0, 2>(j, pack_src<K>(q, rows), buf + pair_off<M>(q), buf + pair_off<M>(q) + PL, tb);
    }
    const int sp = t - spec_base<NTH>(NQ * NP);
    if (sp >= 0 && sp < NQ) bc_block0<M, 0, 2>(pack_src<K>(sp, rows), buf + pair_off<M>(sp), buf + pair_off<M>(sp) + PL, tb);
}

// Two-state packing: plane pairs 0..2 base state, 3..5 perturbation, 6 = (-kS | -kS') of both
template <int M, int NTH>
SDDC_HD void bc_inv_dfx(int t, const double* __restrict__ rows0, const double* __restrict__ rows1, double* __restrict__ buf,
                        const Tables& tb) {
    constexpr int NP = Cfg<M>::NBLK / 2, PL = Cfg<M>::PL, K = Cfg<M>::K;
    const PackSrc ss{rows0 + 6 * K, rows1 + 6 * K};
    if (M >= 384) {
        using WM = WarpMap<NTH, 7>;
        const int warp = t >> 5, id = WM::first(t);
#define SDDC_BC_DFX(Q, QS, ROWS)                                                                                      \
    if (WM::mine(Q, warp))                                                                                            \
        bc_transform<M, 0, 1, WM::STEP>(id, PackSrc{ROWS + row_a(QS) * K, ROWS + row_b(QS) * K}, buf + pair_off<M>(Q), \
                                        buf + pair_off<M>(Q) + PL, tb);
        SDDC_BC_DFX(0, 0, rows0) SDDC_BC_DFX(1, 1, rows0) SDDC_BC_DFX(2, 2, rows0)
        SDDC_BC_DFX(3, 0, rows1) SDDC_BC_DFX(4, 1, rows1) SDDC_BC_DFX(5, 2, rows1)
#undef SDDC_BC_DFX
        if (WM::mine(6, warp)) bc_transform<M, 1, 1, WM::STEP>(id, ss, buf + pair_off<M>(6), buf + pair_off<M>(6) + PL, tb);
        return;
    }
    for (int u = t; u < 7 * NP; u += NTH) {
        const int q = u / NP, j = u - q * NP + 1;
        if (q < 6) bc_unit<M, 0, 1>(j, pack_src<K>(q < 3 ? q : q - 3, q < 3 ? rows0 : rows1), buf + pair_off<M>(q), buf + pair_off<M>(q) + PL, tb);
        else bc_unit<M, 1, 1>(j, ss, buf + pair_off<M>(6), buf + pair_off<M>(6) + PL, tb);
    }
    const int sp = t - spec_base<NTH>(7 * NP);
    if (sp >= 0 && sp < 6) {
        const PackSrc s = pack_src<K>(sp < 3 ? sp : sp - 3, sp < 3 ? rows0 : rows1);
        bc_block0<M, 0, 1>(s, buf + pair_off<M>(sp), buf + pair_off<M>(sp) + PL, tb);
    } else if (sp == 6) {
        bc_block0<M, 1, 1>(ss, buf + pair_off<M>(6), buf + pair_off<M>(6) + PL, tb);
    }
}

// twiddles e^{2 pi i k2 n1 / M}, k2 = 1..5, of the radix-6 column n1
template <int M>
SDDC_HD void load_tw6(int n1, const Tables& tb, C (&t6)[5]) {
#pragma unroll
    for (int k2 = 1; k2 < 6; ++k2) t6[k2 - 1] = C{tb.t6c[(k2 - 1) * Cfg<M>::L + n1], tb.t6s[(k2 - 1) * Cfg<M>::L + n1]};
}
SDDC_HD void inv6t(const double* __restrict__ re, const double* __restrict__ im, const int (&pos)[6], const C (&t6)[5],
                   C (&z)[6]) {
    C x[6];
#pragma unroll
    for (int k2 = 0; k2 < 6; ++k2) {
        x[k2] = C{re[pos[k2]], im[pos[k2]]};
        if (k2 > 0) x[k2] = cmul(x[k2], t6[k2 - 1].r, t6[k2 - 1].i);
    }
    dft6<+1>(x, z);
}
// forward radix-6 of (xr + i xi) over the six points of a column, twiddle e^{-2 pi i k2 n1 / M}, into the planes (re, im)
SDDC_HD void fwd6t(const double (&xr)[6], const double (&xi)[6], double* __restrict__ re, double* __restrict__ im,
                   const int (&pos)[6], const C (&t6)[5]) {
    C x[6], y[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) x[m] = C{xr[m], xi[m]};
    dft6<-1>(x, y);
#pragma unroll
    for (int k2 = 0; k2 < 6; ++k2) {
        if (k2 > 0) y[k2] = cmulc(y[k2], t6[k2 - 1].r, t6[k2 - 1].i);
        re[pos[k2]] = y[k2].r;
        im[pos[k2]] = y[k2].i;
    }
}

// ---- i3f1: last inverse pass + Jacobian products + first forward pass (one state) -------------------------------------
// in : plane pairs 0..3 = (JT|om) (DT|Dpsi) (DS|-kT) (0|-kS) after the radix-8 and radix-RD passes
// out: plane pair 0 <- JT*om + i Dpsi*om   (sine | cosine type),  plane pair 1 <- N_T + i N_S  (cosine | cosine),
//      N_T = JT*DT - Dpsi*kT, N_S = JT*DS - Dpsi*kS (Matrix_Operators.py:791-793), through the forward radix-6 pass.
// The (-1)^j signs of the sine-type fields cancel in the three cosine-type products and turn the first one into the
// DCT-II input of a sine analysis (fft_core.h).
template <int M, int NTH>
SDDC_HD void i3f1_fx(int t, double* __restrict__ buf, const Tables& tb) {
    constexpr int L = Cfg<M>::L, PL = Cfg<M>::PL;
    for (int n1 = t; n1 < L; n1 += NTH) {
        int pos[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pos[m] = at(6 * (n1 >> 3) + m, n1 & 7);
        C t6[5];
        load_tw6<M>(n1, tb, t6);
        double jt[6], dp[6], NT[6];
        {
            C z0[6], z1[6];
            inv6t(buf, buf + PL, pos, t6, z0);               // JT | omega
            inv6t(buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6, z1);  // DT | Dpsi
            double P1[6], Q[6];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                jt[m] = z0[m].r;
                dp[m] = z1[m].i;
                P1[m] = jt[m] * z0[m].i;
                Q[m] = dp[m] * z0[m].i;
                NT[m] = jt[m] * z1[m].r;
            }
            fwd6t(P1, Q, buf, buf + PL, pos, t6);   // only this thread touches these positions of plane pair 0
        }
        double NS[6];
        {
            C z2[6];
            inv6t(buf + pair_off<M>(2), buf + pair_off<M>(2) + PL, pos, t6, z2);  // DS | -kT
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                NT[m] -= dp[m] * z2[m].i;
                NS[m] = jt[m] * z2[m].r;
            }
        }
        {
            C z3[6];
            inv6t(buf + pair_off<M>(3), buf + pair_off<M>(3) + PL, pos, t6, z3);  // 0 | -kS
#pragma unroll
            for (int m = 0; m < 6; ++m) NS[m] -= dp[m] * z3[m].i;
        }
        fwd6t(NT, NS, buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6);
    }
}

// ---- two-state products (NLIN_DFX, Matrix_Operators.py:884-887) -------------------------------------------------------
// in : plane pairs 0..2 base (JT|om) (DT|Dpsi) (DS|-kT), 3..5 the same of the perturbation, 6 = (-kS | -kS')
template <int M, int NTH>
SDDC_HD void i3f1_dfx(int t, double* __restrict__ buf, const Tables& tb) {
    constexpr int L = Cfg<M>::L, PL = Cfg<M>::PL;
    for (int n1 = t; n1 < L; n1 += NTH) {
        int pos[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pos[m] = at(6 * (n1 >> 3) + m, n1 & 7);
        C t6[5];
        load_tw6<M>(n1, tb, t6);
        double jt[6], jtp[6], dp[6], dpp[6], NT[6];
        {
            double P1[6], Q[6];
            C z0[6], z0p[6];
            inv6t(buf, buf + PL, pos, t6, z0);                 // JT | omega
            inv6t(buf + pair_off<M>(3), buf + pair_off<M>(3) + PL, pos, t6, z0p);   // JT' | omega'
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                jt[m] = z0[m].r;
                jtp[m] = z0p[m].r;
                P1[m] = jt[m] * z0p[m].i + jtp[m] * z0[m].i;
            }
            {
                C z1[6], z1p[6];
                inv6t(buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6, z1);    // DT | Dpsi
                inv6t(buf + pair_off<M>(4), buf + pair_off<M>(4) + PL, pos, t6, z1p);   // DT' | Dpsi'
#pragma unroll
                for (int m = 0; m < 6; ++m) {
                    dp[m] = z1[m].i;
                    dpp[m] = z1p[m].i;
                    Q[m] = dp[m] * z0p[m].i + dpp[m] * z0[m].i;
                    NT[m] = jt[m] * z1p[m].r + jtp[m] * z1[m].r;
                }
            }
            fwd6t(P1, Q, buf, buf + PL, pos, t6);
        }
        double NS[6];
        {
            C z2[6], z2p[6];
            inv6t(buf + pair_off<M>(2), buf + pair_off<M>(2) + PL, pos, t6, z2);      // DS | -kT
            inv6t(buf + pair_off<M>(5), buf + pair_off<M>(5) + PL, pos, t6, z2p);   // DS' | -kT'
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                NT[m] -= dp[m] * z2p[m].i + dpp[m] * z2[m].i;
                NS[m] = jt[m] * z2p[m].r + jtp[m] * z2[m].r;
            }
        }
        {
            C z3[6];
            inv6t(buf + pair_off<M>(6), buf + pair_off<M>(6) + PL, pos, t6, z3);    // -kS | -kS'
#pragma unroll
            for (int m = 0; m < 6; ++m) NS[m] -= dp[m] * z3[m].i + dpp[m] * z3[m].r;
        }
        fwd6t(NT, NS, buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6);
    }
}

// ---- cached base state (sddc_jvp_set_base / sddc_jvp_apply) ----------------------------------------------------------------
// Inside one linear solve the base state X of PDFX(dv, X) is fixed, so its seven grid fields are synthesised ONCE
// (i3f1_grid) and kept in HBM in exactly the order the product phase consumes them: grid[field][m][n1], the value of
// `field` at grid point n1 + L m held by thread n1 (512 contiguous bytes per (field, m) at M = 384).  Every product then
// transforms only the perturbation -- four inverse transforms like the one-state kernel instead of the seven of a pair of
// states -- and reads 7 M doubles per row (i3f1_jvpc).  Field order: 0 JT, 1 omega, 2 DT, 3 Dpsi, 4 DS, 5 kT, 6 kS.
template <int M, int NTH>
SDDC_HD void i3f1_grid(int t, const double* __restrict__ buf, const Tables& tb, double* __restrict__ grid) {
    constexpr int L = Cfg<M>::L, PL = Cfg<M>::PL;
    for (int n1 = t; n1 < L; n1 += NTH) {
        int pos[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pos[m] = at(6 * (n1 >> 3) + m, n1 & 7);
        C t6[5];
        load_tw6<M>(n1, tb, t6);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            C z[6];
            inv6t(buf + pair_off<M>(q), buf + pair_off<M>(q) + PL, pos, t6, z);
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                if (q < 3) grid[((2 * q) * 6 + m) * L + n1] = z[m].r;
                grid[((q < 3 ? 2 * q + 1 : 6) * 6 + m) * L + n1] = z[m].i;
            }
        }
    }
}

// in : plane pairs 0..3 = the perturbation's (JT'|om') (DT'|Dpsi') (DS'|-kT') (0|-kS'); grid = the base state's fields
// out: as i3f1_fx, with the bilinear products of Matrix_Operators.py:884-887
template <int M, int NTH>
SDDC_HD void i3f1_jvpc(int t, double* __restrict__ buf, const Tables& tb, const double* __restrict__ grid) {
    constexpr int L = Cfg<M>::L, PL = Cfg<M>::PL;
    for (int n1 = t; n1 < L; n1 += NTH) {
        int pos[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pos[m] = at(6 * (n1 >> 3) + m, n1 & 7);
        C t6[5];
        load_tw6<M>(n1, tb, t6);
        const double* g = grid + n1;
        double jt[6], dp[6], jtp[6], dpp[6], NT[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) {
            jt[m] = g[(0 * 6 + m) * L];
            dp[m] = g[(3 * 6 + m) * L];
        }
        {
            C z0[6], z1[6];
            inv6t(buf, buf + PL, pos, t6, z0);                                       // JT' | omega'
            inv6t(buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6, z1);    // DT' | Dpsi'
            double P1[6], Q[6];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const double om = g[(1 * 6 + m) * L], dT = g[(2 * 6 + m) * L];
                jtp[m] = z0[m].r;
                dpp[m] = z1[m].i;
                P1[m] = jt[m] * z0[m].i + jtp[m] * om;
                Q[m] = dp[m] * z0[m].i + dpp[m] * om;
                NT[m] = jt[m] * z1[m].r + jtp[m] * dT;
            }
            fwd6t(P1, Q, buf, buf + PL, pos, t6);
        }
        double NS[6];
        {
            C z2[6];
            inv6t(buf + pair_off<M>(2), buf + pair_off<M>(2) + PL, pos, t6, z2);    // DS' | -kT'
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const double dS = g[(4 * 6 + m) * L], kT = g[(5 * 6 + m) * L];
                NT[m] -= dp[m] * z2[m].i + dpp[m] * kT;
                NS[m] = jt[m] * z2[m].r + jtp[m] * dS;
            }
        }
        {
            C z3[6];
            inv6t(buf + pair_off<M>(3), buf + pair_off<M>(3) + PL, pos, t6, z3);    // 0 | -kS'
#pragma unroll
            for (int m = 0; m < 6; ++m) NS[m] -= dp[m] * z3[m].i + dpp[m] * g[(6 * 6 + m) * L];
        }
        fwd6t(NT, NS, buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, pos, t6);
    }
}

// ---- cp: forward radix-8 pass + separation of the packed sequences, scaling, truncation to K ---------------------------
// out: [4][K] = DST(JT*om), DST(kDpsi*om + Dpsi*kom) = -k DCT(Dpsi*om), DCT(N_T), DCT(N_S)  (sinusoid indexing;
// Transforms.py:28-39,56-70).  (A, Bv) = V at kappa, (Cc, D) = V at M - kappa; C_k = Re[conj(w_k) V_k] of the two packed
// fields, a sine-type field has its coefficient k at index M - k.
// QM: 0 / 1 = transform index known at compile time, 2 = per thread (argument q)
template <int M, bool BOTH, int QM>
SDDC_HD void cp_emit(int q, int k, C vk, C vkp, double* __restrict__ oa, double* __restrict__ ob, bool kp_ok,
                     const Tables& tb) {
    constexpr double sc = 2.0 / M;   // the table holds w_k / 2, which absorbs the 1/2 of the Hermitian split
    const int kp = M - k;
    const int pk = spec_pos(k, Cfg<M>::K), pkp = spec_pos(kp, Cfg<M>::K);   // parity-split row (fft_core.h)
    const double hc = sc * tb.wkc[k], hs = sc * tb.wks[k];
    const double s = vk.r + vkp.r, d = vk.i - vkp.i, e = vk.i + vkp.i, f = vk.r - vkp.r;
    // field a at kappa: cosine type  hc s + hs d ; sine type (its coefficient kappa sits at M - kappa)  hs s - hc d
    const bool q0 = QM == 2 ? q == 0 : QM == 0;
    const double cbk = hc * e - hs * f;
    if (QM == 2) {
        const double al = q0 ? hs : hc, be = q0 ? -hc : hs;
        oa[pk] = al * s + be * d;
        ob[pk] = q0 ? -(double)k * cbk : cbk;
    } else if (QM == 0) {
        oa[pk] = hs * s - hc * d;
        ob[pk] = -(double)k * cbk;
    } else {
        oa[pk] = hc * s + hs * d;
        ob[pk] = cbk;
    }
    if (BOTH && kp_ok) {
        const double cbkp = hs * e + hc * f;
        if (QM == 2) {
            const double al2 = q0 ? hc : hs, be2 = q0 ? hs : -hc;
            oa[pkp] = al2 * s + be2 * d;
            ob[pkp] = q0 ? -(double)kp * cbkp : cbkp;
        } else if (QM == 0) {
            oa[pkp] = hc * s + hs * d;
            ob[pkp] = -(double)kp * cbkp;
        } else {
            oa[pkp] = hs * s - hc * d;
            ob[pkp] = cbkp;
        }
    }
}

template <int M, int QM>
SDDC_HD void cp_unit(int q, int j, const double* __restrict__ re, const double* __restrict__ im, double* __restrict__ out,
                     const Tables& tb) {
    constexpr int NB = Cfg<M>::NBLK, K = Cfg<M>::K;
    const int jb = NB - j;
    C x[8], VA[8], VB[8];
    load_block(re, im, j, x);
    Dft<8, -1>::run(x, VA);
    load_block(re, im, jb, x);
    Dft<8, -1>::run(x, VB);
    double* oa = out + 2 * q * K;
    double* ob = oa + K;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = j + NB * c;
        if (c == 3) cp_emit<M, true, QM>(q, k, VA[c], VB[7 - c], oa, ob, true, tb);
        else cp_emit<M, false, QM>(q, k, VA[c], VB[7 - c], oa, ob, false, tb);
    }
    if (j != jb) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = jb + NB * c;
            if (c == 3) cp_emit<M, true, QM>(q, k, VB[c], VA[7 - c], oa, ob, true, tb);
            else if (c == 2) cp_emit<M, true, QM>(q, k, VB[c], VA[7 - c], oa, ob, M - k < K, tb);
            else cp_emit<M, false, QM>(q, k, VB[c], VA[7 - c], oa, ob, false, tb);
        }
    }
}

template <int M, int QM>
SDDC_HD void cp_block0(int q, const double* __restrict__ re, const double* __restrict__ im, double* __restrict__ out,
                       const Tables& tb) {
    constexpr int NB = Cfg<M>::NBLK, K = Cfg<M>::K;
    C x[8], V[8];
    load_block(re, im, 0, x);
    Dft<8, -1>::run(x, V);
    double* oa = out + 2 * q * K;
    double* ob = oa + K;
    // k = 0: DCT mean (Transforms.py:28-39); a sine-type coefficient 0 does not exist and -k DCT(.) vanishes
    oa[0] = q == 0 ? 0.0 : V[0].r * (1.0 / M);   // spec_pos(0) = 0
    ob[0] = q == 0 ? 0.0 : V[0].i * (1.0 / M);
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        if (c == 3) cp_emit<M, true, QM>(q, NB * c, V[c], V[8 - c], oa, ob, true, tb);
        else cp_emit<M, false, QM>(q, NB * c, V[c], V[8 - c], oa, ob, false, tb);
    }
    cp_emit<M, false, QM>(q, M / 2, V[4], V[4], oa, ob, false, tb);
}

template <int M, int Q, int STEP>
SDDC_HD void cp_transform(int id, const double* __restrict__ re, const double* __restrict__ im, double* __restrict__ out,
                          const Tables& tb) {
    constexpr int NP = Cfg<M>::NBLK / 2;
    for (int j = id + 1; j <= NP; j += STEP) cp_unit<M, Q>(Q, j, re, im, out, tb);
    if (id == NP % STEP) cp_block0<M, Q>(Q, re, im, out, tb);
}

template <int M, int NTH>
SDDC_HD void cp_fwd(int t, const double* __restrict__ buf, double* __restrict__ out, const Tables& tb) {
    constexpr int NP = Cfg<M>::NBLK / 2, PL = Cfg<M>::PL;
    if (M >= 384) {
        using WM = WarpMap<NTH, 2>;
        const int warp = t >> 5, id = WM::first(t);
        if (WM::mine(0, warp)) cp_transform<M, 0, WM::STEP>(id, buf, buf + PL, out, tb);
        if (WM::mine(1, warp)) cp_transform<M, 1, WM::STEP>(id, buf + pair_off<M>(1), buf + pair_off<M>(1) + PL, out, tb);
        return;
    }
    for (int u = t; u < 2 * NP; u += NTH) {
        const int q = u / NP, j = u - q * NP + 1;
        cp_unit<M, 2>(q, j, buf + pair_off<M>(q), buf + pair_off<M>(q) + PL, out, tb);
    }
    const int sp = t - spec_base<NTH>(2 * NP);
    if (sp >= 0 && sp < 2) cp_block0<M, 2>(sp, buf + pair_off<M>(sp), buf + pair_off<M>(sp) + PL, out, tb);
}

// ---- kinetic energy (Main.py:71-134) with the fused packing ----------------------------------------------------------
// The 3K grid keeps only K = M/3 coefficients per field.  ke_stage copies the cosine-type row (scaled by asc = 1/r) and the
// sine-type row into shared memory, zero padded to the 2M/3 entries bc_transform reads, so that packing and the radix-8 pass
// run from registers exactly as for the nonlinear term (one shared-memory round trip less than build_ke + pass_c).
template <int M>
SDDC_HD void ke_stage(int t, const double* __restrict__ ca, const double* __restrict__ sb, double asc,
                      double* __restrict__ rows) {
    constexpr int Kc = M / 3, K2 = Cfg<M>::K;
    for (int k = t; k < K2; k += NTW) {
        rows[k] = k < Kc ? asc * ca[k] : 0.0;
        rows[K2 + k] = k < Kc ? sb[k] : 0.0;
    }
}
template <int M>
SDDC_HD void ke_pack(int t, const double* __restrict__ rows, double* __restrict__ buf, const Tables& tb) {
    bc_transform<M, 0, 1, NTW>(t, PackSrc{rows, rows + Cfg<M>::K}, buf, buf + Cfg<M>::PL, tb);
}

// ---- staged schedule of the one-state kernel (nlin_fft_staged_kernel): per-warp ownership of transforms ------------------
// Worker buffer: plane pairs 0..3 (8 PL doubles), then 3 K doubles (DS | -kT | -kS) of the staged row; the rows
// (JT, omega) / (DT, Dpsi) of the staged row sit in the planes of pair 2 / 3.
// what warp `warp` of a worker does between two product phases; shared by the kernel and tests/fft_emul.cpp
template <int M>
SDDC_HD void staged_pack(int warp, int lane, double* __restrict__ buf, const Tables& tb, const C (&tw)[Cfg<M>::RD], int stage) {
    constexpr int PL = Cfg<M>::PL, K = Cfg<M>::K;
    double* extra = buf + 8 * PL;   // DS | -kT | -kS
    double* own = buf + (4 + 2 * warp) * PL;   // staged rows of transform `warp`, later the planes of transform 2 + warp
    if (stage == 0) {
        bc_transform<M, 0, 1, 32>(lane, PackSrc{own, own + K}, buf + 2 * warp * PL, buf + (2 * warp + 1) * PL, tb);
    } else if (stage == 1) {
        if (warp == 0) bc_transform<M, 0, 1, 32>(lane, PackSrc{extra, extra + K}, own, own + PL, tb);
        else bc_transform<M, 0, 0, 32>(lane, PackSrc{nullptr, extra + 2 * K}, own, own + PL, tb);
    } else {
        pass_d_warp<M, 2, +1>(lane, buf + 2 * warp * PL, own, tw);
    }
}
template <int M>
SDDC_HD void staged_unpack(int warp, int lane, double* __restrict__ buf, double* __restrict__ out, const Tables& tb,
                           const C (&tw)[Cfg<M>::RD], int stage) {
    constexpr int PL = Cfg<M>::PL;
    double* re = buf + 2 * warp * PL;
    if (stage == 0) pass_d_warp<M, 1, -1>(lane, re, re, tw);
    else if (warp == 0) cp_transform<M, 0, 32>(lane, re, re + PL, out, tb);
    else cp_transform<M, 1, 32>(lane, re, re + PL, out, tb);
}

}  // namespace fftp
}  // namespace sddc
