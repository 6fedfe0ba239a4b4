// FFT formulation of the latitudinal transforms of NLIN_FX / NLIN_DFX (Matrix_Operators.py:743-898,
// Transforms.py:73-129), written as per-thread "phase" functions that compile for the device (k_nlin_fft.cuh) and
// for the host (tests/fft_emul.cpp runs the very same code thread by thread on the CPU against the test-only oracle).
//
// Shared pieces: the plane layout, the register butterflies, the radix-8 / radix-RD passes, the tables, and the
// kinetic-energy transform.  The packing / product / unpacking phases of the nonlinear term live in fft_fused.h.
//   Makhoul's reordering: for a field with DCT-III input X_k (cos type: X_k = c_k, k < K; sine type: X_k = s_{M-k},
//   using sin(k th_j) = (-1)^j cos((M-k) th_j))
//       v_n = Re sum_k X_k w_k e^{2 pi i k n / M},  w_k = e^{i pi k / 2M},  v_n = y_{2n}, v_{M-1-n} = y_{2n+1}
//   and with the Hermitian completion V_k = w_k (X_k - i X_{M-k}) / 2 (V_0 = X_0) two fields a, b share one complex
//   inverse DFT of Z_k = V^a_k + i V^b_k:  Re z_n = v^a_n, Im z_n = v^b_n.
//   inverse : length-M complex DFT as 8 x RD x 6 Cooley-Tukey passes in shared memory (M = 6 L, L = 8 RD).
//   forward : the same passes backwards with conjugated twiddles; the two packed real sequences are separated and
//             scaled afterwards (DCT: (2/M) Re[conj(w_k) V_k], k = 0 halved; DST: the same at index M - k).
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define SDDC_HD __host__ __device__ __forceinline__
#else
#define SDDC_HD inline
#endif

namespace sddc {
namespace fftp {

constexpr int NTW = 64;  // threads per worker (default; the M = 768 kernels use 128: template parameter NTH)

// Shared-memory layout of one plane (re or im) of a length-M sequence: 8-element blocks, block index
//     j = 6 d + k2   (spectral side: k = 6 (RD c + d) + k2 is element c of block j = k mod 6RD;
//                     grid side:     n1 = a + 8 b of sub-transform k2 is element a of block 6 b + k2)
// stored XOR-swizzled: element c of block j sits at 8 (j ^ bit3(j)) + (c ^ (j & 7)).  Bank = address mod 16 (doubles),
// a bijection of (j mod 16, c) for fixed c, so every access pattern of the passes is conflict free per half-warp:
//   16 consecutive blocks, same element (build, radix-8 pass, post)        -> 16 distinct residues of j
//   blocks (2i, 2i+1), elements 0..7 (radix-RD pass)                          -> bit0(j) flips the upper bank half
//   elements 0..15 of blocks j, j+6 (radix-6 pass; 2-way conflict only when j mod 8 < 2)
template <int M_>
struct Cfg {
    static constexpr int M = M_;
    static constexpr int K = 2 * M_ / 3;
    static constexpr int L = M_ / 6;
    static constexpr int RD = L / 8;
    static constexpr int NBLK = M_ / 8;  // = 6 RD
    static constexpr int PL = M_;        // doubles per plane
    // RD = 32 (M = 1536) exists for the kinetic-energy transform only: its middle pass runs in two steps (pass_d32_*)
    static_assert(M_ % 48 == 0 && (RD == 4 || RD == 8 || RD == 16 || RD == 32), "supported grids: M = 192, 384, 768 (1536: KE)");
};
// Position of sinusoid index k inside a [K] row of the analysed products: even indices first, then the odd ones.  A
// back-substitution chain walks one parity, so the coefficients it needs two chain steps apart are 16 contiguous bytes
// (k_solve_hot.cuh gathers them with one cp.async).
SDDC_HD constexpr int spec_pos(int k, int K) { return (k & 1) * (K >> 1) + (k >> 1); }
SDDC_HD int at(int j, int c) { return ((j ^ ((j >> 3) & 1)) << 3) | (c ^ (j & 7)); }
// offset of the (re, im) plane pair of transform q inside a worker's buffer.  On the small grid (M = 192), where the lanes
// of a half-warp straddle two transforms (blocks .., 11, 12 of pair q | blocks 1, 2, .. of pair q + 1), odd pairs are
// shifted by half a bank period; for M >= 384 every warp round stays inside one transform.
template <int M>
SDDC_HD constexpr int pair_off(int q) { return q * (2 * Cfg<M>::PL + (M < 384 ? 8 : 0)); }   // 2 PL is a multiple of 16
template <int M>
SDDC_HD constexpr int pairs_doubles(int nq) { return nq * (2 * Cfg<M>::PL + (M < 384 ? 8 : 0)); }

// table sizes (doubles): wk cos | wk sin | t6 cos | t6 sin | tL cos | tL sin
template <int M> SDDC_HD constexpr int tab_wk_doubles() { return M / 2 + 1; }
template <int M> SDDC_HD constexpr int tab_t6_doubles() { return 5 * (M / 6); }
template <int M> SDDC_HD constexpr int tab_tL_doubles() { return 9 * (M / 48); }
template <int M> SDDC_HD constexpr int tab_doubles() { return 2 * (tab_wk_doubles<M>() + tab_t6_doubles<M>() + tab_tL_doubles<M>()); }

struct Tables {
    const double *wkc, *wks;  // [M/2+1] : cos, sin (pi k / 2M) / 2
    const double *t6c, *t6s;  // [5][L]  : cos, sin (2 pi k2 n1 / M), k2 = 1..5 (row k2 - 1)
    const double *tLc, *tLs;  // [RD][9] : cos, sin (2 pi d a / L), a < 8 (row stride 9: conflict-free column reads)
};
template <int M>
SDDC_HD Tables make_tables(const double* base) {
    Tables tb;
    tb.wkc = base;
    tb.wks = tb.wkc + tab_wk_doubles<M>();
    tb.t6c = tb.wks + tab_wk_doubles<M>();
    tb.t6s = tb.t6c + tab_t6_doubles<M>();
    tb.tLc = tb.t6s + tab_t6_doubles<M>();
    tb.tLs = tb.tLc + tab_tL_doubles<M>();
    return tb;
}

struct C {
    double r, i;
};
SDDC_HD C operator+(C a, C b) { return C{a.r + b.r, a.i + b.i}; }
SDDC_HD C operator-(C a, C b) { return C{a.r - b.r, a.i - b.i}; }
SDDC_HD C cmul(C a, double c, double s) { return C{a.r * c - a.i * s, a.r * s + a.i * c}; }   // a * (c + i s)
SDDC_HD C cmulc(C a, double c, double s) { return C{a.r * c + a.i * s, a.i * c - a.r * s}; }  // a * (c - i s)
template <int SIGN>
SDDC_HD C mul_i(C a) {  // a * (SIGN i)
    return SIGN > 0 ? C{-a.i, a.r} : C{a.i, -a.r};
}

// cos / sin of 2 pi k / 16
SDDC_HD constexpr double cos16(int k) {
    return k == 0 ? 1.0
         : k == 1 ? 0.92387953251128673848
         : k == 2 ? 0.70710678118654752440
         : k == 3 ? 0.38268343236508977173
         : k == 4 ? 0.0
         : k == 5 ? -0.38268343236508977173
         : k == 6 ? -0.70710678118654752440
         : k == 7 ? -0.92387953251128673848
                  : -1.0;
}
SDDC_HD constexpr double sin16(int k) { return k <= 4 ? cos16(4 - k) : cos16(k - 4); }

// O * e^{SIGN 2 pi i k / R}
template <int R, int K_, int SIGN>
SDDC_HD C twmul(C o) {
    if (K_ == 0) return o;
    if (4 * K_ == R) return mul_i<SIGN>(o);
    constexpr int k16 = K_ * (16 / R);
    constexpr double c = cos16(k16), s = SIGN * sin16(k16);
    if (8 * K_ == R || 8 * K_ == 3 * R) {
        // |c| = |s| = 1/sqrt2
        constexpr double h = 0.70710678118654752440;
        const C m = (8 * K_ == R) ? (o + mul_i<SIGN>(o)) : (mul_i<SIGN>(o) - o);
        return C{m.r * h, m.i * h};
    }
    return cmul(o, c, s);
}

template <int R, int SIGN>
struct Dft;
template <int SIGN>
struct Dft<2, SIGN> {
    static SDDC_HD void run(const C (&in)[2], C (&out)[2]) {
        out[0] = in[0] + in[1];
        out[1] = in[0] - in[1];
    }
};
template <int R, int SIGN, int K_>
struct Combine {
    static SDDC_HD void run(const C (&E)[R / 2], const C (&O)[R / 2], C (&out)[R]) {
        const C t = twmul<R, K_, SIGN>(O[K_]);
        out[K_] = E[K_] + t;
        out[K_ + R / 2] = E[K_] - t;
        Combine<R, SIGN, K_ + 1>::run(E, O, out);
    }
};
template <int R, int SIGN>
struct Combine<R, SIGN, R / 2> {
    static SDDC_HD void run(const C (&)[R / 2], const C (&)[R / 2], C (&)[R]) {}
};
// radix-2 decimation in time, natural order in and out, everything in registers
template <int R, int SIGN>
struct Dft {
    static SDDC_HD void run(const C (&in)[R], C (&out)[R]) {
        C e[R / 2], o[R / 2], E[R / 2], O[R / 2];
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            e[j] = in[2 * j];
            o[j] = in[2 * j + 1];
        }
        Dft<R / 2, SIGN>::run(e, E);
        Dft<R / 2, SIGN>::run(o, O);
        Combine<R, SIGN, 0>::run(E, O, out);
    }
};

template <int SIGN>
SDDC_HD void dft3(C x0, C x1, C x2, C& X0, C& X1, C& X2) {
    constexpr double h = 0.86602540378443864676;
    const C t1 = x1 + x2;
    X0 = x0 + t1;
    const C t2 = C{x0.r - 0.5 * t1.r, x0.i - 0.5 * t1.i};
    const C d = x1 - x2;
    const C t3 = mul_i<SIGN>(C{h * d.r, h * d.i});
    X1 = t2 + t3;
    X2 = t2 - t3;
}
// length-6 DFT by the prime-factor map n = (3 n1 + 2 n2) mod 6, k = (3 k1 + 4 k2) mod 6 (no twiddles)
template <int SIGN>
SDDC_HD void dft6(const C (&x)[6], C (&X)[6]) {
    C A0, A1, A2, B0, B1, B2;
    dft3<SIGN>(x[0], x[2], x[4], A0, A1, A2);
    dft3<SIGN>(x[3], x[5], x[1], B0, B1, B2);
    X[0] = A0 + B0;
    X[3] = A0 - B0;
    X[4] = A1 + B1;
    X[1] = A1 - B1;
    X[2] = A2 + B2;
    X[5] = A2 - B2;
}

// storage offset of spectral index k (input order of the inverse, output order of the forward transform)
template <int M>
SDDC_HD int kpos(int k) {
    constexpr int NBLK = Cfg<M>::NBLK;
    return at(k % NBLK, k / NBLK);
}

// ---- radix-8 pass over c (the eight elements of a block) ---------------------------------------------------------------
// inverse (SIGN = +1): DFT over c -> a;  forward (SIGN = -1): DFT over a -> c.  The twiddle e^{+-2 pi i d a / L} between
// the two passes of the length-L transform is applied by pass_d, where it depends on the thread only.
template <int M, int NF, int SIGN, int NTH = NTW>
SDDC_HD void pass_c(int t, double* __restrict__ buf) {
    constexpr int NBLK = Cfg<M>::NBLK, PL = Cfg<M>::PL;
    for (int u = t; u < NF * NBLK; u += NTH) {
        const int q = u / NBLK, j = u - q * NBLK;
        const int sw = j & 7;
        double* re = buf + (2 * q) * PL + ((j ^ ((j >> 3) & 1)) << 3);
        double* im = re + PL;
        C x[8], y[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = C{re[c ^ sw], im[c ^ sw]};
        Dft<8, SIGN>::run(x, y);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            re[c ^ sw] = y[c].r;
            im[c ^ sw] = y[c].i;
        }
    }
}

// twiddles e^{2 pi i d a / L}, d < RD, of thread t: pass_d always works on element a = t & 7 (48 and the worker size
// are multiples of 8)
template <int M>
SDDC_HD void load_tw(int t, const Tables& tb, C (&tw)[Cfg<M>::RD]) {
#pragma unroll
    for (int d = 0; d < Cfg<M>::RD; ++d) tw[d] = C{tb.tLc[d * 9 + (t & 7)], tb.tLs[d * 9 + (t & 7)]};
}

// ---- radix-RD pass over d (same element of the blocks 6 d + k2) ------------------------------------------------------
// inverse: twiddle e^{+2 pi i d a / L}, then DFT over d -> b;  forward: DFT over b -> d, then twiddle e^{-2 pi i d a / L}
template <int M, int NF, int SIGN, int NTH = NTW>
SDDC_HD void pass_d(int t, double* __restrict__ buf, const C (&tw)[Cfg<M>::RD]) {
    constexpr int RD = Cfg<M>::RD, PL = Cfg<M>::PL;
    for (int u = t; u < NF * 48; u += NTH) {
        const int q = u / 48, rem = u - q * 48, k2 = rem >> 3, a = rem & 7;
        double* re = buf + pair_off<M>(q);
        double* im = re + PL;
        int o[RD];
#pragma unroll
        for (int d = 0; d < RD; ++d) o[d] = at(6 * d + k2, a);
        C x[RD], y[RD];
#pragma unroll
        for (int d = 0; d < RD; ++d) x[d] = C{re[o[d]], im[o[d]]};
        if (SIGN > 0) {
#pragma unroll
            for (int d = 1; d < RD; ++d) x[d] = cmul(x[d], tw[d].r, tw[d].i);
        }
        Dft<RD, SIGN>::run(x, y);
        if (SIGN < 0) {
#pragma unroll
            for (int d = 1; d < RD; ++d) y[d] = cmulc(y[d], tw[d].r, tw[d].i);
        }
#pragma unroll
        for (int d = 0; d < RD; ++d) {
            re[o[d]] = y[d].r;
            im[o[d]] = y[d].i;
        }
    }
}

// The same pass run by ONE warp on the transforms it owns (nlin_fft_staged_kernel): NTR plane pairs at `pa` and `pb`
// (NTR = 1: pa only), 48 units each.
template <int M, int NTR, int SIGN>
SDDC_HD void pass_d_warp(int lane, double* __restrict__ pa, double* __restrict__ pb, const C (&tw)[Cfg<M>::RD]) {
    constexpr int RD = Cfg<M>::RD, PL = Cfg<M>::PL;
    for (int u = lane; u < NTR * 48; u += 32) {
        const int sel = u / 48, rem = u - sel * 48, k2 = rem >> 3, a = rem & 7;
        double* re = sel ? pb : pa;
        double* im = re + PL;
        int o[RD];
#pragma unroll
        for (int d = 0; d < RD; ++d) o[d] = at(6 * d + k2, a);
        C x[RD], y[RD];
#pragma unroll
        for (int d = 0; d < RD; ++d) x[d] = C{re[o[d]], im[o[d]]};
        if (SIGN > 0) {
#pragma unroll
            for (int d = 1; d < RD; ++d) x[d] = cmul(x[d], tw[d].r, tw[d].i);
        }
        Dft<RD, SIGN>::run(x, y);
        if (SIGN < 0) {
#pragma unroll
            for (int d = 1; d < RD; ++d) y[d] = cmulc(y[d], tw[d].r, tw[d].i);
        }
#pragma unroll
        for (int d = 0; d < RD; ++d) {
            re[o[d]] = y[d].r;
            im[o[d]] = y[d].i;
        }
    }
}

// ---- radix-32 middle pass (M = 1536, inverse only), in place in two steps ----------------------------------------------
// 32 complex values do not fit into registers next to their twiddles, so the DFT over d = d1 + 4 d2 is split as
//   y[8 b1 + b2] = sum_d1 e^{2 pi i d1 b1 / 4} [ e^{2 pi i d1 b2 / 32} sum_d2 X[d1 + 4 d2] e^{2 pi i d2 b2 / 8} ],   X[d] = x[d] e^{2 pi i d a / L}
// step a: per (k2, a, d1) an 8-point DFT over d2, written back to the positions it read (d1 + 4 b2);
// step b: per (k2, a, b2) a 4-point DFT over d1 on the positions 4 b2 + d1, in place: position 4 b2 + b1 then holds
// y[8 b1 + b2], i.e. the pass leaves its output in the order dperm() and the last pass reads it through that map.
template <int M>
SDDC_HD constexpr int dperm(int b) { return Cfg<M>::RD == 32 ? (((b & 7) << 2) | (b >> 3)) : b; }

SDDC_HD void root32(int k, double& c, double& s) {   // e^{2 pi i k / 32}, 0 <= k < 32
    constexpr double q[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                             0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                             0.19509032201612826785, 0.0};
    auto cosq = [&](int m) {   // cos(2 pi m / 32) for any m >= 0
        m &= 31;
        if (m > 16) m = 32 - m;
        return m <= 8 ? q[m] : -q[16 - m];
    };
    c = cosq(k);
    s = cosq(k + 24);          // sin x = cos(x - pi/2) = cos(x + 3 pi/2)
}

template <int M>
SDDC_HD void pass_d32_a(int t, double* __restrict__ buf, const Tables& tb) {
    constexpr int PL = Cfg<M>::PL;
    static_assert(Cfg<M>::RD == 32, "two-step middle pass");
    double* re = buf;
    double* im = buf + PL;
    for (int u = t; u < 4 * 48; u += NTW) {
        const int d1 = u / 48, rem = u - d1 * 48, k2 = rem >> 3, a = rem & 7;
        int o[8];
        C x[8], y[8];
#pragma unroll
        for (int d2 = 0; d2 < 8; ++d2) {
            const int d = d1 + 4 * d2;
            o[d2] = at(6 * d + k2, a);
            x[d2] = cmul(C{re[o[d2]], im[o[d2]]}, tb.tLc[d * 9 + a], tb.tLs[d * 9 + a]);
        }
        Dft<8, +1>::run(x, y);
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) {
            double c, s;
            root32(d1 * b2, c, s);
            const C v = cmul(y[b2], c, s);
            re[o[b2]] = v.r;
            im[o[b2]] = v.i;
        }
    }
}
template <int M>
SDDC_HD void pass_d32_b(int t, double* __restrict__ buf) {
    constexpr int PL = Cfg<M>::PL;
    double* re = buf;
    double* im = buf + PL;
    for (int u = t; u < 8 * 48; u += NTW) {
        const int b2 = u / 48, rem = u - b2 * 48, k2 = rem >> 3, a = rem & 7;
        int o[4];
        C x[4], y[4];
#pragma unroll
        for (int d1 = 0; d1 < 4; ++d1) {
            o[d1] = at(6 * (4 * b2 + d1) + k2, a);
            x[d1] = C{re[o[d1]], im[o[d1]]};
        }
        Dft<4, +1>::run(x, y);
#pragma unroll
        for (int b1 = 0; b1 < 4; ++b1) {
            re[o[b1]] = y[b1].r;
            im[o[b1]] = y[b1].i;
        }
    }
}

// last inverse pass of one transform at column n1: six grid values z[n2] <-> grid point n = n1 + L n2
template <int M>
SDDC_HD void inv6(const double* __restrict__ re, const double* __restrict__ im, const int (&pos)[6], int n1,
                  const Tables& tb, C (&z)[6]) {
    constexpr int L = Cfg<M>::L;
    C x[6];
#pragma unroll
    for (int k2 = 0; k2 < 6; ++k2) {
        x[k2] = C{re[pos[k2]], im[pos[k2]]};
        if (k2 > 0) x[k2] = cmul(x[k2], tb.t6c[(k2 - 1) * L + n1], tb.t6s[(k2 - 1) * L + n1]);
    }
    dft6<+1>(x, z);
}

// ---- kinetic energy (Main.py:71-134): one complex transform per radial row on the M = 3K grid -----------------------
// (packing + radix-8 pass: fftp::ke_stage / ke_pack in fft_fused.h)
// last inverse pass + weighted sum of squares: returns this thread's share of
//   sum_n Wn[n] (Jpsi(n)^2 + Dpsi(n)^2),  Wn = theta trapezoid weight * sin(theta) in the transform's point order
template <int M>
SDDC_HD double ke6(int t, const double* __restrict__ buf, const Tables& tb, const double* __restrict__ Wn) {
    constexpr int L = Cfg<M>::L, PL = Cfg<M>::PL;
    double acc = 0.0;
    for (int n1 = t; n1 < L; n1 += NTW) {
        int pos[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pos[m] = at(6 * dperm<M>(n1 >> 3) + m, n1 & 7);
        C z[6];
        inv6<M>(buf, buf + PL, pos, n1, tb, z);
#pragma unroll
        for (int m = 0; m < 6; ++m) acc += Wn[n1 + L * m] * (z[m].r * z[m].r + z[m].i * z[m].i);
    }
    return acc;
}

// host: theta weights of Kinetic_Energy in the point order of the transform (n <-> grid index 2n, or 2(M-1-n)+1)
template <int M>
inline void fill_ke_weights(double* Wn) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int n = 0; n < M; ++n) {
        const int j = n < M / 2 ? 2 * n : 2 * (M - 1 - n) + 1;
        const long double th = pi * (2.0L * j + 1.0L) / (2.0L * M), dth = pi / M;
        Wn[n] = (double)(((j == 0 || j == M - 1) ? 0.5L * dth : dth) * sinl(th));   // np.trapz over the midpoint nodes only
    }
}

// host: fill the tables (tab_doubles<M>() doubles, layout of make_tables), long-double arguments
template <int M>
inline void fill_tables(double* out) {
    constexpr int L = Cfg<M>::L, RD = Cfg<M>::RD;
    const long double pi = 3.14159265358979323846264338327950288L;
    double* wkc = out;
    double* wks = wkc + tab_wk_doubles<M>();
    double* t6c = wks + tab_wk_doubles<M>();
    double* t6s = t6c + tab_t6_doubles<M>();
    double* tLc = t6s + tab_t6_doubles<M>();
    double* tLs = tLc + tab_tL_doubles<M>();
    for (int k = 0; k <= M / 2; ++k) {
        const long double x = pi * k / (2.0L * M);
        wkc[k] = (double)(0.5L * cosl(x));
        wks[k] = (double)(0.5L * sinl(x));
    }
    for (int k2 = 1; k2 < 6; ++k2)
        for (int n1 = 0; n1 < L; ++n1) {
            const long double x = 2.0L * pi * ((k2 * n1) % M) / M;
            t6c[(k2 - 1) * L + n1] = (double)cosl(x);
            t6s[(k2 - 1) * L + n1] = (double)sinl(x);
        }
    for (int d = 0; d < RD; ++d)
        for (int a = 0; a < 9; ++a) {
            const long double x = 2.0L * pi * ((d * a) % L) / L;
            tLc[d * 9 + a] = a < 8 ? (double)cosl(x) : 0.0;
            tLs[d * 9 + a] = a < 8 ? (double)sinl(x) : 0.0;
        }
}

}  // namespace fftp
}  // namespace sddc
