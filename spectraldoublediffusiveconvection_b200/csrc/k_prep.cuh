// Stage 0/1 of a member-step: theta-coupling suffix sums (scan) and the radial "prep" stage that turns a
// state X into (a) the nine spectral coefficient arrays of Derivatives (Matrix_Operators.py:630-740) in the
// row-major, parity-split layout the synthesis GEMM consumes and (b) the linear right-hand sides of
// Step_Python (Main.py:266-280).
#pragma once
#include "common.cuh"

namespace sddc {

// JJ[b][m][i], m = 0..K :  JJ[0] = S[0],  JJ[m] = (m+1) psi^(m) + 2 S[m]  (m >= 1)
// with S[m] = sum_{p = m+2, m+4, .. <= K} psi^(p)  accumulated from the highest mode downward exactly like the
// running vectors b / f_e of J_theta_RT (Matrix_Operators.py:436-472) and A2_SINE(_R2) (192-245, 475-526).
// JJ[j] (j < K) is cosine block j of J_theta_RT(psi); JJ[m] is the bracket of A2_SINE / A2_SINE_R2 for sine mode m.
__global__ void __launch_bounds__(128) scan_kernel(const double* __restrict__ X, long long x_stride,
                                                   double* __restrict__ JJ, Geo g, int B) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int n = g.n, K = g.K;
    if (gid >= (long long)B * 2 * n) return;
    const int i = (int)(gid % n);
    const int ch = (int)((gid / n) & 1);
    const int b = (int)(gid / (2 * n));
    const double* psi = X + (long long)b * x_stride;  // block m-1 holds sine mode m
    double* out = JJ + (long long)b * (K + 1) * n;
    const bool zero_chain = g.symmetric && ch == 1;   // odd sine modes are masked out (Main.py:157-164)
    int m = K - ch;
    double S = 0.0;
    double pm = zero_chain ? 0.0 : psi[(long long)(m - 1) * n + i];
    out[(long long)m * n + i] = (m + 1.0) * pm;
#pragma unroll 8
    for (m -= 2; m >= 0; m -= 2) {
        S += pm;  // += psi^(m+2)
        pm = (m >= 1 && !zero_chain) ? psi[(long long)(m - 1) * n + i] : 0.0;
        out[(long long)m * n + i] = (m >= 1) ? ((m + 1.0) * pm + 2.0 * S) : S;
    }
}

struct PrepParams {
    const double* X;        // [B][3N] state (or perturbation)
    long long x_stride;
    const double* JJ;       // [B][K+1][n] from scan_kernel on the same X
    double* coef;           // coefficient set, tile-major [B][Khp/8][2 ks][2 par][9*n8][4] (see k_synth.cuh), or null
    long long coef_stride;  // member stride of coef in doubles
    double* lin;            // [B][3N] linear right-hand side, or null
    const double* Ra;       // [B]
    const double* Ras;      // [B]
    const double *DrT, *D2rT, *DsqT;  // transposed, row-padded operators [n][n8]: M^T[i'][i]
    const double *ir2, *ir4, *r2, *dT0, *gb;  // [n]
    Geo g;
    int B;
};

constexpr int PREP_TC = 32;  // sinusoid columns per CTA

__host__ __device__ inline size_t prep_smem_bytes(int n, int n8) {
    return sizeof(double) * (size_t)(4 * (PREP_TC + 1) * n + 3 * n * n8);
}

// One CTA per (member, tile of 32 sinusoid columns).  Phase A writes the coefficient arrays (k fastest),
// phase B the linear right-hand sides (radial index fastest, the state layout).
__global__ void __launch_bounds__(256) prep_kernel(PrepParams p) {
    extern __shared__ __align__(128) double smem[];
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K, N = g.N;
    const int b = blockIdx.y, c0 = blockIdx.x * PREP_TC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TQ = PREP_TC + 1;
    double* sP = smem;              // psi blocks c0-1 .. c0+31   (q = 0..32)
    double* sT = sP + TQ * n;       // T   blocks c0   .. c0+32
    double* sS = sT + TQ * n;       // S   blocks c0   .. c0+32
    double* sJ = sS + TQ * n;       // JJ  modes  c0   .. c0+32
    double* mDr = sJ + TQ * n;      // [n][n8]
    double* mD2r = mDr + n * n8;
    double* mDsq = mD2r + n * n8;

    const double* Xb = p.X + (long long)b * p.x_stride;
    const double* Jb = p.JJ + (long long)b * (K + 1) * n;
    for (int idx = tid; idx < TQ * n; idx += 256) {
        const int q = idx / n, i = idx - q * n;
        const int bp = c0 - 1 + q, bt = c0 + q;
        double vp = 0.0, vt = 0.0, vs = 0.0, vj = 0.0;
        if (bp >= 0 && bp < K && !(g.symmetric && (bp & 1) == 0)) vp = Xb[(long long)bp * n + i];
        if (bt < K && !(g.symmetric && (bt & 1) == 1)) {
            vt = Xb[(long long)N + (long long)bt * n + i];
            vs = Xb[2LL * N + (long long)bt * n + i];
        }
        if (bt <= K) vj = Jb[(long long)bt * n + i];
        sP[idx] = vp; sT[idx] = vt; sS[idx] = vs; sJ[idx] = vj;
    }
    for (int idx = tid; idx < n * n8; idx += 256) {
        mDr[idx] = p.DrT[idx]; mD2r[idx] = p.D2rT[idx]; mDsq[idx] = p.DsqT[idx];
    }
    __syncthreads();

    // ---- phase A: coefficient arrays; warp item = (field group, quad of radial rows); lane = column ----
    if (p.coef != nullptr) {
        const int nq = n8 / 4;
        const int c = c0 + lane;
        const int par = c & 1, kp = c >> 1;
        double* cf = p.coef + (long long)b * p.coef_stride;
        const int R9 = 9 * n8;
        // element (field a, row i, parity, k') -> tile-major offset; fs = field stride inside one [row][4] plane
        const long long fs = (long long)n8 * 4;
        const long long cbase = ((long long)((kp >> 2) * 2 + par) * R9) * 4 + (kp & 3);
        for (int it = warp; it < 3 * nq; it += 8) {
            const int grp = it / nq, i0 = (it - grp * nq) * 4;
            const double* sx = (grp == 0) ? sP : (grp == 1 ? sT : sS);
            double d[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
            if (grp == 0) {
                for (int ip = 0; ip < n; ++ip) {
                    const double x = sx[lane * n + ip];
                    const double2 a0 = *reinterpret_cast<const double2*>(&mDr[ip * n8 + i0]);
                    const double2 a1 = *reinterpret_cast<const double2*>(&mDr[ip * n8 + i0 + 2]);
                    const double2 b0 = *reinterpret_cast<const double2*>(&mD2r[ip * n8 + i0]);
                    const double2 b1 = *reinterpret_cast<const double2*>(&mD2r[ip * n8 + i0 + 2]);
                    d[0] = fma(a0.x, x, d[0]); d[1] = fma(a0.y, x, d[1]);
                    d[2] = fma(a1.x, x, d[2]); d[3] = fma(a1.y, x, d[3]);
                    e[0] = fma(b0.x, x, e[0]); e[1] = fma(b0.y, x, e[1]);
                    e[2] = fma(b1.x, x, e[2]); e[3] = fma(b1.y, x, e[3]);
                }
            } else {
                for (int ip = 0; ip < n; ++ip) {
                    const double x = sx[lane * n + ip];
                    const double2 a0 = *reinterpret_cast<const double2*>(&mDr[ip * n8 + i0]);
                    const double2 a1 = *reinterpret_cast<const double2*>(&mDr[ip * n8 + i0 + 2]);
                    d[0] = fma(a0.x, x, d[0]); d[1] = fma(a0.y, x, d[1]);
                    d[2] = fma(a1.x, x, d[2]); d[3] = fma(a1.y, x, d[3]);
                }
            }
            if (c < K) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = i0 + r;
                    if (i >= n) break;
                    const long long o = cbase + (long long)i * 4;
                    if (grp == 0) {
                        const double jj = sJ[lane * n + i];
                        double om = 0.0, dps = 0.0;
                        if (c >= 1) {
                            om = e[r] - (double)c * (p.ir4[i] * jj);
                            dps = d[r];
                        }
                        cf[0 * fs + o] = jj;                 // JT
                        cf[1 * fs + o] = (double)c * dps;    // k Dpsi
                        cf[2 * fs + o] = (double)c * om;     // k omega
                        cf[5 * fs + o] = om;                 // omega
                        cf[6 * fs + o] = dps;                // Dpsi
                    } else {
                        const double x = sx[lane * n + i];
                        cf[(grp == 1 ? 3 : 4) * fs + o] = d[r];               // DT / DS
                        cf[(grp == 1 ? 7 : 8) * fs + o] = -(double)c * x;     // -k T / -k S
                    }
                }
            }
        }
    }

    // ---- phase B: linear right-hand sides, state layout (radial index fastest) ----
    if (p.lin != nullptr) {
        double* lb = p.lin + (long long)b * 3 * N;
        const double Ra = p.Ra[b], Ras = p.Ras[b];
        const double dtPr = g.dt * g.Pr;
        for (int idx = tid; idx < PREP_TC * n; idx += 256) {
            const int ql = idx / n, i = idx - ql * n;
            const int blk = c0 + ql;
            if (blk >= K) continue;
            // psi equation: A2_SINE(psi) + dt Pr G(Ra T - Ra_s S), sine mode m = blk + 1
            const int m = blk + 1;
            const double* xp = &sP[(ql + 1) * n];
            double mv = 0.0;
            for (int ip = 0; ip < n; ++ip) mv = fma(mDsq[ip * n8 + i], xp[ip], mv);
            double a2 = mv - (double)m * (p.ir2[i] * sJ[(ql + 1) * n + i]);
            if (m <= K - 1) {
                const double w = Ra * sT[(ql + 1) * n + i] - Ras * sS[(ql + 1) * n + i];
                a2 += dtPr * ((-(double)m * p.gb[i]) * w);
            }
            lb[(long long)blk * n + i] = a2;
            // T, S equations: r^2 T - dt * dT0 * J_theta(psi), cosine mode k = blk
            const double pT0 = p.dT0[i] * sJ[ql * n + i];
            lb[(long long)N + (long long)blk * n + i] = p.r2[i] * sT[ql * n + i] - g.dt * pT0;
            lb[2LL * N + (long long)blk * n + i] = p.r2[i] * sS[ql * n + i] - g.dt * pT0;
        }
    }
}

// Generic single-field linear operators of the reference API (sddc_linear_op): in/out [B][K][n].
struct LinopParams {
    const double* in;
    const double* JJ;   // scan of `in` as psi (ops 0-3)
    double* out;
    const double* matT; // [n][n8] transposed operator for A2_SINE (Dsq) / A2_SINE_R2 (D2r)
    const double* vec;  // [n] per-row factor: dT0 | ir2 | ir4 | gb | r2
    Geo g;
    int B, op;
};

__global__ void __launch_bounds__(256) linop_kernel(LinopParams p) {
    extern __shared__ __align__(128) double smem[];
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K;
    const int b = blockIdx.y, c0 = blockIdx.x * PREP_TC, tid = threadIdx.x;
    double* sP = smem;            // blocks c0 .. c0+31
    double* mM = sP + PREP_TC * n;
    const double* xb = p.in + (long long)b * g.N;
    const double* Jb = p.JJ ? p.JJ + (long long)b * (K + 1) * n : nullptr;
    double* ob = p.out + (long long)b * g.N;
    const bool need_mat = (p.op == 2 || p.op == 3);
    if (need_mat) {
        for (int idx = tid; idx < PREP_TC * n; idx += 256) {
            const int blk = c0 + idx / n;
            double v = 0.0;
            if (blk < K && !(g.symmetric && (blk & 1) == 0)) v = xb[(long long)c0 * n + idx];
            sP[idx] = v;
        }
        for (int idx = tid; idx < n * n8; idx += 256) mM[idx] = p.matT[idx];
        __syncthreads();
    }
    for (int idx = tid; idx < PREP_TC * n; idx += 256) {
        const int ql = idx / n, i = idx - ql * n, blk = c0 + ql;
        if (blk >= K) continue;
        double r;
        switch (p.op) {
            case 0: r = Jb[(long long)blk * n + i]; break;                                    // J_theta_RT
            case 1: r = p.vec[i] * Jb[(long long)blk * n + i]; break;                         // DT0_theta
            case 2:
            case 3: {                                                                         // A2_SINE(_R2)
                const int m = blk + 1;
                double mv = 0.0;
                for (int ip = 0; ip < n; ++ip) mv = fma(mM[ip * n8 + i], sP[ql * n + ip], mv);
                r = mv - (double)m * (p.vec[i] * Jb[(long long)m * n + i]);
                break;
            }
            case 4:                                                                           // kGR_RT.dot
                r = (blk + 1 < K) ? (-(double)(blk + 1) * p.vec[i]) * xb[(long long)(blk + 1) * n + i] : 0.0;
                break;
            default: r = p.vec[i] * xb[(long long)blk * n + i]; break;                        // R2.dot
        }
        ob[(long long)blk * n + i] = r;
    }
}

}  // namespace sddc
