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
    // chain of sine modes m = K-ch, K-ch-2, ... ; loads are issued in batches of U so that the memory latency is
    // paid once per batch, the additions stay strictly sequential (same order as the reference)
    constexpr int U = 16;
    double S = 0.0, prev = 0.0;  // prev = psi^(m+2)
    bool first = true;
    for (int m0 = K - ch; m0 >= 0; m0 -= 2 * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 - 2 * u;
            v[u] = (m >= 1 && !zero_chain) ? psi[(long long)(m - 1) * n + i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 - 2 * u;
            if (m < 0) break;
            if (!first) S += prev;
            first = false;
            prev = v[u];
            out[(long long)m * n + i] = (m >= 1) ? ((m + 1.0) * v[u] + 2.0 * S) : S;
        }
    }
}

struct PrepParams {
    const double* X;        // [B][3N] state (or perturbation)
    long long x_stride;
    const double* JJ;       // [B][K+1][n] from scan_kernel on the same X
    double* coef;           // coefficient set, tile-major [B][Khp/8][2 ks][2 par][9*n8][4] (see k_synth.cuh), or null;
                            // FFTL kernels: the seven spectral rows [B][n][7][K] of fft_core.h instead
    long long coef_stride;  // member stride of coef in doubles
    double* lin;            // linear right-hand side in solve-major order [3][K][bstride][n8+2] (k_solve.cuh), or null
    long long bstride;      // members per (field, mode) slab of lin
    const double* Ra;       // [B]
    const double* Ras;      // [B]
    const double *DrP, *D2rP, *DsqP;  // row-major operators zero-padded to [n8][n8+4]
    const double *ir2, *ir4, *r2, *dT0, *gb;  // [n]
    Geo g;
    int B;
    int nstage;             // 2: the next tile's copies overlap the current tile (n8 >= 40: one CTA per SM anyway);
                            // 1: n8 <= 32, where three single-stage CTAs per SM measured faster than two double-buffered
                            //    ones (0.128 vs 0.138 ms), and n8 = 64, where two stages do not fit
};

constexpr int PREP_TC = 32;  // sinusoid columns per CTA

__host__ __device__ inline size_t prep_smem_bytes(int n8, int nstage) {
    return sizeof(double) * (size_t)(nstage * 4 * (PREP_TC + 1) + 3 * n8) * (n8 + 4);
}

// CTA c takes the tiles (member, 32 sinusoid columns) c, c + gridDim.x, ...  For n8 >= 40 (one CTA per SM) the launch
// is persistent: the operators are loaded once per CTA and the four state tiles of the next tile are in flight
// (cp.async, two stages) while the current one goes through the MMAs and the epilogue stores (0.45 -> 0.39 ms at
// (40,512)).  For n8 <= 32 the grid has one CTA per tile, three per SM: measured faster than two double-buffered or
// three single-stage persistent CTAs (0.128 vs 0.138 / 0.142 ms at (30,256)).  2*nt8 warps: warp = (8-row radial tile mt, half of the columns).
// The radial derivatives Dr psi, D2r psi, Dsq psi, Dr T, Dr S are DMMA GEMMs
//     OUT[i, c] = sum_i' Mat[i, i'] * X[c][i']      (A = operator, B = 32 mode blocks of the state tile)
// whose accumulator fragments are written straight into (a) the tile-major coefficient arrays of the synthesis
// GEMM -- one accumulator tile is one contiguous 256-byte A-fragment block there -- and (b) the linear
// right-hand sides of Step_Python (Main.py:266-280) in the state layout.
// FFTL selects the output of the FFT formulation (k_nlin_fft.cuh): seven row-major spectral rows per radial point
// (JT, omega, DT, Dpsi, DS, -kT, -kS: one (cosine | sine) pair per transform of fft_fused.h) and the natural MMA column order, so that a quad writes 64 contiguous bytes.
template <int NT8, bool FFTL = false>
__global__ void __launch_bounds__(64 * NT8, (NT8 <= 4) ? 3 : 1) prep_kernel(PrepParams p, int ntiles) {
    extern __shared__ __align__(128) double smem[];
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K, N = g.N, LDX = n8 + 4;
    const int nkt = (K + PREP_TC - 1) / PREP_TC;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    constexpr int TQ = PREP_TC + 1;
    const int TS = 4 * TQ * LDX;     // doubles per stage: psi | T | S | JJ tiles
    const int nst = p.nstage;
    double* mDr = smem + nst * TS;   // [n8][LDX]
    double* mD2r = mDr + n8 * LDX;
    double* mDsq = mD2r + n8 * LDX;

    // all tile loads are asynchronous copies issued back to back (zero fill for masked / padded entries)
    auto issue = [&](int tile, int stage) {
        const int b = tile / nkt, c0 = (tile - b * nkt) * PREP_TC;
        double* sP = smem + stage * TS;  // psi blocks c0-1 .. c0+31   (q = 0..32), row stride LDX, padded entries zero
        double* sT = sP + TQ * LDX;      // T   blocks c0   .. c0+32
        double* sS = sT + TQ * LDX;      // S   blocks c0   .. c0+32
        double* sJ = sS + TQ * LDX;      // JJ  modes  c0   .. c0+32
        const double* Xb = p.X + (long long)b * p.x_stride;
        const double* Jb = p.JJ + (long long)b * (K + 1) * n;
        for (int idx = tid; idx < TQ * LDX; idx += nthr) {
            const int q = idx / LDX, i = idx - q * LDX;
            const int bp = c0 - 1 + q, bt = c0 + q;
            const bool in = i < n;
            const bool okp = in && bp >= 0 && bp < K && !(g.symmetric && (bp & 1) == 0);
            const bool okt = in && bt < K && !(g.symmetric && (bt & 1) == 1);
            const bool okj = in && bt <= K;
            cp_async8_zfill(&sP[idx], okp ? Xb + (long long)bp * n + i : Xb, okp);
            cp_async8_zfill(&sT[idx], okt ? Xb + (long long)N + (long long)bt * n + i : Xb, okt);
            cp_async8_zfill(&sS[idx], okt ? Xb + 2LL * N + (long long)bt * n + i : Xb, okt);
            cp_async8_zfill(&sJ[idx], okj ? Jb + (long long)bt * n + i : Jb, okj);
        }
    };
    pdl_launch_dependents();
    for (int idx = tid; idx < n8 * LDX / 2; idx += nthr) {
        cp_async16(&mDr[2 * idx], p.DrP + 2 * idx);
        cp_async16(&mD2r[2 * idx], p.D2rP + 2 * idx);
        cp_async16(&mDsq[2 * idx], p.DsqP + 2 * idx);
    }
    pdl_wait();   // the state and its suffix sums come from the previous back-substitution
    int tile = blockIdx.x, stage = 0;
    if (tile < ntiles) issue(tile, 0);
    cp_async_commit();

    const int mt = warp >> 1, nh = warp & 1;
    const bool want_lin = p.lin != nullptr;
    const int i = mt * 8 + gq;
    const double dtPr = g.dt * g.Pr;
    const int R9 = 9 * n8;
    const int LDG = n8 + 2;
    double* lb = p.lin;
    const int arow = (mt * 8 + gq) * LDX + tq;
    // MMA column slot (nl, j) -> sinusoid column of the 16-column half tile: odd columns (odd wavenumbers) in
    // natural order, even columns class-split like chunk_pos, so that the four lanes sharing an accumulator row
    // write four consecutive doubles of the tile-major coefficient arrays for both parities
    int bcol[2];
#pragma unroll
    for (int nl = 0; nl < 2; ++nl) {
        const int e = gq & 1, tt = gq >> 1;
        const int kloc = (e == 0) ? 2 * tt + nl : 4 * nl + tt;
        bcol[nl] = (FFTL ? (nh * 16 + nl * 8 + gq) : (nh * 16 + 2 * kloc + e)) * LDX + tq;
    }

    for (; tile < ntiles; tile += gridDim.x, stage = (nst == 2 ? stage ^ 1 : 0)) {
        const int nxt = tile + gridDim.x;
        if (nst == 2) {
            if (nxt < ntiles) issue(nxt, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int b = tile / nkt, c0 = (tile - b * nkt) * PREP_TC;
        const double* sP = smem + stage * TS;
        const double* sT = sP + TQ * LDX;
        const double* sS = sT + TQ * LDX;
        const double* sJ = sS + TQ * LDX;
        double dps[2][2] = {}, d2r[2][2] = {}, dsq[2][2] = {}, dT[2][2] = {}, dS[2][2] = {};
        for (int ks = 0; ks < n8 / 4; ++ks) {
            const double aDr = mDr[arow + ks * 4], aD2r = mD2r[arow + ks * 4], aDsq = mDsq[arow + ks * 4];
#pragma unroll
            for (int nl = 0; nl < 2; ++nl) {
                const int o = bcol[nl] + ks * 4;
                const double bP = sP[o], bT = sT[o], bS = sS[o];
                mma884(dps[nl][0], dps[nl][1], aDr, bP);
                mma884(d2r[nl][0], d2r[nl][1], aD2r, bP);
                mma884(dT[nl][0], dT[nl][1], aDr, bT);
                mma884(dS[nl][0], dS[nl][1], aDr, bS);
                if (want_lin) mma884(dsq[nl][0], dsq[nl][1], aDsq, sP[o + LDX]);  // psi block c (sine mode c+1)
            }
        }
        if (i < n) {
        const double ir2 = p.ir2[i], ir4 = p.ir4[i], r2 = p.r2[i], dT0 = p.dT0[i], gb = p.gb[i];
        double* cf = p.coef ? p.coef + (long long)b * p.coef_stride : nullptr;
        auto lin_off = [&](int f, int blk) { return (((long long)f * K + blk) * p.bstride + b) * LDG + i; };
        const double Ra = want_lin ? p.Ra[b] : 0.0, Ras = want_lin ? p.Ras[b] : 0.0;
#pragma unroll
        for (int nl = 0; nl < 2; ++nl) {
            if (FFTL && p.coef) {
                // columns (col, col+1) of the accumulator pair; K is even and c0 a multiple of 32, so both are < K together
                const int col = nh * 16 + nl * 8 + 2 * tq, c = c0 + col;
                if (c < K) {
                    const double j0 = sJ[col * LDX + i], j1 = sJ[(col + 1) * LDX + i];
                    double2 om, dp;
                    om.x = (c >= 1) ? d2r[nl][0] - (double)c * (ir4 * j0) : 0.0;
                    dp.x = (c >= 1) ? dps[nl][0] : 0.0;
                    om.y = d2r[nl][1] - (double)(c + 1) * (ir4 * j1);
                    dp.y = dps[nl][1];
                    double* r7 = p.coef + ((long long)b * n + i) * 7 * K + c;
                    *reinterpret_cast<double2*>(r7) = make_double2(j0, j1);                       // JT
                    *reinterpret_cast<double2*>(r7 + (long long)K) = om;                          // omega
                    *reinterpret_cast<double2*>(r7 + 2LL * K) = make_double2(dT[nl][0], dT[nl][1]);
                    *reinterpret_cast<double2*>(r7 + 3LL * K) = dp;                               // Dpsi
                    *reinterpret_cast<double2*>(r7 + 4LL * K) = make_double2(dS[nl][0], dS[nl][1]);
                    // -k T, -k S: the sine series of the theta-derivatives (Matrix_Operators.py:711-716)
                    const double m0 = -(double)c, m1 = -(double)(c + 1);
                    *reinterpret_cast<double2*>(r7 + 5LL * K) = make_double2(m0 * sT[col * LDX + i], m1 * sT[(col + 1) * LDX + i]);
                    *reinterpret_cast<double2*>(r7 + 6LL * K) = make_double2(m0 * sS[col * LDX + i], m1 * sS[(col + 1) * LDX + i]);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = FFTL ? nh * 16 + nl * 8 + 2 * tq + e
                                     : nh * 16 + 2 * ((e == 0) ? 2 * tq + nl : 4 * nl + tq) + e;   // see the slot map above
                const int c = c0 + col;
                if (c >= K) continue;
                const double jj = sJ[col * LDX + i];
                const double tv = sT[col * LDX + i], sv = sS[col * LDX + i];
                if (FFTL) {
                    // written above, two adjacent columns per 16-byte store
                } else if (cf) {
                    const int kp = chunk_pos(c >> 1, e);  // parity of c is e (c0 is a multiple of 32); position in the chunk
                    const long long o = (((long long)((kp >> 2) * 2 + e) * R9) + i) * 4 + (kp & 3);
                    const long long fs = (long long)n8 * 4;
                    double om = 0.0, dp = 0.0;
                    if (c >= 1) {
                        om = d2r[nl][e] - (double)c * (ir4 * jj);
                        dp = dps[nl][e];
                    }
                    cf[0 * fs + o] = jj;                  // JT
                    cf[1 * fs + o] = (double)c * dp;      // k Dpsi
                    cf[2 * fs + o] = (double)c * om;      // k omega
                    cf[3 * fs + o] = dT[nl][e];           // DT
                    cf[4 * fs + o] = dS[nl][e];           // DS
                    cf[5 * fs + o] = om;                  // omega
                    cf[6 * fs + o] = dp;                  // Dpsi
                    cf[7 * fs + o] = -(double)c * tv;     // -k T
                    cf[8 * fs + o] = -(double)c * sv;     // -k S
                }
                if (lb) {
                    // psi equation, block c <-> sine mode m = c+1: A2_SINE(psi) + dt Pr G(Ra T - Ra_s S)
                    const int m = c + 1;
                    double a2 = dsq[nl][e] - (double)m * (ir2 * sJ[(col + 1) * LDX + i]);
                    if (m <= K - 1) {
                        const double w = Ra * sT[(col + 1) * LDX + i] - Ras * sS[(col + 1) * LDX + i];
                        a2 += dtPr * ((-(double)m * gb) * w);
                    }
                    lb[lin_off(0, c)] = a2;
                    // T, S equations, cosine mode c: r^2 T - dt * dT0 * J_theta(psi)
                    const double pT0 = dT0 * jj;
                    lb[lin_off(1, c)] = r2 * tv - g.dt * pT0;
                    lb[lin_off(2, c)] = r2 * sv - g.dt * pT0;
                }
            }
        }
        }
        __syncthreads();   // every read of this stage is done before the copies of the tile after next refill it
        if (nst == 1) {
            if (nxt < ntiles) issue(nxt, 0);
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
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
