// Stage 2 of the dense formulation, hot-path variant for 24 < n <= 32: persistent, warp-specialised synthesis kernel
// with the second mirror level ("quarter-wave" split).  Schedule: one persistent CTA per SM walks over (member, column
// tile) work items; 12 MMA warps (3 per SM sub-partition, so the four tensor pipes are evenly loaded) run the DMMA main
// loop, a producer warp streams the operand stages with TMA bulk copies through an mbarrier ring that runs ahead across
// tile boundaries, and 8 epilogue warps turn the accumulators of the *previous* tile into folded grid products while
// the MMA warps already work on the next tile.  Other shapes use synth_kernel (k_synth.cuh).
//
// Mirror pairs j' of the grid come in orbits (L = j'', R = M/2-1-j''), j'' < M/4, with theta_R = pi/2 - theta_L, hence
//     cos(2k' theta_R) = (-1)^k' cos(2k' theta_L),      sin(2k' theta_R) = (-1)^(k'+1) sin(2k' theta_L):
// the even-wavenumber half E of every synthesis splits once more into the classes k' even (EE) and k' odd (EO),
//     cosine rows: E(L) = EE + EO, E(R) = EE - EO;        sine rows: E(L) = EE + EO, E(R) = EO - EE,
// each a contraction over K/4 wavenumbers at the L angles only.  The odd-wavenumber half O is contracted at L and
// at R as before.  Per orbit and row that is 2*K/2 + 2*K/4 multiply-adds instead of 4*K/2: 25 % less tensor work;
// the analysis GEMM gets the same saving from the doubly folded products written here (k_analysis.cuh, quarter mode).
//
// One tile = 8 orbits.  Each of the 12 MMA warps owns 3 of the 36 row tiles for BOTH parities: per 8-wavenumber
// stage 12 DMMAs for O (2 column tiles x 2 k-steps) and 6 for EE/EO (k-step 0 holds the k' even coefficients,
// k-step 1 the k' odd ones: chunk_pos in common.cuh) -- every warp does identical work.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "k_synth.cuh"

namespace sddc {

// pipeline granularity: MMA k-steps (4 wavenumbers each) per stage, and ring depth
constexpr int SWS_KS = 2, SWS_STAGES = 3;
constexpr int SWS_NT = 2, SWS_W = 16, SWS_NEW = 8;  // MMA + 1 producer + epilogue warps
constexpr int SWS_NMMA = 12, SWS_NTHR = 32 * (SWS_NMMA + 1 + SWS_NEW);

// MODE: SWS_FX   products of the fields of one state (NLIN_FX);
//       SWS_GRID no products: the nine grid fields of the state are stored to p.gridc (base state of a Newton /
//                GMRES solve: computed once, reused by every Jacobian-vector product);
//       SWS_JVPC the MMA operand is the perturbation dv, the base-state grid fields are read back from p.gridc and
//                the bilinear products of NLIN_DFX (Matrix_Operators.py:884-887) are formed -- a JVP then costs one
//                synthesis instead of two.
enum { SWS_FX = 0, SWS_GRID = 1, SWS_JVPC = 2 };

constexpr int SWQ_TPW = 3;  // row tiles per MMA warp (36 / 12)

__host__ __device__ inline size_t synth_wsq_smem_doubles(int n8) {
    const size_t rs = 9 * (size_t)n8;
    const size_t stage = (size_t)SWS_KS * 2 * rs * 4 + (size_t)SWS_KS * 4 * SWS_W * 4;
    return SWS_STAGES * stage + 2 * rs * SWS_W + 2 * (size_t)4 * n8 * 8 + (size_t)n8 * (n8 + 4);
}

template <int NT8, int MODE>
__global__ void __launch_bounds__(SWS_NTHR, 1) synth_wsq_kernel(SynthParams p, int ntiles_j, int nwork) {
    constexpr int NF = 9, RS = NF * NT8 * 8, W = SWS_W, KS = 2, TPW = SWQ_TPW;
    constexpr int A_SET = KS * 2 * RS * 4, B_ST = KS * 4 * W * 4, STAGE = A_SET + B_ST;
    constexpr int NS = 3, NMMA = 12, NEW = SWS_NEW, NTHR_E = 32 * NEW;
    constexpr int n8 = NT8 * 8, ROWS3 = 3 * n8, LDE = 8;
    static_assert(NF * NT8 == NMMA * TPW, "laid out for 36 row tiles");
    static_assert(SWS_NMMA == 12 && SWS_KS == 2 && SWS_STAGES == 3, "launch geometry");
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_eo_full, bar_eo_free;
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int nchunk = g.Khp / 8, n = g.n, Mq = g.Mh / 2;
    constexpr int LDD = n8 + 4;
    double* sEO = smem + (size_t)NS * STAGE;      // [4 classes: O_L, O_R, EE, EO][RS][8]
    double* sA1 = sEO + (size_t)4 * RS * LDE;     // [4 points][n8][8]  (rows >= n stay zero)
    double* sQ = sA1 + (size_t)4 * n8 * 8;        // [4 points][n8][8]
    double* sDr = sQ + (size_t)4 * n8 * 8;        // [n8][LDD] zero padded

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NMMA); }
        mbar_init(&bar_eo_full, NMMA);
        mbar_init(&bar_eo_free, NEW);
        mbar_fence_init();
    }
    for (int idx = tid; idx < n8 * LDD; idx += SWS_NTHR) {
        const int r = idx / LDD, c = idx - r * LDD;
        sDr[idx] = (r < n && c < n) ? p.Dr[r * n + c] : 0.0;
    }
    for (int idx = tid; idx < 4 * n8 * 8; idx += SWS_NTHR) sA1[idx] = 0.0;
    __syncthreads();

    if (warp == NMMA) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int st = 0, ph = 0;
            long long it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
                const int b = w / ntiles_j, jt = w - b * ntiles_j;
                const double* gA = p.coef0 + (long long)b * p.coef_stride;
                const double* gB = p.tab + (long long)jt * nchunk * B_ST;
                for (int c = 0; c < nchunk; ++c, ++it) {
                    if (it >= NS) mbar_wait(&bar_empty[st], ph ^ 1);
                    double* sA = smem + (size_t)st * STAGE;
                    mbar_expect_tx(&bar_full[st], (unsigned)(STAGE * sizeof(double)));
                    bulk_g2s(sA, gA + (long long)c * A_SET, A_SET * sizeof(double), &bar_full[st]);
                    bulk_g2s(sA + A_SET, gB + (long long)c * B_ST, B_ST * sizeof(double), &bar_full[st]);
                    if (++st == NS) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp < NMMA) {
        // ---------------- MMA warps: 3 row tiles, both parities ----------------
        const int tile0 = warp * TPW;
        const int split = min(TPW, max(0, 5 * NT8 - tile0));   // tiles [0, split) cosine rows, [split, TPW) sine rows
        const int a0_off = (tile0 * 8 + gq) * 4 + tq;                    // parity-0 rows (+ ks*2*RS*4 + mt*32)
        const int a1_off = (RS + tile0 * 8 + gq) * 4 + tq;               // parity-1 rows
        const int b_off = A_SET + gq * 4 + tq;                           // + ks*4*W*4 + (ty*2+par)*W*4 + nt*32
        int st = 0, ph = 0, tcount = 0;
        double accO[TPW][2][2], accE[TPW][2][2];   // O at the L / R column tile; EE / EO at the L column tile
        auto stage_mma = [&](const double* sS, auto split_tag) {
            constexpr int SPLIT = decltype(split_tag)::value;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double* sB = sS + b_off + ks * 4 * W * 4;
                double b1c[2], b1s[2], b0c = 0.0, b0s = 0.0, a0[TPW], a1[TPW];
                if (SPLIT > 0) { b1c[0] = sB[(0 * 2 + 1) * W * 4]; b1c[1] = sB[(0 * 2 + 1) * W * 4 + 32]; b0c = sB[0]; }
                if (SPLIT < TPW) { b1s[0] = sB[(1 * 2 + 1) * W * 4]; b1s[1] = sB[(1 * 2 + 1) * W * 4 + 32]; b0s = sB[(1 * 2 + 0) * W * 4]; }
#pragma unroll
                for (int mt = 0; mt < TPW; ++mt) {
                    a0[mt] = sS[a0_off + ks * 2 * RS * 4 + mt * 32];
                    a1[mt] = sS[a1_off + ks * 2 * RS * 4 + mt * 32];
                }
#pragma unroll
                for (int mt = 0; mt < TPW; ++mt) {
                    mma884(accO[mt][0][0], accO[mt][0][1], a1[mt], mt < SPLIT ? b1c[0] : b1s[0]);
                    mma884(accO[mt][1][0], accO[mt][1][1], a1[mt], mt < SPLIT ? b1c[1] : b1s[1]);
                    mma884(accE[mt][ks][0], accE[mt][ks][1], a0[mt], mt < SPLIT ? b0c : b0s);
                }
            }
        };
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tcount) {
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt)
#pragma unroll
                for (int x = 0; x < 2; ++x) accO[mt][x][0] = accO[mt][x][1] = accE[mt][x][0] = accE[mt][x][1] = 0.0;
            for (int c = 0; c < nchunk; ++c) {
                mbar_wait(&bar_full[st], ph);
                const double* sS = smem + (size_t)st * STAGE;
                if (split == TPW) stage_mma(sS, std::integral_constant<int, TPW>{});
                else if (split == 0) stage_mma(sS, std::integral_constant<int, 0>{});
                else stage_mma(sS, std::integral_constant<int, (5 * NT8) % TPW>{});
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[st]);
                if (++st == NS) { st = 0; ph ^= 1; }
            }
            if (tcount > 0) mbar_wait(&bar_eo_free, (tcount - 1) & 1);
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt) {
                const size_t row = (size_t)(tile0 + mt) * 8 + gq;
                *reinterpret_cast<double2*>(&sEO[(0 * RS + row) * LDE + 2 * tq]) = make_double2(accO[mt][0][0], accO[mt][0][1]);
                *reinterpret_cast<double2*>(&sEO[(1 * RS + row) * LDE + 2 * tq]) = make_double2(accO[mt][1][0], accO[mt][1][1]);
                *reinterpret_cast<double2*>(&sEO[(2 * RS + row) * LDE + 2 * tq]) = make_double2(accE[mt][0][0], accE[mt][0][1]);
                *reinterpret_cast<double2*>(&sEO[(3 * RS + row) * LDE + 2 * tq]) = make_double2(accE[mt][1][0], accE[mt][1][1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_eo_full);
        }
    } else {
        // ---------------- epilogue warps: one (radial row, orbit) item = 4 grid points per thread ----------------
        const int et = tid - 32 * (NMMA + 1);
        const int nitem = n * 8;
        constexpr int PTS = (n8 * 8 + NTHR_E - 1) / NTHR_E;
        int tcount = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tcount) {
            const int b = w / ntiles_j, jt = w - b * ntiles_j;
            double* prd = p.prd + (long long)b * 2 * g.Mhp * ROWS3;
            const long long pps = (long long)g.Mhp * ROWS3;   // parity stride
            auto prd_off = [&](int f, int i, int pos) { return ((long long)(pos >> 2) * ROWS3 + f * n8 + i) * 4 + (pos & 3); };
            double* gc = (MODE != SWS_FX) ? p.gridc + (long long)b * 9 * 2 * n8 * g.Mhp : nullptr;
            const long long gms = (long long)n8 * g.Mhp;
            mbar_wait(&bar_eo_full, tcount & 1);
#pragma unroll
            for (int s = 0; s < PTS; ++s) {
                const int it = et + s * NTHR_E;
                if (it < nitem) {
                    const int i = it >> 3, c = it & 7;
                    const int jq = jt * 8 + c;                      // orbit index: L = jq, R = Mh-1-jq
                    double nt[4], ns[4];   // [L, mirror(L), R, mirror(R)]
                    // the two mirror pairs of the orbit are processed one after the other to bound register use
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        double f[9][2];    // [field][point of the pair, its mirror point]
#pragma unroll
                        for (int a = 0; a < 9; ++a) {
                            const size_t r = ((size_t)a * n8 + i) * LDE + c;
                            const double o = sEO[(size_t)pr * RS * LDE + r];          // O at L (pr 0) or R (pr 1)
                            const double ee = sEO[(size_t)2 * RS * LDE + r], eo = sEO[(size_t)3 * RS * LDE + r];
                            if (a < 5) {
                                const double e = pr == 0 ? ee + eo : ee - eo;
                                f[a][0] = e + o; f[a][1] = e - o;
                            } else {
                                const double e = pr == 0 ? ee + eo : eo - ee;
                                f[a][0] = o + e; f[a][1] = o - e;
                            }
                        }
                        const long long gp = (long long)i * g.Mhp + (pr == 0 ? jq : g.Mh - 1 - jq);
                        if (MODE == SWS_GRID) {
#pragma unroll
                            for (int a = 0; a < 9; ++a) {
                                gc[(a * 2 + 0) * gms + gp] = f[a][0];
                                gc[(a * 2 + 1) * gms + gp] = f[a][1];
                            }
                            continue;
                        }
                        double a1[2], qv[2];
                        if (MODE == SWS_FX) {
#pragma unroll
                            for (int x = 0; x < 2; ++x) {
                                a1[x] = f[0][x] * f[5][x];
                                qv[x] = f[1][x] * f[5][x] + f[6][x] * f[2][x];
                                nt[2 * pr + x] = f[0][x] * f[3][x] - f[6][x] * f[7][x];
                                ns[2 * pr + x] = f[0][x] * f[4][x] - f[6][x] * f[8][x];
                            }
                        } else {
                            double h[9][2];   // perturbation fields (just synthesised); f <- cached base-state fields
#pragma unroll
                            for (int a = 0; a < 9; ++a) {
                                h[a][0] = f[a][0]; h[a][1] = f[a][1];
                                f[a][0] = gc[(a * 2 + 0) * gms + gp];
                                f[a][1] = gc[(a * 2 + 1) * gms + gp];
                            }
#pragma unroll
                            for (int x = 0; x < 2; ++x) {
                                a1[x] = f[0][x] * h[5][x] + h[0][x] * f[5][x];
                                qv[x] = (f[1][x] * h[5][x] + f[6][x] * h[2][x]) + (h[1][x] * f[5][x] + h[6][x] * f[2][x]);
                                nt[2 * pr + x] = (h[0][x] * f[3][x] - h[6][x] * f[7][x]) + (f[0][x] * h[3][x] - f[6][x] * h[7][x]);
                                ns[2 * pr + x] = (h[0][x] * f[4][x] - h[6][x] * f[8][x]) + (f[0][x] * h[4][x] - f[6][x] * h[8][x]);
                            }
                        }
                        sA1[((size_t)(2 * pr + 0) * n8 + i) * 8 + c] = a1[0];
                        sA1[((size_t)(2 * pr + 1) * n8 + i) * 8 + c] = a1[1];
                        sQ[((size_t)(2 * pr + 0) * n8 + i) * 8 + c] = qv[0];
                        sQ[((size_t)(2 * pr + 1) * n8 + i) * 8 + c] = qv[1];
                    }
                    if (MODE != SWS_GRID) {
                        // cosine-type analysis (T, S).  odd k: difference of a pair at its own position (L -> jq,
                        // R -> Mq+jq).  even k: pair sums, folded once more into the k' even / k' odd classes.
                        const long long oT0 = prd_off(1, i, jq), oT1 = prd_off(1, i, Mq + jq);
                        const long long oS0 = prd_off(2, i, jq), oS1 = prd_off(2, i, Mq + jq);
                        const double tL = nt[0] + nt[1], tR = nt[2] + nt[3], sL = ns[0] + ns[1], sR = ns[2] + ns[3];
                        prd[pps + oT0] = nt[0] - nt[1]; prd[pps + oT1] = nt[2] - nt[3];
                        prd[oT0] = tL + tR;             prd[oT1] = tL - tR;
                        prd[pps + oS0] = ns[0] - ns[1]; prd[pps + oS1] = ns[2] - ns[3];
                        prd[oS0] = sL + sR;             prd[oS1] = sL - sR;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NTHR_E) : "memory");
            if (lane == 0) mbar_arrive(&bar_eo_free);
            if (MODE != SWS_GRID) {
                // Dr @ (n x 32 points) on the tensor pipe: DFMAs issued from here would queue behind the MMA warps'
                // DMMAs on the shared fp64 pipe one by one.  Warp e < NT8 owns radial rows 8e..8e+7 of all four
                // point classes; accumulator element (row g, columns 2t, 2t+1).
                const int ew = warp - (NMMA + 1);
                if (ew < NT8) {
                    double acc[4][2] = {};
                    const double* ar = sDr + (ew * 8 + gq) * LDD + tq;
                    const double* br = sA1 + tq * 8 + gq;
#pragma unroll
                    for (int ks = 0; ks < n8 / 4; ++ks) {
                        const double av = ar[ks * 4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) mma884(acc[x][0], acc[x][1], av, br[(x * n8 + ks * 4) * 8]);
                    }
                    const int i = ew * 8 + gq;
                    if (i < n) {
                        const int c = 2 * tq, jq = jt * 8 + c;
                        double v[4][2];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const double2 q = *reinterpret_cast<const double2*>(&sQ[((size_t)x * n8 + i) * 8 + c]);
                            v[x][0] = acc[x][0] - q.x; v[x][1] = acc[x][1] - q.y;
                        }
                        // sine-type analysis (psi).  odd k: pair sums at their own positions.  even k: pair
                        // differences, class k' even: L - R, class k' odd: L + R.   (columns c, c+1: one 16-byte store)
                        const long long o0 = prd_off(0, i, jq), o1 = prd_off(0, i, Mq + jq);
                        double2 s0, s1, d0, d1;
                        s0.x = v[0][0] + v[1][0]; s0.y = v[0][1] + v[1][1];
                        s1.x = v[2][0] + v[3][0]; s1.y = v[2][1] + v[3][1];
                        const double dLx = v[0][0] - v[1][0], dLy = v[0][1] - v[1][1];
                        const double dRx = v[2][0] - v[3][0], dRy = v[2][1] - v[3][1];
                        d0.x = dLx - dRx; d0.y = dLy - dRy;
                        d1.x = dLx + dRx; d1.y = dLy + dRy;
                        *reinterpret_cast<double2*>(&prd[pps + o0]) = s0;
                        *reinterpret_cast<double2*>(&prd[pps + o1]) = s1;
                        *reinterpret_cast<double2*>(&prd[o0]) = d0;
                        *reinterpret_cast<double2*>(&prd[o1]) = d1;
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(NTHR_E) : "memory");
            }
        }
    }
}

}  // namespace sddc
