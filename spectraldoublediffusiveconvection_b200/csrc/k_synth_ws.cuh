// Stage 2, hot-path variant: persistent, warp-specialised synthesis kernel for NLIN_FX.
//
// Same mathematics, operand layouts and output as synth_kernel<NT8, EPI_FX> (k_synth.cuh), different schedule:
//   * one persistent CTA per SM walks over (member, column tile) work items, W = 16 mirror pairs per tile;
//   * 12 MMA warps (3 per SM sub-partition, so the four tensor pipes are evenly loaded; each warp owns 6 of the 36
//     8-row tiles of one parity and both column tiles) run the DMMA main loop, a producer warp streams the operand
//     stages with TMA bulk copies through an mbarrier ring that runs ahead across tile boundaries, and 8 epilogue
//     warps turn the E/O accumulators of the *previous* tile into folded grid products while the MMA warps already
//     work on the next tile -- the tensor pipe no longer idles during prologue and epilogue.
//   Instantiated for NT8 = 4 (24 < n <= 32, i.e. N_r = 26..33); other shapes use synth_kernel.
#pragma once
#include "common.cuh"
#include "k_synth.cuh"
#include <type_traits>

namespace sddc {

#ifndef SWS_KS_PER_STAGE
#define SWS_KS_PER_STAGE 2
#endif
// pipeline granularity: MMA k-steps (4 wavenumbers each) per stage, and ring depth (same bytes in flight either way)
constexpr int SWS_KS = SWS_KS_PER_STAGE, SWS_STAGES = 6 / SWS_KS_PER_STAGE;
#ifndef SWS_NMMA_CFG
#define SWS_NMMA_CFG 12
#endif
#ifndef SWS_NEW_CFG
#define SWS_NEW_CFG 8
#endif
constexpr int SWS_NT = 2, SWS_W = 16, SWS_NEW = SWS_NEW_CFG;  // MMA + 1 producer + epilogue warps
constexpr int SWS_NMMA = SWS_NMMA_CFG, SWS_TPW = 72 / SWS_NMMA, SWS_NTHR = 32 * (SWS_NMMA + 1 + SWS_NEW);

__host__ __device__ inline size_t synth_ws_smem_doubles(int n, int n8) {
    const size_t rs = 9 * (size_t)n8;
    const size_t stage = (size_t)SWS_KS * 2 * rs * 4 + (size_t)SWS_KS * 4 * SWS_W * 4;
    return SWS_STAGES * stage + 2 * rs * SWS_W + 2 * (size_t)n * SWS_W + (size_t)n * n;
}

// MODE: SWS_FX   products of the fields of one state (NLIN_FX);
//       SWS_GRID no products: the nine grid fields of the state are stored to p.gridc (base state of a Newton /
//                GMRES solve: computed once, reused by every Jacobian-vector product);
//       SWS_JVPC the MMA operand is the perturbation dv, the base-state grid fields are read back from p.gridc and
//                the bilinear products of NLIN_DFX (Matrix_Operators.py:884-887) are formed -- a JVP then costs one
//                synthesis instead of two.
enum { SWS_FX = 0, SWS_GRID = 1, SWS_JVPC = 2 };

template <int NT8, int MODE>
__global__ void __launch_bounds__(SWS_NTHR, 1) synth_ws_kernel(SynthParams p, int ntiles_j, int nwork) {
    constexpr int NF = 9, RS = NF * NT8 * 8, NT = SWS_NT, W = SWS_W, KS = SWS_KS, LDE = W;
    constexpr int A_SET = KS * 2 * RS * 4, B_ST = KS * 4 * W * 4, STAGE = A_SET + B_ST;
    constexpr int NS = SWS_STAGES, NMMA = SWS_NMMA, NEW = SWS_NEW, NTHR_E = 32 * NEW;
    constexpr int n8 = NT8 * 8, ROWS3 = 3 * n8;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_eo_full, bar_eo_free;
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int nchunk = g.Khp / (4 * KS), n = g.n;   // a stage is KS consecutive [ks] blocks of the tile-major operands
    double* sEO = smem + (size_t)NS * STAGE;      // [2 par][RS][LDE]
    double* sA1 = sEO + (size_t)2 * RS * LDE;     // [2 mirror][n][W]
    double* sDr = sA1 + (size_t)2 * n * W;        // [n][n]

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NMMA); }
        mbar_init(&bar_eo_full, NMMA);
        mbar_init(&bar_eo_free, NEW);
        mbar_fence_init();
    }
    for (int idx = tid; idx < n * n; idx += SWS_NTHR) sDr[idx] = p.Dr[idx];
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
        // ---------------- MMA warps: (parity, group of 6 row tiles) ----------------
        constexpr int TPW = SWS_TPW;
        constexpr int WPP = NMMA / 2;   // MMA warps per parity
        static_assert(NF * NT8 == WPP * TPW, "warp-specialised synthesis is laid out for 36 row tiles per parity");
        const int par = warp / WPP, q = warp - par * WPP;
        const int tile0 = q * TPW;                       // first 8-row tile of this warp (fields = tile / NT8)
        // table type per tile: fields 0-4 cosine (tiles 0..19), fields 5-8 sine (tiles 20..35); only q == 3 is mixed
        const int split = min(TPW, max(0, 5 * NT8 - tile0));   // tiles [0, split) cosine, [split, TPW) sine
        const int a_off = ((par * RS) + tile0 * 8 + gq) * 4 + tq;
        const int b_cos = A_SET + ((0 * 2 + par) * W + gq) * 4 + tq;
        const int b_sin = A_SET + ((1 * 2 + par) * W + gq) * 4 + tq;
        int st = 0, ph = 0, tcount = 0;
        double acc[TPW][NT][2];
        // SPLIT is the compile-time number of leading cosine tiles (6, 2 or 0 for NT8 = 4)
        auto chunk_mma = [&](const double* sS, auto split_tag) {
            constexpr int SPLIT = decltype(split_tag)::value;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double bc[NT], bs[NT], af[TPW];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (SPLIT > 0) bc[nt] = sS[b_cos + ks * 4 * W * 4 + nt * 32];
                    if (SPLIT < TPW) bs[nt] = sS[b_sin + ks * 4 * W * 4 + nt * 32];
                }
#pragma unroll
                for (int mt = 0; mt < TPW; ++mt) af[mt] = sS[a_off + ks * 2 * RS * 4 + mt * 32];
#pragma unroll
                for (int mt = 0; mt < TPW; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        mma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], mt < SPLIT ? bc[nt] : bs[nt]);
            }
        };
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tcount) {
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            for (int c = 0; c < nchunk; ++c) {
                mbar_wait(&bar_full[st], ph);
                const double* sS = smem + (size_t)st * STAGE;
                if (split == TPW) chunk_mma(sS, std::integral_constant<int, TPW>{});
                else if (split == 0) chunk_mma(sS, std::integral_constant<int, 0>{});
                else chunk_mma(sS, std::integral_constant<int, (5 * NT8) % TPW>{});
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[st]);
                if (++st == NS) { st = 0; ph ^= 1; }
            }
            // hand the accumulators to the epilogue warps
            if (tcount > 0) mbar_wait(&bar_eo_free, (tcount - 1) & 1);
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt) {
                const int row = (tile0 + mt) * 8 + gq;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2*>(&sEO[((size_t)par * RS + row) * LDE + nt * 8 + 2 * tq]) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_eo_full);
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int et = tid - 32 * (NMMA + 1);  // 0 .. NTHR_E-1
        const double* E = sEO;
        const double* O = sEO + (size_t)RS * LDE;
        const int npts = n * W;
        constexpr int PTS = (n8 * W + NTHR_E - 1) / NTHR_E;
        int tcount = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tcount) {
            const int b = w / ntiles_j, jt = w - b * ntiles_j;
            double* prd = p.prd + (long long)b * 2 * g.Mhp * ROWS3;
            const long long pps = (long long)g.Mhp * ROWS3;
            auto prd_off = [&](int f, int i, int c) {
                const int jp = jt * W + c;
                return ((long long)(jp >> 2) * ROWS3 + f * n8 + i) * 4 + (jp & 3);
            };
            mbar_wait(&bar_eo_full, tcount & 1);
            double qv[PTS][2];
            // cached base-state grid fields: gridc[b][a][mirror][i][j']
            double* gc = (MODE != SWS_FX) ? p.gridc + (long long)b * 9 * 2 * n8 * g.Mhp : nullptr;
            const long long gms = (long long)n8 * g.Mhp;   // mirror stride; field stride = 2*gms
#pragma unroll
            for (int s = 0; s < PTS; ++s) {
                const int pt = et + s * NTHR_E;
                qv[s][0] = qv[s][1] = 0.0;
                if (pt < npts) {
                    const int i = pt / W, c = pt - i * W;
                    double f0[9], f1[9];
#pragma unroll
                    for (int a = 0; a < 9; ++a) {
                        const double e = E[(size_t)(a * n8 + i) * LDE + c], o = O[(size_t)(a * n8 + i) * LDE + c];
                        if (a < 5) { f0[a] = e + o; f1[a] = e - o; } else { f0[a] = o + e; f1[a] = o - e; }
                    }
                    if (MODE == SWS_GRID) {
                        const long long go = (long long)i * g.Mhp + jt * W + c;
#pragma unroll
                        for (int a = 0; a < 9; ++a) {
                            gc[(a * 2 + 0) * gms + go] = f0[a];
                            gc[(a * 2 + 1) * gms + go] = f1[a];
                        }
                        continue;
                    }
                    double a1_0, a1_1, nt0, nt1, ns0, ns1;
                    if (MODE == SWS_FX) {
                        a1_0 = f0[0] * f0[5];                     a1_1 = f1[0] * f1[5];
                        qv[s][0] = f0[1] * f0[5] + f0[6] * f0[2]; qv[s][1] = f1[1] * f1[5] + f1[6] * f1[2];
                        nt0 = f0[0] * f0[3] - f0[6] * f0[7];      nt1 = f1[0] * f1[3] - f1[6] * f1[7];
                        ns0 = f0[0] * f0[4] - f0[6] * f0[8];      ns1 = f1[0] * f1[4] - f1[6] * f1[8];
                    } else {
                        // h = perturbation fields (just synthesised), f = base-state fields (cached)
                        double h0[9], h1[9];
                        const long long go = (long long)i * g.Mhp + jt * W + c;
#pragma unroll
                        for (int a = 0; a < 9; ++a) {
                            h0[a] = f0[a]; h1[a] = f1[a];
                            f0[a] = gc[(a * 2 + 0) * gms + go];
                            f1[a] = gc[(a * 2 + 1) * gms + go];
                        }
                        a1_0 = f0[0] * h0[5] + h0[0] * f0[5];
                        a1_1 = f1[0] * h1[5] + h1[0] * f1[5];
                        qv[s][0] = (f0[1] * h0[5] + f0[6] * h0[2]) + (h0[1] * f0[5] + h0[6] * f0[2]);
                        qv[s][1] = (f1[1] * h1[5] + f1[6] * h1[2]) + (h1[1] * f1[5] + h1[6] * f1[2]);
                        nt0 = (h0[0] * f0[3] - h0[6] * f0[7]) + (f0[0] * h0[3] - f0[6] * h0[7]);
                        nt1 = (h1[0] * f1[3] - h1[6] * f1[7]) + (f1[0] * h1[3] - f1[6] * h1[7]);
                        ns0 = (h0[0] * f0[4] - h0[6] * f0[8]) + (f0[0] * h0[4] - f0[6] * h0[8]);
                        ns1 = (h1[0] * f1[4] - h1[6] * f1[8]) + (f1[0] * h1[4] - f1[6] * h1[8]);
                    }
                    sA1[(size_t)i * W + c] = a1_0;
                    sA1[(size_t)(n + i) * W + c] = a1_1;
                    const long long oT = prd_off(1, i, c), oS = prd_off(2, i, c);
                    prd[oT] = nt0 + nt1;
                    prd[pps + oT] = nt0 - nt1;
                    prd[oS] = ns0 + ns1;
                    prd[pps + oS] = ns0 - ns1;
                }
            }
            // all E/O reads of this tile are done -> the MMA warps may overwrite sEO; sA1 is complete
            asm volatile("bar.sync 1, %0;" ::"n"(NTHR_E) : "memory");
            if (lane == 0) mbar_arrive(&bar_eo_free);
            if (MODE != SWS_GRID) {   // compile-time: the grid-cache mode has no Dr stage
#pragma unroll
            for (int s = 0; s < PTS; ++s) {
                const int pt = et + s * NTHR_E;
                if (pt < npts) {
                    const int i = pt / W, c = pt - i * W;
                    double v0 = 0.0, v1 = 0.0;
                    for (int ip = 0; ip < n; ++ip) {
                        const double dr = sDr[i * n + ip];
                        v0 = fma(dr, sA1[(size_t)ip * W + c], v0);
                        v1 = fma(dr, sA1[(size_t)(n + ip) * W + c], v1);
                    }
                    v0 -= qv[s][0];
                    v1 -= qv[s][1];
                    const long long o = prd_off(0, i, c);
                    prd[pps + o] = v0 + v1;
                    prd[o] = v0 - v1;
                }
            }
            // sA1 is rewritten by the next tile's first phase only after every epilogue thread is past this point
            asm volatile("bar.sync 2, %0;" ::"n"(NTHR_E) : "memory");
            }
        }
    }
}

}  // namespace sddc
