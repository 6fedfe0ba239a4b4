// Stage 3: latitudinal analysis (scaled DCT-II / DST-II, truncated to K modes) as a parity-split DMMA GEMM of the
// folded grid products against the (scaled) cosine/sine tables, fused with the un-shift of the sine coefficients,
// the equatorial-symmetry mask, the -dt factor and the linear right-hand side.
//
// Reference semantics: NLIN_FX step 4-5 (Matrix_Operators.py:797-804), Step_Python (Main.py:262,271,276,280).
//     F_hat[(f,i), k=2k'+p] = sum_{j'} PRD[f][p][i][j'] * TAB2[type(f)][p][k'][j']
// Both operands are tile-major in global memory (one contiguous block per pipeline stage, fetched by one TMA
// bulk copy each):   PRD  : [b][par][chunk][ks][row = f*n8+i][4]      (written by the synthesis epilogue)
//                    TAB2 : [par][column tile][chunk][ks][type][col][4]  with the 2/M (1/M) scaling folded in
#pragma once
#include "common.cuh"

namespace sddc {

struct AnaParams {
    const double* prd;
    const double* tab2;
    double* out;         // F(X): state layout [B][3N] (bstride == 0) or solve-major [3][K][bstride][n8+2] (k_solve.cuh);
                         // the -dt factor and the linear terms are applied by the solve kernel
    long long bstride;
    int quarter;         // products / table follow the second mirror level (k_synth_wsq.cuh): even-k CTAs contract over
                         // M/4 positions only and own one class (k' even or k' odd) of output columns
    Geo g;
};

constexpr int ANA_KC = 8;
constexpr int ANA_KS = ANA_KC / 4;
constexpr int ANA_MAX_STAGES = 4;

#ifndef ANA_NT
#define ANA_NT 2
#endif
// column tiles per warp: 2 keeps the kernel at <= 80 registers so that two CTAs share an SM and one CTA's prologue /
// scattered-store epilogue overlaps the other's DMMA main loop
__host__ __device__ constexpr int ana_nt_for(int nt8) { return nt8 <= 5 ? ANA_NT : 2; }

__host__ __device__ inline size_t ana_stage_doubles(int nt8) {
    return (size_t)ANA_KS * (3 * nt8 * 8) * 4 + (size_t)ANA_KS * 2 * (ana_nt_for(nt8) * 32) * 4;
}

// grid = (Khp2 / KT3, 2 parities, B), block = 13 warps: consumer warp = (field f, column group cg) owns the NT8 row
// tiles of its field (psi rows use the sine table, T and S rows the cosine table) and NT3 column tiles; the last
// warp is the TMA producer.
template <int NT8>
__global__ void __launch_bounds__(416, (NT8 <= 5 && ANA_NT == 2) ? 2 : 1) analysis_kernel(AnaParams p, int nstage) {
    constexpr int NT3 = ana_nt_for(NT8), KT3 = NT3 * 32, KS = ANA_KS, NCW = 12, NTHR = 416;
    constexpr int ROWS3 = 3 * NT8 * 8;
    constexpr int A_ST = KS * ROWS3 * 4, B_ST = KS * 2 * KT3 * 4, STAGE = A_ST + B_ST;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[ANA_MAX_STAGES], bar_empty[ANA_MAX_STAGES];
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int wf = warp >> 2, cg = warp & 3;
    const int kt = blockIdx.x, par = blockIdx.y, b = blockIdx.z;
    const int n = g.n, K = g.K, N = g.N, Mhp = g.Mhp;
    const bool sm = p.bstride != 0;
    const int LDG = NT8 * 8 + SDDC_SM_PAD;
    // element (field f, block blk, radial i) of member b
    auto out_at = [&](int f, int blk, int i) -> double& {
        return sm ? p.out[(((long long)f * K + blk) * p.bstride + b) * LDG + i]
                  : p.out[(long long)b * 3 * N + (long long)f * N + (long long)blk * n + i];
    };

    if (g.symmetric && par == 1) {
        // every odd-k output block is masked (Vecs_to_X symmetric branch, Matrix_Operators.py:536-556)
        for (int idx = tid; idx < KT3 * n; idx += NTHR) {
            const int kp = kt * KT3 + idx / n, i = idx % n;
            if (kp >= g.Kh) continue;
            const int k = 2 * kp + 1;
            out_at(0, k - 1, i) = 0.0;
            out_at(1, k, i) = 0.0;
            out_at(2, k, i) = 0.0;
        }
        return;
    }

    const bool qpar0 = p.quarter && par == 0;
    const int nkt_half = gridDim.x / 2, cls = qpar0 ? kt / nkt_half : 0;
    const int nchunk_full = Mhp / ANA_KC;
    const int nchunk = qpar0 ? nchunk_full / 2 : nchunk_full;   // even-k CTAs: positions of their class only
    const int chunk0 = qpar0 ? cls * nchunk : 0;
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    double acc[NT8][NT3][2];
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    if (warp == NCW) {
        if (lane == 0) {
            const double* gA = p.prd + ((long long)b * 2 + par) * Mhp * ROWS3 + (long long)chunk0 * A_ST;
            const double* gB = p.tab2 + ((long long)par * gridDim.x + kt) * nchunk_full * B_ST;
            int st = 0, ph = 0;
            for (int c = 0; c < nchunk; ++c) {
                if (c >= nstage) mbar_wait(&bar_empty[st], ph ^ 1);
                double* sA = smem + (size_t)st * STAGE;
                mbar_expect_tx(&bar_full[st], (unsigned)(STAGE * sizeof(double)));
                bulk_g2s(sA, gA + (long long)c * A_ST, A_ST * sizeof(double), &bar_full[st]);
                bulk_g2s(sA + A_ST, gB + (long long)c * B_ST, B_ST * sizeof(double), &bar_full[st]);
                if (++st == nstage) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        const int ty = (wf == 0) ? 1 : 0;
        const int a_off = (wf * NT8 * 8 + gq) * 4 + tq;                       // + ks*ROWS3*4 + mt*32
        const int b_off = A_ST + (ty * KT3 + cg * NT3 * 8 + gq) * 4 + tq;     // + ks*2*KT3*4 + nt*32
        int st = 0, ph = 0;
        for (int c = 0; c < nchunk; ++c) {
            mbar_wait(&bar_full[st], ph);
            const double* sS = smem + (size_t)st * STAGE;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double bf[NT3], af[NT8];
#pragma unroll
                for (int nt = 0; nt < NT3; ++nt) bf[nt] = sS[b_off + ks * 2 * KT3 * 4 + nt * 32];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt) af[mt] = sS[a_off + ks * ROWS3 * 4 + mt * 32];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT3; ++nt) mma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            if (++st == nstage) { st = 0; ph ^= 1; }
        }
    }
    if (warp >= NCW) return;

    // ---- epilogue: C fragment (row = (f,i), col = k') -> state layout [f][block][i] ----
    const int f = wf;
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt) {
        const int i = mt * 8 + gq;
        if (i < n) {
#pragma unroll
            for (int nt = 0; nt < NT3; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int colt = (cg * NT3 + nt) * 8 + 2 * tq + e;
                    const int kp = qpar0 ? 2 * ((kt % nkt_half) * KT3 + colt) + cls : kt * KT3 + colt;
                    if (kp >= g.Kh) continue;
                    const int k = 2 * kp + par;
                    const double v = acc[mt][nt][e];
                    if (f == 0) {
                        // sinusoid index k -> code block k-1; k = 0 is dropped and block K-1 gets no
                        // nonlinear contribution (Matrix_Operators.py:802)
                        const int blk = (k >= 1) ? k - 1 : K - 1;
                        const double fv = (k >= 1) ? v : 0.0;
                        const bool keep = !(g.symmetric && (blk & 1) == 0);
                        out_at(0, blk, i) = keep ? fv : 0.0;
                    } else {
                        out_at(f, k, i) = v;  // odd k never reaches here when symmetric
                    }
                }
            }
        }
    }
}

}  // namespace sddc
