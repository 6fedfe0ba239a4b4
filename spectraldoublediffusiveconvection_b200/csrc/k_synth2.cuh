// Stage 2, hot-path variant for NLIN_FX at NT8 = 4 (N_r = 26..33): same mathematics, operand layouts and output as
// synth_kernel<4, EPI_FX> (k_synth.cuh) with a schedule tuned on ncu evidence:
//   * 12 MMA warps per CTA (3 per SM sub-partition -> the four tensor pipes are evenly loaded), each owning 6 of the
//     36 8-row tiles of one parity and both column tiles of a 16-column tile;
//   * small CTAs (416 threads, <= 78 registers, 88 KB shared memory) so that TWO CTAs share an SM: one CTA's TMA
//     prologue, barrier bubbles and pointwise epilogue are covered by the other CTA's DMMA main loop;
//   * one MMA k-step (4 wavenumbers) per pipeline stage, 4-stage TMA/mbarrier ring.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "k_synth.cuh"

namespace sddc {

constexpr int S2_NT = 2, S2_W = 16, S2_NMMA = 12, S2_TPW = 6, S2_STAGES = 4, S2_NTHR = 32 * (S2_NMMA + 1);

__host__ __device__ inline size_t synth2_smem_doubles(int n, int n8) {
    const size_t rs = 9 * (size_t)n8;
    const size_t stage = (size_t)2 * rs * 4 + (size_t)4 * S2_W * 4;
    const size_t epi = 2 * rs * S2_W + 2 * (size_t)n * S2_W + (size_t)n * n;
    return S2_STAGES * stage > epi ? S2_STAGES * stage : epi;
}

template <int NT8>
__global__ void __launch_bounds__(S2_NTHR, 2) synth2_kernel(SynthParams p) {
    constexpr int NF = 9, RS = NF * NT8 * 8, NT = S2_NT, W = S2_W, LDE = W, TPW = S2_TPW;
    constexpr int A_SET = 2 * RS * 4, B_ST = 4 * W * 4, STAGE = A_SET + B_ST;   // one [ks] block of each operand
    constexpr int NS = S2_STAGES, NMMA = S2_NMMA, NTHR = S2_NTHR;
    constexpr int n8 = NT8 * 8, ROWS3 = 3 * n8;
    static_assert(NF * NT8 == 6 * TPW, "laid out for 36 row tiles per parity");
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS];
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, jt = blockIdx.x;
    const int nchunk = g.Khp / 4, n = g.n;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NMMA); }
        mbar_fence_init();
    }
    __syncthreads();

    const int par = warp / 6, q = warp - par * 6;   // garbage for the producer warp
    const int tile0 = q * TPW;
    double acc[TPW][NT][2];
#pragma unroll
    for (int mt = 0; mt < TPW; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    if (warp == NMMA) {
        if (lane == 0) {
            const double* gA = p.coef0 + (long long)b * p.coef_stride;
            const double* gB = p.tab + (long long)jt * nchunk * B_ST;
            int st = 0, ph = 0;
            for (int c = 0; c < nchunk; ++c) {
                if (c >= NS) mbar_wait(&bar_empty[st], ph ^ 1);
                double* sA = smem + (size_t)st * STAGE;
                mbar_expect_tx(&bar_full[st], (unsigned)(STAGE * sizeof(double)));
                bulk_g2s(sA, gA + (long long)c * A_SET, A_SET * sizeof(double), &bar_full[st]);
                bulk_g2s(sA + A_SET, gB + (long long)c * B_ST, B_ST * sizeof(double), &bar_full[st]);
                if (++st == NS) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // table type per tile: fields 0-4 cosine (tiles 0..19), fields 5-8 sine (tiles 20..35); only q == 3 is mixed
        const int split = min(TPW, max(0, 5 * NT8 - tile0));
        const int a_off = ((par * RS) + tile0 * 8 + gq) * 4 + tq;
        const int b_cos = A_SET + ((0 * 2 + par) * W + gq) * 4 + tq;
        const int b_sin = A_SET + ((1 * 2 + par) * W + gq) * 4 + tq;
        auto kstep = [&](const double* sS, auto split_tag) {
            constexpr int SPLIT = decltype(split_tag)::value;
            double bc[NT], bs[NT], af[TPW];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (SPLIT > 0) bc[nt] = sS[b_cos + nt * 32];
                if (SPLIT < TPW) bs[nt] = sS[b_sin + nt * 32];
            }
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt) af[mt] = sS[a_off + mt * 32];
#pragma unroll
            for (int mt = 0; mt < TPW; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    mma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], mt < SPLIT ? bc[nt] : bs[nt]);
        };
        int st = 0, ph = 0;
        for (int c = 0; c < nchunk; ++c) {
            mbar_wait(&bar_full[st], ph);
            const double* sS = smem + (size_t)st * STAGE;
            if (split == TPW) kstep(sS, std::integral_constant<int, TPW>{});
            else if (split == 0) kstep(sS, std::integral_constant<int, 0>{});
            else kstep(sS, std::integral_constant<int, (5 * NT8) % TPW>{});
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            if (++st == NS) { st = 0; ph ^= 1; }
        }
    }
    __syncthreads();  // every stage has been consumed; the staging area is reused below

    double* sEO = smem;                          // [2 par][RS][LDE]
    double* sA1 = sEO + (size_t)2 * RS * LDE;    // [2 mirror][n][W]
    double* sDr = sA1 + (size_t)2 * n * W;       // [n][n]
    if (warp < NMMA) {
#pragma unroll
        for (int mt = 0; mt < TPW; ++mt) {
            const int row = (tile0 + mt) * 8 + gq;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(&sEO[((size_t)par * RS + row) * LDE + nt * 8 + 2 * tq]) =
                    make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
    }
    for (int idx = tid; idx < n * n; idx += NTHR) sDr[idx] = p.Dr[idx];
    __syncthreads();

    const double* E = sEO;
    const double* O = sEO + (size_t)RS * LDE;
    const int npts = n * W;
    constexpr int PTS = (n8 * W + NTHR - 1) / NTHR;
    double qv[PTS][2];
    double* prd = p.prd + (long long)b * 2 * g.Mhp * ROWS3;
    const long long pps = (long long)g.Mhp * ROWS3;
    auto prd_off = [&](int f, int i, int c) {
        const int jp = jt * W + c;
        return ((long long)(jp >> 2) * ROWS3 + f * n8 + i) * 4 + (jp & 3);
    };
#pragma unroll
    for (int s = 0; s < PTS; ++s) {
        const int pt = tid + s * NTHR;
        qv[s][0] = qv[s][1] = 0.0;
        if (pt < npts) {
            const int i = pt / W, c = pt - i * W;
            double f0[9], f1[9];
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                const double e = E[(size_t)(a * n8 + i) * LDE + c], o = O[(size_t)(a * n8 + i) * LDE + c];
                if (a < 5) { f0[a] = e + o; f1[a] = e - o; } else { f0[a] = o + e; f1[a] = o - e; }
            }
            sA1[(size_t)i * W + c] = f0[0] * f0[5];
            sA1[(size_t)(n + i) * W + c] = f1[0] * f1[5];
            qv[s][0] = f0[1] * f0[5] + f0[6] * f0[2];
            qv[s][1] = f1[1] * f1[5] + f1[6] * f1[2];
            const double nt0 = f0[0] * f0[3] - f0[6] * f0[7], nt1 = f1[0] * f1[3] - f1[6] * f1[7];
            const double ns0 = f0[0] * f0[4] - f0[6] * f0[8], ns1 = f1[0] * f1[4] - f1[6] * f1[8];
            const long long oT = prd_off(1, i, c), oS = prd_off(2, i, c);
            prd[oT] = nt0 + nt1;
            prd[pps + oT] = nt0 - nt1;
            prd[oS] = ns0 + ns1;
            prd[pps + oS] = ns0 - ns1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PTS; ++s) {
        const int pt = tid + s * NTHR;
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
}

}  // namespace sddc
