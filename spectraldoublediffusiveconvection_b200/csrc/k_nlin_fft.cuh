// Stage 2+3 of a member-step in the FFT formulation (fft_core.h): the nonlinear term NLIN_FX / NLIN_DFX
// (Matrix_Operators.py:743-898) with the latitudinal transforms as mixed-radix complex FFTs in shared memory.
//
//   nlin_fft_kernel : persistent, one CTA per SM.  A CTA holds NW independent workers of 64 threads (two warps);
//                     a worker takes one (member, radial row) at a time through
//                         build | radix-8 | radix-RD | radix-6 + products + radix-6 | radix-RD | radix-8 | post
//                     with a named barrier (bar.sync id, 64) between phases, so the workers of a CTA are never in
//                     lock step and fill each other's barrier / memory latencies.  All transform data stays in the
//                     worker's shared-memory planes; HBM sees 7 coefficient rows in and 4 spectral rows out.
//   post_kernel     : Dr @ DST(JT*omega) - DST(..) per wavenumber, un-shift of the sine coefficients
//                     (Matrix_Operators.py:797-802), equatorial-symmetry mask, and the transposition into the state /
//                     solve-major layout.
#pragma once
#include "common.cuh"
#include "fft_core.h"

namespace sddc {

struct NlinFftParams {
    const double* coef0;  // [rows][7][K] spectral rows of the (base) state, rows = B * n
    const double* coef1;  // [rows][7][K] rows of the perturbation (two-state mode)
    double* spec;         // [rows][4][K] analysed products
    const double* tab;    // fftp::tab_doubles<M>() table doubles (fftp::fill_tables)
    int nrows;
    // fused finishing stage (what post_kernel does otherwise): the worker that completes the last radial row of a
    // member applies Dr @ and transposes that member's products into the state / solve-major layout
    int* next_row;        // row counter, zeroed before every launch: workers claim rows dynamically, so that a worker
                          // delayed by a finishing stage simply takes fewer rows (with static striding the delayed
                          // worker arrives last again and again and ends up finishing every member: measured 30x slower)
    int* done;            // [B] arrival counters, zero between launches; nullptr: post_kernel runs as a separate launch
    const double* DrT;    // [n][n8]: DrT[i'][i] = Dr[i][i']
    double* out;          // F(X): state layout [B][3N] (bstride == 0) or solve-major [3][K][bstride][n8+2]
    long long bstride;
    int ftc;              // column tile of the finishing stage: 4 n (ftc + 1) doubles fit in one worker's planes
    Geo g;
};

#ifndef NLIN_FFT_NW
#define NLIN_FFT_NW 7  // workers (of 64 threads) per CTA of the one-state kernel: 14 warps cap registers at 128 (a few spilled
                       // bytes) but measured 2.5 % faster than 6 workers at 168 registers
#endif

template <int M>
__host__ __device__ constexpr int nlin_fft_tab_pad() { return (fftp::tab_doubles<M>() + 15) / 16 * 16; }
__host__ __device__ inline int nlin_fft_dr_pad(int n, int n8) { return (n * n8 + 15) / 16 * 16; }
template <int M, bool DFX>
__host__ __device__ constexpr size_t nlin_fft_worker_doubles() { return (size_t)2 * (DFX ? 9 : 5) * fftp::Cfg<M>::PL; }
// dynamic shared memory: tables | DrT (fused finishing stage only) | nw workers
template <int M, bool DFX>
__host__ __device__ inline size_t nlin_fft_smem_bytes(int nw, int dr_doubles) {
    return sizeof(double) * ((size_t)nlin_fft_tab_pad<M>() + dr_doubles + (size_t)nw * nlin_fft_worker_doubles<M, DFX>());
}

template <int NT = 64>
__device__ __forceinline__ void worker_sync(int w) {
    asm volatile("bar.sync %0, %1;" ::"r"(w + 1), "n"(NT) : "memory");
}

// Finishing stage of member b by one worker (64 threads), tile by tile of `tc` sinusoid columns:
//   F_psi[i][k] = sum_i' Dr[i][i'] P1[i'][k] - P2[i][k]  -> code block k-1 (Matrix_Operators.py:791,797-802),
//   F_T, F_S    = analysed products, equatorial-symmetry mask, state / solve-major layout.
// The products were written by other SMs; the caller has fenced after observing the last arrival.
__device__ __forceinline__ void finish_member(const NlinFftParams& p, int b, int w, int t, double* sT, const double* sD) {
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K, N = g.N, tc = p.ftc, LDT = tc + 1;
    const double* sb = p.spec + (size_t)b * n * 4 * K;
    const bool sm = p.bstride != 0;
    const int LDG = n8 + 2;
    auto out_at = [&](int f, int blk, int i) -> double& {
        return sm ? p.out[(((long long)f * K + blk) * p.bstride + b) * LDG + i]
                  : p.out[(long long)b * 3 * N + (long long)f * N + (long long)blk * n + i];
    };
    for (int k0 = 0; k0 < K; k0 += tc) {
        // all tile loads in flight at once (asynchronous copies; none of these lines can be in this SM's L1: they were
        // written by stores, which do not allocate, and L1 does not survive a kernel boundary)
        for (int idx = t; idx < 4 * n * tc; idx += 64) {
            const int c = idx % tc, fi = idx / tc, f = fi & 3, i = fi >> 2;
            const bool ok = k0 + c < K;
            cp_async8_zfill(&sT[(f * n + i) * LDT + c], ok ? sb + ((size_t)i * 4 + f) * K + k0 + c : sb, ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
        worker_sync(w);
        for (int u = t; u < tc * ((n + 3) / 4); u += 64) {
            const int col = u % tc, i0 = (u / tc) * 4;   // i0 + 3 < n8: padded operator columns are zero
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int ip = 0; ip < n; ++ip) {
                const double p1 = sT[ip * LDT + col];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = fma(sD[ip * n8 + i0 + r], p1, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (i0 + r < n) sT[(n + i0 + r) * LDT + col] = acc[r] - sT[(n + i0 + r) * LDT + col];
        }
        worker_sync(w);
        for (int idx = t; idx < 3 * tc * n; idx += 64) {
            const int i = idx % n, fc = idx / n, col = fc % tc, f = fc / tc;
            const int k = k0 + col;
            if (k >= K) continue;
            const bool masked = g.symmetric && (k & 1);  // every odd sinusoid index is masked (Matrix_Operators.py:536-556)
            const double v = masked ? 0.0 : sT[((f + 1) * n + i) * LDT + col];
            if (f == 0) {
                if (k == 0) out_at(0, K - 1, i) = 0.0;   // block K-1 gets no nonlinear contribution (Matrix_Operators.py:802)
                else out_at(0, k - 1, i) = v;
            } else {
                out_at(f, k, i) = v;
            }
        }
        worker_sync(w);
    }
}

// NT = threads per worker: 64, or 128 where the last pass has 128 columns (M = 768)
template <int M, bool DFX, int NW, int NT = 64>
__global__ void __launch_bounds__(NT * NW, 1) nlin_fft_kernel(NlinFftParams p) {
    using namespace fftp;
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NF = DFX ? 9 : 5;
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_last[NW], s_row[NW];
    double* stab = smem;
    double* sD = smem + nlin_fft_tab_pad<M>();
    const int dr_doubles = p.done ? nlin_fft_dr_pad(p.g.n, p.g.n8) : 0;
    for (int i = threadIdx.x; i < tab_doubles<M>(); i += NT * NW) stab[i] = p.tab[i];
    if (p.done)
        for (int i = threadIdx.x; i < p.g.n * p.g.n8; i += NT * NW) sD[i] = p.DrT[i];
    __syncthreads();
    const Tables tb = make_tables<M>(stab);
    const int w = threadIdx.x / NT, t = threadIdx.x % NT;
    double* buf = sD + dr_doubles + (size_t)w * nlin_fft_worker_doubles<M, DFX>();
    C tw[Cfg<M>::RD];
    load_tw<M>(t, tb, tw);
    const int stride = gridDim.x * NW;
    for (;;) {
        if (t == 0) s_row[w] = atomicAdd(p.next_row, 1);
        worker_sync<NT>(w);
        const int row = s_row[w];
        if (row >= p.nrows) break;
        if (DFX) {
            build<M, 1, NT>(t, p.coef0 + (size_t)row * 7 * K, buf, tb, p.coef1 + (size_t)row * 7 * K);
            build<M, 2, NT>(t, p.coef1 + (size_t)row * 7 * K, buf + 10 * PL, tb);
        } else {
            build<M, 0, NT>(t, p.coef0 + (size_t)row * 7 * K, buf, tb);
        }
        // pull the row that will be claimed one round from now from HBM into L2 while this one is transformed
        if (row + stride < p.nrows) {
            const char* nx = reinterpret_cast<const char*>(p.coef0 + (size_t)(row + stride) * 7 * K);
            for (int o = t * 128; o < 7 * K * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
            if (DFX) {
                const char* nx1 = reinterpret_cast<const char*>(p.coef1 + (size_t)(row + stride) * 7 * K);
                for (int o = t * 128; o < 7 * K * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx1 + o));
            }
        }
        worker_sync<NT>(w);
        pass_c<M, NF, +1, NT>(t, buf);
        worker_sync<NT>(w);
        pass_d<M, NF, +1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        i3f1<M, DFX, NT>(t, buf, tb);
        worker_sync<NT>(w);
        pass_d<M, 2, -1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        pass_c<M, 2, -1, NT>(t, buf);
        worker_sync<NT>(w);
        post<M, NT>(t, buf, p.spec + (size_t)row * 4 * K, tb);
        if (p.done) {
            // last-arriver pattern: publish this row, count it, and let the worker that completes a member finish it
            __threadfence();
            worker_sync<NT>(w);
            const int b = row / p.g.n;
            if (t == 0) s_last[w] = (atomicAdd(p.done + b, 1) == p.g.n - 1);
            worker_sync<NT>(w);
            if (s_last[w]) {
                __threadfence();
                if (NT == 64) finish_member(p, b, w, t, buf, sD);   // the opt-in fused stage exists for 64-thread workers only
                if (t == 0) p.done[b] = 0;
            }
        }
        worker_sync<NT>(w);
    }
}

// Two-state (JVP) kernel with the perturbation transformed two fields at a time against the resident base planes:
// 14 planes per worker instead of 18, so five 64-thread workers fit at M = 384 (four with the one-round kernel) and two
// 128-thread workers at M = 768 (one).  One radix-6 column per thread (NT == M / 6); what a thread carries between the
// two rounds (JT', Dpsi' and the partial products at its six points) stays in registers across the barriers.  Same
// arithmetic as nlin_fft_kernel<M, true>: results are bit-identical (tests/test_fft_core_cpu.py).
template <int M>
__host__ __device__ constexpr size_t nlin_fft2_worker_doubles() { return (size_t)14 * fftp::Cfg<M>::PL; }
template <int M>
__host__ __device__ constexpr size_t nlin_fft2_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)nlin_fft_tab_pad<M>() + (size_t)nw * nlin_fft2_worker_doubles<M>());
}

template <int M, int NW, int NT>
__global__ void __launch_bounds__(NT * NW, 1) nlin_fft2_kernel(NlinFftParams p) {
    using namespace fftp;
    static_assert(NT == Cfg<M>::L, "one radix-6 column per thread");
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL;
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_row[NW];
    double* stab = smem;
    for (int i = threadIdx.x; i < tab_doubles<M>(); i += NT * NW) stab[i] = p.tab[i];
    __syncthreads();
    const Tables tb = make_tables<M>(stab);
    const int w = threadIdx.x / NT, t = threadIdx.x % NT;
    double* buf = smem + nlin_fft_tab_pad<M>() + (size_t)w * nlin_fft2_worker_doubles<M>();
    double* pair = buf + 10 * PL;
    C tw[Cfg<M>::RD];
    load_tw<M>(t, tb, tw);
    const int stride = gridDim.x * NW;
    for (;;) {
        if (t == 0) s_row[w] = atomicAdd(p.next_row, 1);
        worker_sync<NT>(w);
        const int row = s_row[w];
        if (row >= p.nrows) break;
        const double* r0 = p.coef0 + (size_t)row * 7 * K;
        const double* r1 = p.coef1 + (size_t)row * 7 * K;
        build<M, 1, NT>(t, r0, buf, tb, r1);
        build<M, 3, NT>(t, r1, pair, tb);
        if (row + stride < p.nrows) {
            const char* nx0 = reinterpret_cast<const char*>(p.coef0 + (size_t)(row + stride) * 7 * K);
            const char* nx1 = reinterpret_cast<const char*>(p.coef1 + (size_t)(row + stride) * 7 * K);
            for (int o = t * 128; o < 7 * K * 8; o += NT * 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx0 + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx1 + o));
            }
        }
        worker_sync<NT>(w);
        pass_c<M, 7, +1, NT>(t, buf);
        worker_sync<NT>(w);
        pass_d<M, 7, +1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        Dfx2State st;
        dfx2_first<M>(t, buf, tb, st);
        worker_sync<NT>(w);                 // every column of the first pair is consumed before its planes are refilled
        build<M, 4, NT>(t, r1, pair, tb);
        worker_sync<NT>(w);
        pass_c<M, 2, +1, NT>(t, pair);
        worker_sync<NT>(w);
        pass_d<M, 2, +1, NT>(t, pair, tw);
        worker_sync<NT>(w);
        dfx2_second<M>(t, buf, tb, st);
        worker_sync<NT>(w);
        pass_d<M, 2, -1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        pass_c<M, 2, -1, NT>(t, buf);
        worker_sync<NT>(w);
        post<M, NT>(t, buf, p.spec + (size_t)row * 4 * K, tb);
        worker_sync<NT>(w);
    }
}

// Kinetic energy on the 3K grid (Main.py:71-134) in the FFT formulation: one complex transform per radial row
// (J_theta(psi)/r and Dr psi packed), weighted sum of squares in the last pass.  kepart[row] = wr[i] * sum_theta.
struct KeFftParams {
    const double* rows;   // [B n][2][K]; or the seven-row blocks of prep_kernel (row_stride = 7K, b_off = K, ascale = 1/r)
    long long row_stride; // doubles between consecutive (member, radial point) rows
    int b_off;            // offset of the sine-type row inside a block
    const double* ascale; // [n] factor of the cosine-type row, or null
    const double* tab;    // fftp::fill_tables<3K>
    const double* Wn;     // [3K] fftp::fill_ke_weights<3K>
    const double* wr;     // [n] radial trapezoid weights
    double* kepart;       // [B n]
    int nrows, n;
};
template <int M>
__host__ __device__ constexpr int ke_fft_shared_doubles() { return nlin_fft_tab_pad<M>() + M; }
template <int M>
__host__ __device__ constexpr size_t ke_fft_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)ke_fft_shared_doubles<M>() + (size_t)nw * 2 * fftp::Cfg<M>::PL);
}

template <int M, int NW>
__global__ void __launch_bounds__(64 * NW, 1) ke_fft_kernel(KeFftParams p) {
    using namespace fftp;
    constexpr int PL = Cfg<M>::PL;
    extern __shared__ __align__(128) double smem[];
    __shared__ double s_part[NW][2];
    double* stab = smem;
    double* sW = smem + nlin_fft_tab_pad<M>();
    for (int i = threadIdx.x; i < tab_doubles<M>(); i += 64 * NW) stab[i] = p.tab[i];
    for (int i = threadIdx.x; i < M; i += 64 * NW) sW[i] = p.Wn[i];
    __syncthreads();
    const Tables tb = make_tables<M>(stab);
    const int w = threadIdx.x >> 6, t = threadIdx.x & 63;
    double* buf = smem + ke_fft_shared_doubles<M>() + (size_t)w * 2 * PL;
    C tw[Cfg<M>::RD];
    load_tw<M>(t, tb, tw);
    for (int row = blockIdx.x * NW + w; row < p.nrows; row += gridDim.x * NW) {
        const double* r = p.rows + (size_t)row * p.row_stride;
        build_ke<M>(t, r, r + p.b_off, p.ascale ? p.ascale[row % p.n] : 1.0, buf, tb);
        worker_sync(w);
        pass_c<M, 1, +1>(t, buf);
        worker_sync(w);
        pass_d<M, 1, +1>(t, buf, tw);
        worker_sync(w);
        const double part = warp_sum(ke6<M>(t, buf, tb, sW));
        if ((t & 31) == 0) s_part[w][t >> 5] = part;
        worker_sync(w);
        if (t == 0) p.kepart[row] = p.wr[row % p.n] * (s_part[w][0] + s_part[w][1]);
        worker_sync(w);
    }
}

struct PostParams {
    const double* spec;  // [B][n][4][K]
    const double* DrT;   // [n][n8]: DrT[i'][i] = Dr[i][i']
    double* out;         // F(X): state layout [B][3N] (bstride == 0) or solve-major [3][K][bstride][n8+2]
    long long bstride;
    Geo g;
};

constexpr int POST_TC = 32;

__host__ __device__ inline size_t post_smem_bytes(int n, int n8) {
    return sizeof(double) * ((size_t)2 * 4 * n * (POST_TC + 1) + (size_t)n * n8);
}

// Persistent: CTA c takes the tiles (member b, 32 sinusoid columns) c, c + gridDim.x, ...; the four product tiles of the
// next tile are in flight (cp.async, two stages) while the current one is contracted with Dr and stored.
__global__ void __launch_bounds__(256) post_kernel(PostParams p, int ntiles) {
    extern __shared__ __align__(128) double smem[];
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K, N = g.N, LDT = POST_TC + 1;
    const int tid = threadIdx.x, nkt = (K + POST_TC - 1) / POST_TC, TS = 4 * n * LDT;
    double* sD = smem + 2 * TS;   // [n][n8]
    for (int idx = tid; idx < n * n8; idx += 256) sD[idx] = p.DrT[idx];
    auto issue = [&](int tile, int stage) {
        const int b = tile / nkt, k0 = (tile - b * nkt) * POST_TC;
        const double* sb = p.spec + (size_t)b * n * 4 * K;
        double* sT = smem + stage * TS;
        for (int idx = tid; idx < 4 * n * POST_TC; idx += 256) {
            const int c = idx & (POST_TC - 1), fi = idx / POST_TC, f = fi & 3, i = fi >> 2;
            const bool ok = k0 + c < K;
            cp_async8_zfill(&sT[(f * n + i) * LDT + c], ok ? sb + ((size_t)i * 4 + f) * K + k0 + c : sb, ok);
        }
    };
    const bool sm = p.bstride != 0;
    const int LDG = n8 + 2;
    int tile = blockIdx.x, stage = 0;
    if (tile < ntiles) issue(tile, 0);
    cp_async_commit();
    for (; tile < ntiles; tile += gridDim.x, stage ^= 1) {
        const int nxt = tile + gridDim.x;
        if (nxt < ntiles) issue(nxt, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int b = tile / nkt, k0 = (tile - b * nkt) * POST_TC;
        double* sT = smem + stage * TS;   // [4][n][33]
        // F_psi[i][k] = sum_i' Dr[i][i'] P1[i'][k] - P2[i][k]: thread = (column, four consecutive rows); the operator row
        // segment is a warp-wide broadcast, the column read is conflict free.  The result replaces P2 (read by no one else).
        for (int i0 = (tid >> 5) * 4; i0 < n; i0 += 32) {   // i0 + 3 < n8: padded operator columns are zero
            const int col = tid & 31;
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double2* d2 = reinterpret_cast<const double2*>(sD + i0);   // 32-byte aligned: two 16-byte broadcasts per row
            const int rs = n8 >> 1;
#pragma unroll 4
            for (int ip = 0; ip < n; ++ip) {
                const double p1 = sT[ip * LDT + col];
                const double2 a = d2[ip * rs], c2 = d2[ip * rs + 1];
                acc[0] = fma(a.x, p1, acc[0]);
                acc[1] = fma(a.y, p1, acc[1]);
                acc[2] = fma(c2.x, p1, acc[2]);
                acc[3] = fma(c2.y, p1, acc[3]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (i0 + r < n) sT[(n + i0 + r) * LDT + col] = acc[r] - sT[(n + i0 + r) * LDT + col];
        }
        __syncthreads();
        // store phase, radial index fastest (runs of n doubles in either layout); (i, column, field) advance incrementally
        {
            int i = tid % n, fc = tid / n;
            const int di = 256 % n, dfc = 256 / n;
            double* ob = sm ? p.out + (long long)b * LDG : p.out + (long long)b * 3 * N;
            const long long blk_stride = sm ? p.bstride * LDG : n, fld_stride = sm ? (long long)K * p.bstride * LDG : N;
            for (; fc < 3 * POST_TC; ) {
                const int col = fc & (POST_TC - 1), f = fc >> 5;
                const int k = k0 + col;
                if (k < K) {
                    const bool masked = g.symmetric && (k & 1);  // every odd sinusoid index is masked (Matrix_Operators.py:536-556)
                    const double v = masked ? 0.0 : sT[((f + 1) * n + i) * LDT + col];
                    // psi: sinusoid index k -> code block k-1; k = 0 is dropped and block K-1 gets no nonlinear
                    // contribution (Matrix_Operators.py:802)
                    const int blk = f == 0 ? (k == 0 ? K - 1 : k - 1) : k;
                    ob[f * fld_stride + blk * blk_stride + i] = (f == 0 && k == 0) ? 0.0 : v;
                }
                i += di; fc += dfc;
                if (i >= n) { i -= n; ++fc; }
            }
        }
        __syncthreads();   // this stage is refilled two iterations from now, by copies issued after this point
    }
    cp_async_wait<0>();
}

}  // namespace sddc
