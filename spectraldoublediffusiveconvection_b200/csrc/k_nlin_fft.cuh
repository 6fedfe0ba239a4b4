// Stage 2+3 of a member-step in the FFT formulation (fft_core.h, fft_fused.h): the nonlinear term NLIN_FX / NLIN_DFX
// (Matrix_Operators.py:743-898) with the latitudinal transforms as mixed-radix complex FFTs in shared memory.
//
//   nlin_fft_kernel : persistent, one CTA per SM.  A CTA holds NW independent workers of NT threads; a worker takes
//                     one (member, radial row) at a time through
//                         pack + radix-8 | radix-RD | radix-6 + products + radix-6 | radix-RD | radix-8 + unpack
//                     with a named barrier (bar.sync id, NT) between phases, so the workers of a CTA are never in
//                     lock step and fill each other's barrier / memory latencies.  All transform data stays in the
//                     worker's shared-memory planes; HBM sees 7 coefficient rows in and 4 spectral rows out.
//   post_kernel     : Dr @ DST(JT*omega) - DST(..) per wavenumber, un-shift of the sine coefficients
//                     (Matrix_Operators.py:797-802), equatorial-symmetry mask, and the transposition into the state /
//                     solve-major layout.
#pragma once
#include "common.cuh"
#include "fft_core.h"
#include "fft_fused.h"
#include "k_misc.cuh"

namespace sddc {

struct NlinFftParams {
    const double* coef0;  // [rows][7][K] spectral rows of the (base) state, rows = B * n
    const double* coef1;  // [rows][7][K] rows of the perturbation (two-state mode)
    double* spec;         // [rows][4][K] analysed products, every row parity-split (fft_core.h: spec_pos)
    const double* tab;    // fftp::tab_doubles<M>() table doubles (fftp::fill_tables)
    double* grid;         // [rows][7][M] cached grid fields of the base state (MODE 1 writes, MODE 2 reads; fft_fused.h)
    int nrows;
    int* next_row;        // [0] row counter: workers claim rows dynamically; [1] CTAs that have finished -- the last one
                          // zeroes both for the next launch (no memset between the kernels of a step)
};

#ifndef NLIN_FFT_NW
#define NLIN_FFT_NW 8  // workers (of 64 threads) per CTA of the one-state kernel at M <= 384 (24 KB of planes each at M = 384)
#endif

template <int M>
__host__ __device__ constexpr int nlin_fft_tab_pad() { return (fftp::tab_doubles<M>() + 15) / 16 * 16; }
// plane pairs per worker: four transforms of one state, seven of a pair of states (fft_fused.h)
template <int M, bool DFX>
__host__ __device__ constexpr size_t nlin_fft_worker_doubles() { return (size_t)fftp::pairs_doubles<M>(DFX ? 7 : 4); }
// dynamic shared memory: tables | nw workers
template <int M, bool DFX>
__host__ __device__ constexpr size_t nlin_fft_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)nlin_fft_tab_pad<M>() + (size_t)nw * nlin_fft_worker_doubles<M, DFX>());
}

// every claim of this CTA has been made: the last CTA of the launch to get here resets the counters for the next launch
__device__ __forceinline__ void rows_done(int* cnt) {
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(cnt + 1, 1) == (int)gridDim.x - 1) {
        cnt[0] = 0;
        cnt[1] = 0;
    }
}

template <int NT = 64>
__device__ __forceinline__ void worker_sync(int w) {
    asm volatile("bar.sync %0, %1;" ::"r"(w + 1), "n"(NT) : "memory");
}

// NT = threads per worker: 64, or 128 where the radix-6 pass has 128 columns (M = 768)
// MODE (one-state instantiations): 0 products of coef0; 1 GRID: grid fields of coef0 -> p.grid, no analysis;
//                                  2 JVPC: bilinear products of the perturbation rows coef1 with the cached p.grid
template <int M, bool DFX, int NW, int NT = 64, int MODE = 0>
__global__ void __launch_bounds__(NT * NW, 1) nlin_fft_kernel(NlinFftParams p) {
    static_assert(!DFX || MODE == 0, "the cached-base modes are one-state kernels");
    using namespace fftp;
    constexpr int K = Cfg<M>::K, NF = DFX ? 7 : 4;
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_row[NW];
    double* stab = smem;
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < tab_doubles<M>(); i += NT * NW) stab[i] = p.tab[i];
    const int w = threadIdx.x / NT, t = threadIdx.x % NT;
    pdl_wait();   // coefficient rows (prep_kernel) and the row counter (reset by the previous launch's last CTA)
    if (t == 0) s_row[w] = atomicAdd(p.next_row, 1);
    __syncthreads();
    const Tables tb = make_tables<M>(stab);
    double* buf = smem + nlin_fft_tab_pad<M>() + (size_t)w * nlin_fft_worker_doubles<M, DFX>();
    C tw[Cfg<M>::RD];
    load_tw<M>(t, tb, tw);
    const int stride = gridDim.x * NW;
    for (;;) {
        const int row = s_row[w];
        if (row >= p.nrows) break;
        // the next claim travels to L2 and back while this row is transformed
        int next = 0;
        if (t == 0) next = atomicAdd(p.next_row, 1);
        const double* r0 = (MODE == 2 ? p.coef1 : p.coef0) + (size_t)row * 7 * K;
        if (DFX) bc_inv_dfx<M, NT>(t, r0, p.coef1 + (size_t)row * 7 * K, buf, tb);
        else bc_inv_fx<M, NT>(t, r0, buf, tb);
        // pull the row that will be claimed one round from now from HBM into L2 while this one is transformed
        if (row + stride < p.nrows) {
            const char* nx = reinterpret_cast<const char*>((MODE == 2 ? p.coef1 : p.coef0) + (size_t)(row + stride) * 7 * K);
            for (int o = t * 128; o < 7 * K * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
            if (MODE == 2) {
                const char* ng = reinterpret_cast<const char*>(p.grid + (size_t)(row + stride) * 7 * M);
                for (int o = t * 128; o < 7 * M * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ng + o));
            }
            if (DFX) {
                const char* nx1 = reinterpret_cast<const char*>(p.coef1 + (size_t)(row + stride) * 7 * K);
                for (int o = t * 128; o < 7 * K * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx1 + o));
            }
        }
        worker_sync<NT>(w);
        pass_d<M, NF, +1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        if (MODE == 1) {
            i3f1_grid<M, NT>(t, buf, tb, p.grid + (size_t)row * 7 * M);
            if (t == 0) s_row[w] = next;
            worker_sync<NT>(w);
            continue;
        }
        if (DFX) i3f1_dfx<M, NT>(t, buf, tb);
        else if (MODE == 2) i3f1_jvpc<M, NT>(t, buf, tb, p.grid + (size_t)row * 7 * M);
        else i3f1_fx<M, NT>(t, buf, tb);
        worker_sync<NT>(w);
        pass_d<M, 2, -1, NT>(t, buf, tw);
        worker_sync<NT>(w);
        cp_fwd<M, NT>(t, buf, p.spec + (size_t)row * spec_pitch(K), tb);
        if (t == 0) s_row[w] = next;
        worker_sync<NT>(w);   // the planes are free again and the next row index is visible
    }
    rows_done(p.next_row);
}

// ---- staged one-state kernel (64-thread workers: M = 384) -----------------------------------------------------------------
// Same phases, two changes in the schedule:
//  * Ownership.  Warp w of a worker owns the transforms w and w + 2 on the way in and transform w on the way back: packing,
//    radix-8 and radix-RD passes of a transform never leave their warp, so only __syncwarp() separates them.  The two
//    warps meet twice per row -- before and after the radix-6 / product phase, which needs every transform at one grid
//    column -- instead of six times, and drift freely in between.
//  * Staging.  The coefficient rows of the NEXT row arrive by TMA bulk copies (cp.async.bulk + mbarrier, one barrier per
//    warp) while the current row is still being analysed: rows (JT, omega) and (DT, Dpsi) land in the plane pairs 2 and 3,
//    which are dead after the product phase, rows (DS, -kT, -kS) in 6 KB of their own.  Warp w first packs its transform
//    w from the dead planes of pair 2 + w, then overwrites exactly those planes with its transform 2 + w: no global
//    load is left in the packing phase (it was 35 % of all stall samples, most of them on the L2 latency).
#ifndef NLIN_FFT_STAGED_NW
#define NLIN_FFT_STAGED_NW 6  // 30 KB per worker at M = 384; 6 workers = 3 warps per scheduler at 168 registers (no spills) measured faster than 7
#endif
template <int M>
__host__ __device__ constexpr size_t nlin_fft_staged_worker_doubles() { return (size_t)8 * fftp::Cfg<M>::PL + 3 * fftp::Cfg<M>::K; }
template <int M>
__host__ __device__ constexpr size_t nlin_fft_staged_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)nlin_fft_tab_pad<M>() + (size_t)nw * nlin_fft_staged_worker_doubles<M>());
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int M, int NW, int MODE = 0>
__global__ void __launch_bounds__(64 * NW, 1) nlin_fft_staged_kernel(NlinFftParams p) {
    using namespace fftp;
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NT = 64;
    static_assert(Cfg<M>::L == NT, "one radix-6 column per thread");
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_row[NW];
    __shared__ __align__(8) uint64_t s_bar[NW][2];
    double* stab = smem;
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < tab_doubles<M>(); i += NT * NW) stab[i] = p.tab[i];
    const int w = threadIdx.x / NT, t = threadIdx.x % NT, warp = t >> 5, lane = t & 31;
    if (lane == 0) mbar_init(&s_bar[w][warp], 1);
    pdl_wait();   // coefficient rows (prep_kernel) and the row counter (reset by the previous launch's last CTA)
    if (t == 0) s_row[w] = atomicAdd(p.next_row, 1);
    mbar_fence_init();
    __syncthreads();
    const Tables tb = make_tables<M>(stab);
    double* buf = smem + nlin_fft_tab_pad<M>() + (size_t)w * nlin_fft_staged_worker_doubles<M>();
    double* extra = buf + 8 * PL;
    C tw[Cfg<M>::RD];
    load_tw<M>(t, tb, tw);
    // lane 0 of a warp fetches what its warp packs: rows (2 warp, 2 warp + 1) and its share of (DS, -kT | -kS)
    auto fetch = [&](int r) {
        if (lane != 0 || r >= p.nrows) return;
        const double* src = (MODE == 2 ? p.coef1 : p.coef0) + (size_t)r * 7 * K;
        uint64_t* bar = &s_bar[w][warp];
        mbar_expect_tx(bar, (warp == 0 ? 4 : 3) * K * (unsigned)sizeof(double));
        bulk_g2s(buf + (4 + 2 * warp) * PL, src + 2 * warp * K, 2 * K * sizeof(double), bar);
        if (warp == 0) bulk_g2s(extra, src + 4 * K, 2 * K * sizeof(double), bar);
        else bulk_g2s(extra + 2 * K, src + 6 * K, K * sizeof(double), bar);
    };
    int row = s_row[w];
    fetch(row);
    unsigned ph = 0;
    while (row < p.nrows) {
        int next = 0;
        if (t == 0) next = atomicAdd(p.next_row, 1);   // travels to L2 and back while this row is transformed
        mbar_wait(&s_bar[w][warp], ph);
        ph ^= 1;
        staged_pack<M>(warp, lane, buf, tb, tw, 0);
        __syncwarp();
        staged_pack<M>(warp, lane, buf, tb, tw, 1);
        __syncwarp();
        staged_pack<M>(warp, lane, buf, tb, tw, 2);
        if (t == 0) s_row[w] = next;
        worker_sync<NT>(w);
        const int nrow = s_row[w];
        if (MODE == 1) i3f1_grid<M, NT>(t, buf, tb, p.grid + (size_t)row * 7 * M);
        else if (MODE == 2) i3f1_jvpc<M, NT>(t, buf, tb, p.grid + (size_t)row * 7 * M);
        else i3f1_fx<M, NT>(t, buf, tb);
        worker_sync<NT>(w);
        fence_proxy_async();   // the generic-proxy reads of the dead planes precede the bulk copies into them
        fetch(nrow);
        if (MODE == 1) { row = nrow; continue; }
        if (MODE == 2 && nrow < p.nrows) {   // the next row's cached grid fields on their way into L2 while this row is analysed
            const char* ng = reinterpret_cast<const char*>(p.grid + (size_t)nrow * 7 * M);
            for (int o = t * 128; o < 7 * M * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ng + o));
        }
        staged_unpack<M>(warp, lane, buf, p.spec + (size_t)row * spec_pitch(K), tb, tw, 0);
        __syncwarp();
        staged_unpack<M>(warp, lane, buf, p.spec + (size_t)row * spec_pitch(K), tb, tw, 1);
        __syncwarp();
        row = nrow;
    }
    rows_done(p.next_row);
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
// per worker: one plane pair + the two zero-padded coefficient rows (fftp::ke_stage)
template <int M>
__host__ __device__ constexpr int ke_fft_worker_doubles() { return 2 * fftp::Cfg<M>::PL + 2 * fftp::Cfg<M>::K; }
template <int M>
__host__ __device__ constexpr size_t ke_fft_smem_bytes(int nw) {
    return sizeof(double) * ((size_t)ke_fft_shared_doubles<M>() + (size_t)nw * ke_fft_worker_doubles<M>());
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
    double* buf = smem + ke_fft_shared_doubles<M>() + (size_t)w * ke_fft_worker_doubles<M>();
    double* srow = buf + 2 * PL;
    constexpr bool TWO_STEP = Cfg<M>::RD == 32;   // M = 1536: radix-32 middle pass in two in-place steps (fft_core.h)
    C tw[TWO_STEP ? 1 : Cfg<M>::RD];
    if constexpr (!TWO_STEP) load_tw<M>(t, tb, tw);
    for (int row = blockIdx.x * NW + w; row < p.nrows; row += gridDim.x * NW) {
        const double* r = p.rows + (size_t)row * p.row_stride;
        ke_stage<M>(t, r, r + p.b_off, p.ascale ? p.ascale[row % p.n] : 1.0, srow);
        worker_sync(w);
        ke_pack<M>(t, srow, buf, tb);
        worker_sync(w);
        if constexpr (TWO_STEP) {
            pass_d32_a<M>(t, buf, tb);
            worker_sync(w);
            pass_d32_b<M>(t, buf);
        } else {
            pass_d<M, 1, +1>(t, buf, tw);
        }
        worker_sync(w);
        const double part = warp_sum(ke6<M>(t, buf, tb, sW));
        if ((t & 31) == 0) s_part[w][t >> 5] = part;
        worker_sync(w);
        if (t == 0) p.kepart[row] = p.wr[row % p.n] * (s_part[w][0] + s_part[w][1]);
        worker_sync(w);
    }
}

// ---- direct-summation row kernel ---------------------------------------------------------------------------------------
// The same contract as nlin_fft_kernel -- [rows][7][K] spectral rows in, [rows][4][K] analysed products out, then
// post_kernel -- for shapes the FFT and the mirror-split DMMA transforms do not cover: N_fm = 2 (mod 4), whose 3/2-padded
// grid has an odd number of points (the reference only asks for an even N_fm, Matrix_Operators.py:758), and the two-state
// products at N_r > 41 outside N_fm = 128 / 256 / 512.  One CTA per (member, radial row); thread j synthesises the seven
// fields at theta_j by direct sums (exact argument reduction, no tables), forms the products of
// Matrix_Operators.py:791-793 (884-887 for a pair of states), thread k analyses them.  O(K M) per field: a correctness
// path for small or odd shapes, not a fast one.
struct NlinDirectParams {
    const double* coef0;  // [rows][7][K]
    const double* coef1;  // [rows][7][K] second state, or null
    double* spec;         // [rows][4][K]
    int K, M;
};

__device__ __forceinline__ void direct_fields(const double* __restrict__ r, int K, int M, int j, double (&f)[7]) {
    // rows: 0 JT (cos), 1 omega (sin), 2 DT (cos), 3 Dpsi (sin), 4 DS (cos), 5 -kT (sin), 6 -kS (sin)
#pragma unroll
    for (int q = 0; q < 7; ++q) f[q] = 0.0;
    for (int k = 0; k < K; ++k) {
        double c, s;
        trig_kj(k, j, M, c, s);
        f[0] = fma(r[k], c, f[0]);
        f[2] = fma(r[2 * K + k], c, f[2]);
        f[4] = fma(r[4 * K + k], c, f[4]);
        if (k) {   // entry 0 of a sine-type row is ignored (Transforms.py:41-54)
            f[1] = fma(r[K + k], s, f[1]);
            f[3] = fma(r[3 * K + k], s, f[3]);
            f[5] = fma(r[5 * K + k], s, f[5]);
            f[6] = fma(r[6 * K + k], s, f[6]);
        }
    }
}

__global__ void __launch_bounds__(128) nlin_direct_kernel(NlinDirectParams p) {
    extern __shared__ __align__(128) double smem[];
    const int K = p.K, M = p.M, row = blockIdx.x, tid = threadIdx.x;
    const bool two = p.coef1 != nullptr;
    double* sc0 = smem;                     // [7][K]
    double* sc1 = sc0 + 7 * K;              // [7][K] (two-state)
    double* sp = sc1 + (two ? 7 * K : 0);   // [4][M] grid products: JT*om, Dpsi*om, N_T, N_S
    for (int i = tid; i < 7 * K; i += 128) {
        sc0[i] = p.coef0[(size_t)row * 7 * K + i];
        if (two) sc1[i] = p.coef1[(size_t)row * 7 * K + i];
    }
    __syncthreads();
    for (int j = tid; j < M; j += 128) {
        double a[7];
        direct_fields(sc0, K, M, j, a);
        if (!two) {
            sp[j] = a[0] * a[1];
            sp[M + j] = a[3] * a[1];
            sp[2 * M + j] = a[0] * a[2] - a[3] * a[5];     // JT DT - Dpsi kT: the rows -k T, -k S ARE the sine series kT, kS
            sp[3 * M + j] = a[0] * a[4] - a[3] * a[6];
        } else {
            double b[7];
            direct_fields(sc1, K, M, j, b);
            sp[j] = a[0] * b[1] + b[0] * a[1];
            sp[M + j] = a[3] * b[1] + b[3] * a[1];
            sp[2 * M + j] = (a[0] * b[2] + b[0] * a[2]) - (a[3] * b[5] + b[3] * a[5]);
            sp[3 * M + j] = (a[0] * b[4] + b[0] * a[4]) - (a[3] * b[6] + b[3] * a[6]);
        }
    }
    __syncthreads();
    // out: [0] DST(JT om), [1] DST(kDpsi om + Dpsi kom) = -k DCT(Dpsi om), [2] DCT(N_T), [3] DCT(N_S)  (as cp_emit)
    double* o = p.spec + (size_t)row * spec_pitch(K);
    for (int k = tid; k < K; k += 128) {
        double s0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        for (int j = 0; j < M; ++j) {
            double c, s;
            trig_kj(k, j, M, c, s);
            s0 = fma(sp[j], s, s0);
            c1 = fma(sp[M + j], c, c1);
            c2 = fma(sp[2 * M + j], c, c2);
            c3 = fma(sp[3 * M + j], c, c3);
        }
        const double sc = (k == 0 ? 1.0 : 2.0) / M;
        const int pk = fftp::spec_pos(k, K);   // parity-split row, as the FFT kernels write it
        o[pk] = k ? s0 * (2.0 / M) : 0.0;
        o[K + pk] = -(double)k * (c1 * sc);
        o[2 * K + pk] = c2 * sc;
        o[3 * K + pk] = c3 * sc;
    }
}

struct PostParams {
    const double* spec;  // [B][n][4][K], rows parity-split (spec_pos)
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
    pdl_launch_dependents();
    for (int idx = tid; idx < n * n8; idx += 256) sD[idx] = p.DrT[idx];
    pdl_wait();   // the analysed products come from the row kernel
    auto issue = [&](int tile, int stage) {
        const int b = tile / nkt, k0 = (tile - b * nkt) * POST_TC;
        const double* sb = p.spec + (size_t)b * n * spec_pitch(K);
        double* sT = smem + stage * TS;
        for (int idx = tid; idx < 4 * n * POST_TC; idx += 256) {
            const int c = idx & (POST_TC - 1), fi = idx / POST_TC, f = fi & 3, i = fi >> 2;
            const bool ok = k0 + c < K;
            cp_async8_zfill(&sT[(f * n + i) * LDT + c], ok ? sb + (size_t)i * spec_pitch(K) + f * K + fftp::spec_pos(k0 + c, K) : sb, ok);
        }
    };
    const bool sm = p.bstride != 0;
    const int LDG = n8 + SDDC_SM_PAD;
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
