// Batched Arnoldi orthogonalisation for the lock-step Newton / pseudo-arc-length drivers (krylov.py; the role SciPy's
// LGMRES plays inside Main._Newton / _ContinC, Main.py:523-539, 885-953): every member b owns a Krylov basis
// V[b][0..nvec) of vectors of length n, and a new direction w[b] is orthogonalised against it by classical
// Gram-Schmidt applied twice.  HBM-bound (the basis is read, nothing is reused): the passes are organised so that the
// basis crosses HBM three times per Arnoldi step instead of the four of  h = V^T w; w -= V h  done twice --
//   gs_dots_kernel          part1 = V^T w                          (pass 1)
//   gs_update_kernel<true>  w -= V h1;  part2 = V^T w (updated)    (pass 2; the second read of a chunk comes from L2)
//   gs_update_kernel<false> w -= V h2;  |w|^2 partial sums         (pass 3)
// Reductions over the vector length are two-level with a FIXED order (per-chunk partial sums written to global
// memory, summed in chunk order by the consumer): results are bit-reproducible from run to run.
#pragma once
#include "common.cuh"

namespace sddc {

constexpr int GS_CHUNK = 1024;   // vector elements per CTA
constexpr int GS_THREADS = 256;

struct GsParams {
    const double* V;          // [B][ldv][n]
    long long member_stride;  // doubles between the bases of consecutive members (ldv * n)
    int n, nvec;              // vector length, basis vectors in use
    double* w;                // [B][n]
    const double* part_in;    // [B][nchunk][ldp]: per-chunk partial dot products to apply (update kernel)
    double* part_out;         // [B][nchunk][ldp]: per-chunk partial dot products produced; slot nvec = |w_chunk|^2
    double* h_out;            // [B][ldp]: the summed coefficients this pass subtracts (column of the Hessenberg matrix)
    int nchunk, ldp;
};

// dot products of the CTA's chunk of w (in shared memory) with the same chunk of every basis vector: warp per vector
__device__ __forceinline__ void gs_chunk_dots(const GsParams& p, const double* Vb, const double* ws, int x0, int len,
                                              double* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < p.nvec; i += GS_THREADS / 32) {
        const double* row = Vb + (size_t)i * p.n + x0;
        double acc = 0.0;
#pragma unroll 4
        for (int x = lane; x < len; x += 32) acc = fma(row[x], ws[x], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[i] = acc;
    }
}

__global__ void __launch_bounds__(GS_THREADS) gs_dots_kernel(GsParams p) {
    __shared__ double ws[GS_CHUNK];
    const int b = blockIdx.y, c = blockIdx.x, x0 = c * GS_CHUNK, len = min(GS_CHUNK, p.n - x0);
    const double* wb = p.w + (size_t)b * p.n + x0;
    for (int x = threadIdx.x; x < len; x += GS_THREADS) ws[x] = wb[x];
    __syncthreads();
    gs_chunk_dots(p, p.V + (size_t)b * p.member_stride, ws, x0, len, p.part_out + ((size_t)b * p.nchunk + c) * p.ldp);
}

template <bool DOTS>
__global__ void __launch_bounds__(GS_THREADS) gs_update_kernel(GsParams p) {
    extern __shared__ double sm[];
    double* ws = sm;                 // [GS_CHUNK]
    double* hs = sm + GS_CHUNK;      // [nvec]
    __shared__ double red[32];
    const int b = blockIdx.y, c = blockIdx.x, x0 = c * GS_CHUNK, len = min(GS_CHUNK, p.n - x0);
    const double* Vb = p.V + (size_t)b * p.member_stride;
    // coefficients: partial sums of the previous pass, added in chunk order
    for (int i = threadIdx.x; i < p.nvec; i += GS_THREADS) {
        const double* pi = p.part_in + (size_t)b * p.nchunk * p.ldp + i;
        double s = 0.0;
        for (int cc = 0; cc < p.nchunk; ++cc) s += pi[(size_t)cc * p.ldp];
        hs[i] = s;
        if (c == 0) p.h_out[(size_t)b * p.ldp + i] = s;
    }
    __syncthreads();
    double* wb = p.w + (size_t)b * p.n + x0;
    double nrm = 0.0;
    for (int x = threadIdx.x; x < len; x += GS_THREADS) {
        const double* col = Vb + x0 + x;
        double a0 = wb[x], a1 = 0.0;
        int i = 0;
        for (; i + 1 < p.nvec; i += 2) {      // two independent chains
            a0 = fma(-hs[i], col[(size_t)i * p.n], a0);
            a1 = fma(-hs[i + 1], col[(size_t)(i + 1) * p.n], a1);
        }
        if (i < p.nvec) a0 = fma(-hs[i], col[(size_t)i * p.n], a0);
        const double v = a0 + a1;
        wb[x] = v;
        ws[x] = v;
        nrm = fma(v, v, nrm);
    }
    double* po = p.part_out + ((size_t)b * p.nchunk + c) * p.ldp;
    nrm = block_sum(nrm, red);       // contains the barrier that publishes ws
    if (threadIdx.x == 0) po[p.nvec] = nrm;
    if (DOTS) gs_chunk_dots(p, Vb, ws, x0, len, po);
}

}  // namespace sddc
