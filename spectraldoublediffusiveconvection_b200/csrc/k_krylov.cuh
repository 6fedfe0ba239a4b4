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
    const int* member_mask;   // optional [B]: members with 0 are skipped by the update kernel (their w stays as it is)
};

// dot products of the CTA's chunk of w (in shared memory) with the same chunk of every basis vector: a warp takes two
// vectors at a time (eight independent 8-byte loads per lane in flight)
__device__ __forceinline__ void gs_chunk_dots(const GsParams& p, const double* Vb, const double* ws, int x0, int len,
                                              double* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = GS_THREADS / 32;
    for (int i = 2 * warp; i < p.nvec; i += 2 * NW) {
        const bool two = i + 1 < p.nvec;
        const double* r0 = Vb + (size_t)i * p.n + x0;
        const double* r1 = two ? r0 + p.n : r0;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int x = lane;
        for (; x + 96 < len; x += 128) {
            const double w0 = ws[x], w1 = ws[x + 32], w2 = ws[x + 64], w3 = ws[x + 96];
            const double p0 = r0[x], p1 = r0[x + 32], p2 = r0[x + 64], p3 = r0[x + 96];
            const double q0 = r1[x], q1 = r1[x + 32], q2 = r1[x + 64], q3 = r1[x + 96];
            a0 = fma(p0, w0, a0); a1 = fma(p1, w1, a1); a0 = fma(p2, w2, a0); a1 = fma(p3, w3, a1);
            b0 = fma(q0, w0, b0); b1 = fma(q1, w1, b1); b0 = fma(q2, w2, b0); b1 = fma(q3, w3, b1);
        }
        for (; x < len; x += 32) {
            a0 = fma(r0[x], ws[x], a0);
            b0 = fma(r1[x], ws[x], b0);
        }
        const double sa = warp_sum(a0 + a1), sb = warp_sum(b0 + b1);
        if (lane == 0) {
            out[i] = sa;
            if (two) out[i + 1] = sb;
        }
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
    if (p.member_mask && p.member_mask[b] == 0) return;
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
    // thread t owns the elements t, t + 256, t + 512, t + 768 of the chunk: four independent accumulation chains, the
    // loads of one basis vector are coalesced across the warp
    {
        constexpr int NX = GS_CHUNK / GS_THREADS;
        double acc[NX];
        bool in[NX];
#pragma unroll
        for (int u = 0; u < NX; ++u) {
            const int x = threadIdx.x + u * GS_THREADS;
            in[u] = x < len;
            acc[u] = in[u] ? wb[x] : 0.0;
        }
        const double* col = Vb + x0 + threadIdx.x;
        for (int i = 0; i < p.nvec; ++i, col += p.n) {
            const double h = hs[i];
#pragma unroll
            for (int u = 0; u < NX; ++u)
                if (in[u]) acc[u] = fma(-h, col[u * GS_THREADS], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < NX; ++u) {
            const int x = threadIdx.x + u * GS_THREADS;
            if (in[u]) {
                wb[x] = acc[u];
                ws[x] = acc[u];
                nrm = fma(acc[u], acc[u], nrm);
            }
        }
    }
    double* po = p.part_out + ((size_t)b * p.nchunk + c) * p.ldp;
    nrm = block_sum(nrm, red);       // contains the barrier that publishes ws
    if (threadIdx.x == 0) po[p.nvec] = nrm;
    if (DOTS) gs_chunk_dots(p, Vb, ws, x0, len, po);
}

// ---- Hessenberg column of one Arnoldi step, all members (batched_gmres in krylov.py) -----------------------------------------
// One thread per member: takes the projection coefficients h[0..j] and the norm of the orthogonalised vector, applies
// the member's previous Givens rotations, forms the new one, updates the rotated right-hand side g and the residual
// estimate, and retires the member when it has reached its tolerance.  Masked members (live = 0) get an identity column
// and a zero right-hand-side entry, so that the triangular solve at the end of the cycle leaves their solution untouched.
// O(j) work per member per step: this replaces two dozen tiny tensor operations per Arnoldi step, not arithmetic.
struct GmresColParams {
    const double* h;      // [B][ldh] projection coefficients of this step (ldh >= j + 1)
    const double* hn;     // [B] norm of the orthogonalised vector
    double* H;            // [B][m+1][m] rotated Hessenberg matrix (upper triangular part is R)
    double* cs;           // [B][m]
    double* sn;           // [B][m]
    double* g;            // [B][m+1] rotated right-hand side
    double* resid;        // [B]
    const double* tol;    // [B]
    int* live;            // [B] in / out
    int* any_live;        // [1] out: OR of the new live flags (zeroed by the caller)
    int B, j, m, ldh, shifted;
};

__global__ void gmres_column_kernel(GmresColParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int j = p.j, m = p.m;
    double* Hb = p.H + (size_t)b * (m + 1) * m;
    double* gb = p.g + (size_t)b * (m + 1);
    if (!p.live[b]) {
        for (int i = 0; i < j; ++i) Hb[(size_t)i * m + j] = 0.0;
        Hb[(size_t)j * m + j] = 1.0;
        Hb[(size_t)(j + 1) * m + j] = 0.0;
        p.cs[(size_t)b * m + j] = 1.0;
        p.sn[(size_t)b * m + j] = 0.0;
        gb[j] = 0.0;
        gb[j + 1] = 0.0;
        return;
    }
    const double* hb = p.h + (size_t)b * p.ldh;
    const double* csb = p.cs + (size_t)b * m;
    const double* snb = p.sn + (size_t)b * m;
    // column after the previous rotations, written as it is produced (element i is final once rotation i has acted)
    double lo = hb[0] - ((p.shifted && j == 0) ? 1.0 : 0.0);
    for (int i = 0; i < j; ++i) {
        const double hi = hb[i + 1] - ((p.shifted && i + 1 == j) ? 1.0 : 0.0);
        const double c = csb[i], s = snb[i];
        Hb[(size_t)i * m + j] = c * lo + s * hi;
        lo = -s * lo + c * hi;
    }
    const double hnb = p.hn[b];
    const double den = sqrt(lo * lo + hnb * hnb);
    double c = 1.0, s = 0.0, diag = 1.0;
    if (den > 0.0) { c = lo / den; s = hnb / den; diag = c * lo + s * hnb; }
    Hb[(size_t)j * m + j] = diag;
    Hb[(size_t)(j + 1) * m + j] = 0.0;
    p.cs[(size_t)b * m + j] = c;
    p.sn[(size_t)b * m + j] = s;
    const double gj = gb[j];
    double gn = -s * gj;
    gb[j] = c * gj;
    const double r = fabs(gn);
    p.resid[b] = r;
    const bool still = r > p.tol[b];
    if (!still) gn = 0.0;              // a member that stops here
    gb[j + 1] = gn;
    p.live[b] = still ? 1 : 0;
    if (still) atomicOr(p.any_live, 1);
}

}  // namespace sddc
