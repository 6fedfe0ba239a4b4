// Shared device helpers: fp64 tensor-core MMA (DMMA m8n8k4), cp.async staging, reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#ifndef SDDC_SM_PAD
#define SDDC_SM_PAD 2   // solve-major rows: n8 + SDDC_SM_PAD doubles per member (k_solve.cuh)
#endif

namespace sddc {

// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment ownership (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
// 8-byte asynchronous copy with zero fill when !valid (nothing is read from gmem_src then)
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem_src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
    if (pending <= 0) cp_async_wait<0>();
    else if (pending == 1) cp_async_wait<1>();
    else if (pending == 2) cp_async_wait<2>();
    else cp_async_wait<3>();
}

// ---- mbarrier + bulk async copy (TMA 1-D) helpers -------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// non-blocking variant: mbarrier.test_wait polled in a loop (try_wait may suspend the warp for an implementation-defined
// time; a hand-off that is on the critical path wants the shortest reaction time instead)
__device__ __forceinline__ void mbar_spin(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_SPIN:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_SPIN;\n"
        "bra LAB_SPIN;\n"
        "DONE_SPIN:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// contiguous global -> shared copy by the TMA engine; completion is signalled on `bar` (complete_tx::bytes).
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of a contiguous global range by the TMA engine (no shared-memory destination, no completion signal)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// Programmatic dependent launch (the kernels of a member-step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel lets its successor's CTAs become resident as soon as all
// of its own have started, and itself touches nothing its predecessors produced before pdl_wait() -- the successor's
// prologue (tables / operators into shared memory, barrier initialisation) then overlaps the predecessor's tail.  Both
// are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed (deterministic) combination order. `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (warp == 0) {
        s = (lane < nw) ? red[lane] : 0.0;
        s = warp_sum(s);
    }
    return s;  // valid in warp 0
}

// Order of the wavenumbers inside one 8-wide chunk of a parity block of the synthesis operands.  Odd k keep their
// natural order.  Even k (k = 2k') are stored with the k' even ones first (MMA k-step 0) and the k' odd ones second
// (k-step 1), so that the two classes k = 0 and k = 2 (mod 4) can be contracted separately (second mirror level,
// k_synth_ws.cuh); kernels that contract a whole chunk are indifferent to the order as long as coefficients and
// tables agree.
__host__ __device__ __forceinline__ int chunk_pos(int kp, int par) {
    const int r = kp & 7;
    return (kp & ~7) | (par == 0 ? (((r & 1) << 2) | (r >> 1)) : r);
}
__host__ __device__ __forceinline__ int chunk_pos_inv(int pos, int par) {
    const int r = pos & 7;
    return (pos & ~7) | (par == 0 ? (((r & 3) << 1) | (r >> 2)) : r);
}

// Analysed products of the row kernels: spec[b * n + i] is a block of four parity-split rows [4][K] (fft_core.h: spec_pos),
// the blocks SPEC_PAD doubles apart.  Measured on B200 with pads of 4 and 20 doubles (so that the sectors a
// back-substitution chain gathers from the rows of its members do not sit a power of two apart): no difference for the
// gather, post_kernel 10 % slower -- the pad stays 0.
#ifndef SPEC_PAD
#define SPEC_PAD 0
#endif
__host__ __device__ constexpr long long spec_pitch(int K) { return 4LL * K + SPEC_PAD; }

// Geometry shared by all kernels.
struct Geo {
    int n;      // interior radial points
    int n8;     // n rounded up to a multiple of 8 (MMA row tiles)
    int nt8;    // n8 / 8
    int K;      // latitudinal modes N_fm
    int Kh;     // K / 2  (modes per parity)
    int Khp;    // Kh rounded up to a multiple of 8 (cp.async chunking)
    int Khp2;   // Kh rounded up to a multiple of the analysis column tile (128 or 64)
    int M;      // 3K/2 de-aliased grid
    int Mh;     // M / 2  (mirror pairs)
    int Mhp;    // Mh rounded up to a multiple of 32
    int N;      // n*K
    int symmetric;
    double dt, Pr, Tau;
};

}  // namespace sddc
