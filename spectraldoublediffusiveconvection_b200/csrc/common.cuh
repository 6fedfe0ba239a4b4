// Shared device helpers: fp64 tensor-core MMA (DMMA m8n8k4), cp.async staging, reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

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

// Geometry shared by all kernels.
struct Geo {
    int n;      // interior radial points
    int n8;     // n rounded up to a multiple of 8 (MMA row tiles)
    int nt8;    // n8 / 8
    int K;      // latitudinal modes N_fm
    int Kh;     // K / 2  (modes per parity)
    int Khp;    // Kh rounded up to a multiple of 8 (cp.async chunking)
    int Khp2;   // Kh rounded up to a multiple of 128 (analysis column tiles)
    int M;      // 3K/2 de-aliased grid
    int Mh;     // M / 2  (mirror pairs)
    int Mhp;    // Mh rounded up to a multiple of 32
    int N;      // n*K
    int symmetric;
    double dt, Pr, Tau;
};

}  // namespace sddc
