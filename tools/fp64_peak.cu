// Micro-benchmark: measured fp64 peaks on this B200 (DFMA pipe, DMMA m8n8k4 / m16n8k8 / m16n8k16).
// Used only to obtain the fp64 roofline denominator that MEASURED_PEAKS.json lacks.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// dependent-issue latency of DMMA.8x8x4: one warp, one accumulator chain
__global__ void dmma_latency_kernel(double* out, long long* cycles, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (c0 + c1 == 123.456) out[0] = c0;
}

// DMMA fed from shared memory the way the GEMM kernels do it: per k-step MT A-fragments + NT B-fragments (LDS.64,
// conflict-free [row][4] layout), MT*NT independent accumulator tiles.  No global traffic, no barriers.
template <int MT, int NT>
__global__ void __launch_bounds__(640, 1) dmma_smem_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * (i & 127);
    __syncthreads();
    double acc[MT][NT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    const double* base = sm + (warp * 128 + g * 4 + t);
    for (int it = 0; it < iters; ++it) {
        const double* p = base + (it & 7) * 1024;
        double af[MT], bf[NT];
#pragma unroll
        for (int m = 0; m < MT; ++m) af[m] = p[m * 32];
#pragma unroll
        for (int n = 0; n < NT; ++n) bf[n] = p[4096 + n * 32];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[m][n][0]), "+d"(acc[m][n][1]) : "d"(af[m]), "d"(bf[n]));
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) s += acc[m][n][0] + acc[m][n][1];
    if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    double* out; CK(cudaMalloc(&out, 1024));
    const int sms = p.multiProcessorCount;
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        {
            float ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * (double)iters * 256.0 * grid;
            printf("DFMA      ilp16 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 8 * 8 * 4 * 8 * (double)iters * 8.0 * grid;
            printf("DMMA 8x8x4  ilp8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma1688_kernel<8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 8 * 8 * (double)(iters / 4) * 8.0 * grid;
            printf("DMMA 16x8x8 ilp8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma16816_kernel<8><<<grid, 256>>>(out, iters / 8, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 16 * 8 * (double)(iters / 8) * 8.0 * grid;
            printf("DMMA 16x8x16 ilp8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
    }
    {
        long long* cyc; CK(cudaMalloc(&cyc, 8));
        dmma_latency_kernel<<<1, 32>>>(out, cyc, 4096, 1.0000001, 1e-9);
        long long h = 0; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DMMA 8x8x4 dependent-chain latency: %.1f cycles/instr\n", (double)h / 4096.0);
    }
    {
        cudaFuncSetAttribute(dmma_smem_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(dmma_smem_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        const int it2 = 20000;
        float ms = time_ms([&] { dmma_smem_kernel<4, 4><<<sms, 576, 160 * 1024>>>(out, it2); }, 5);
        printf("DMMA from smem 4x4 tiles/warp, 18 warps/SM: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 256 * 16 * it2 * 18.0 * sms / ms * 1e-9);
        ms = time_ms([&] { dmma_smem_kernel<4, 2><<<sms, 576, 160 * 1024>>>(out, it2); }, 5);
        printf("DMMA from smem 4x2 tiles/warp, 18 warps/SM: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 256 * 8 * it2 * 18.0 * sms / ms * 1e-9);
        ms = time_ms([&] { dmma_smem_kernel<4, 4><<<sms, 256, 160 * 1024>>>(out, it2); }, 5);
        printf("DMMA from smem 4x4 tiles/warp, 8 warps/SM: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 256 * 16 * it2 * 8.0 * sms / ms * 1e-9);
        cudaFuncSetAttribute(dmma_smem_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(dmma_smem_kernel<9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        for (int nw : {4, 8, 12, 16, 20}) {
            ms = time_ms([&] { dmma_smem_kernel<6, 2><<<sms, 32 * nw, 160 * 1024>>>(out, it2); }, 3);
            printf("DMMA from smem 6x2 tiles/warp, %2d warps/SM: %.2f TFLOP/s\n", nw, 2.0 * 256 * 12 * it2 * nw * sms / ms * 1e-9);
            ms = time_ms([&] { dmma_smem_kernel<9, 2><<<sms, 32 * nw, 160 * 1024>>>(out, it2); }, 3);
            printf("DMMA from smem 9x2 tiles/warp, %2d warps/SM: %.2f TFLOP/s\n", nw, 2.0 * 256 * 18 * it2 * nw * sms / ms * 1e-9);
            ms = time_ms([&] { dmma_smem_kernel<4, 4><<<sms, 32 * nw, 160 * 1024>>>(out, it2); }, 3);
            printf("DMMA from smem 4x4 tiles/warp, %2d warps/SM: %.2f TFLOP/s\n", nw, 2.0 * 256 * 16 * it2 * nw * sms / ms * 1e-9);
        }
    }
    // HBM copy
    size_t n = (size_t)1 << 27;  // 128 Mi double2 = 2 GiB
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
    CK(cudaMemset(a, 1, n * 16));
    float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
    printf("copy double2 2GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n * 16 / ms * 1e-6);
    CK(cudaDeviceSynchronize());
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
