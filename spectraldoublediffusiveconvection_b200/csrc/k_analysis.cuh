// Stage 3: latitudinal analysis (scaled DCT-II / DST-II, truncated to K modes) as a parity-split DMMA GEMM of the
// folded grid products against the (scaled) cosine/sine tables, fused with the un-shift of the sine coefficients,
// the equatorial-symmetry mask, the -dt factor and the linear right-hand side.
//
// Reference semantics: NLIN_FX step 4-5 (Matrix_Operators.py:797-804), Step_Python (Main.py:262,271,276,280).
//     F_hat[(f,i), k=2k'+p] = sum_{j'} PRD[f][p][i][j'] * TAB2[type(f)][p][k'][j']
#pragma once
#include "common.cuh"

namespace sddc {

struct AnaParams {
    const double* prd;   // [B][3][2][n8][Mhp]
    const double* tab2;  // [2 types][2 par][Khp2][Mhp], scale factors folded in
    const double* lin;   // [B][3N] linear right-hand side (mode RHS) or null (mode F only)
    double* out;         // [B][3N]
    Geo g;
    double mdt;          // -dt (RHS mode)
};

constexpr int ANA_KC = 8;

template <int NT3>
__host__ __device__ inline size_t ana_stage_doubles(int rows3) {
    return (size_t)(ANA_KC / 4) * rows3 * 4 + (size_t)(ANA_KC / 4) * 2 * (NT3 * 64) * 4;
}

// grid = (Khp2 / (64*NT3), 2 parities, B), block = 256; warp w owns column tiles [w*NT3, (w+1)*NT3) and all
// 3*nt8 row tiles (psi rows use the sine table, T and S rows the cosine table).
template <int NT3, int MT3>
__global__ void __launch_bounds__(256) analysis_kernel(AnaParams p, int nstage) {
    constexpr int KT3 = NT3 * 64, KS = ANA_KC / 4;
    extern __shared__ __align__(16) double smem[];
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int kt = blockIdx.x, par = blockIdx.y, b = blockIdx.z;
    const int n = g.n, n8 = g.n8, K = g.K, N = g.N, Mhp = g.Mhp;
    const int rows3 = 3 * n8, TM3 = rows3 >> 3;
    double* outb = p.out + (long long)b * 3 * N;
    const double* linb = p.lin ? p.lin + (long long)b * 3 * N : nullptr;

    if (g.symmetric && par == 1) {
        // every odd-k output block is masked (Vecs_to_X symmetric branch, Matrix_Operators.py:536-556)
        for (int idx = tid; idx < KT3 * n; idx += 256) {
            const int kp = kt * KT3 + idx / n, i = idx % n;
            if (kp >= g.Kh) continue;
            const int k = 2 * kp + 1;
            outb[(long long)(k - 1) * n + i] = 0.0;
            outb[(long long)N + (long long)k * n + i] = 0.0;
            outb[2LL * N + (long long)k * n + i] = 0.0;
        }
        return;
    }

    const double* A = p.prd + (long long)b * 3 * 2 * n8 * Mhp;
    const long long fps = (long long)n8 * Mhp;  // parity stride; field stride = 2*fps
    const long long t2s = (long long)g.Khp2 * Mhp;
    const int A_ST = KS * rows3 * 4;
    const int STAGE = A_ST + KS * 2 * KT3 * 4;
    const int nchunk = Mhp / ANA_KC;

    double acc[MT3][NT3][2];
#pragma unroll
    for (int mt = 0; mt < MT3; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    auto load_stage = [&](int st, int chunk) {
        double* sA = smem + (size_t)st * STAGE;
        double* sB = sA + A_ST;
        const int j0 = chunk * ANA_KC;
        for (int idx = tid; idx < rows3 * 4; idx += 256) {
            const int piece = idx & 3, row = idx >> 2;
            const int f = row / n8, i = row - f * n8;
            const double* src = A + (long long)(f * 2 + par) * fps + (long long)i * Mhp + j0 + piece * 2;
            cp_async16(sA + ((piece >> 1) * rows3 + row) * 4 + (piece & 1) * 2, src);
        }
        for (int idx = tid; idx < 2 * KT3 * 4; idx += 256) {
            const int piece = idx & 3, r = idx >> 2;
            const int col = r % KT3, ty = r / KT3;
            const double* src = p.tab2 + (long long)(ty * 2 + par) * t2s + (long long)(kt * KT3 + col) * Mhp + j0 + piece * 2;
            cp_async16(sB + (((piece >> 1) * 2 + ty) * KT3 + col) * 4 + (piece & 1) * 2, src);
        }
    };

    for (int s = 0; s < nstage - 1; ++s) {
        if (s < nchunk) load_stage(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunk; ++c) {
        cp_async_wait_dyn(nstage - 2);
        __syncthreads();
        {
            const int cn = c + nstage - 1;
            if (cn < nchunk) load_stage(cn % nstage, cn);
            cp_async_commit();
        }
        const double* sA = smem + (size_t)(c % nstage) * STAGE;
        const double* sB = sA + A_ST;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            double bf[2][NT3];
#pragma unroll
            for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                for (int nt = 0; nt < NT3; ++nt)
                    bf[ty][nt] = sB[((ks * 2 + ty) * KT3 + (warp * NT3 + nt) * 8 + gq) * 4 + tq];
            const double* sAk = sA + (ks * rows3 + gq) * 4 + tq;
#pragma unroll
            for (int mt = 0; mt < MT3; ++mt) {
                if (mt < TM3) {
                    const double a = sAk[mt * 32];
                    const bool sn = mt < g.nt8;  // psi rows: sine table
#pragma unroll
                    for (int nt = 0; nt < NT3; ++nt)
                        mma884(acc[mt][nt][0], acc[mt][nt][1], a, sn ? bf[1][nt] : bf[0][nt]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: C fragment (row = (f,i), col = k') -> state layout [f][block][i] ----
#pragma unroll
    for (int mt = 0; mt < MT3; ++mt) {
        if (mt < TM3) {
            const int row = mt * 8 + gq;
            const int f = row / n8, i = row - f * n8;
            if (i < n) {
#pragma unroll
                for (int nt = 0; nt < NT3; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kp = kt * KT3 + (warp * NT3 + nt) * 8 + 2 * tq + e;
                        if (kp >= g.Kh) continue;
                        const int k = 2 * kp + par;
                        const double v = acc[mt][nt][e];
                        if (f == 0) {
                            // sinusoid index k -> code block k-1; k = 0 is dropped and block K-1 gets no
                            // nonlinear contribution (Matrix_Operators.py:802)
                            const int blk = (k >= 1) ? k - 1 : K - 1;
                            const double fv = (k >= 1) ? v : 0.0;
                            const long long o = (long long)blk * n + i;
                            const bool keep = !(g.symmetric && (blk & 1) == 0);
                            double r = keep ? fv : 0.0;
                            if (linb) r = keep ? fma(p.mdt, fv, linb[o]) : 0.0;
                            outb[o] = r;
                        } else {
                            const long long o = (long long)f * N + (long long)k * n + i;
                            double r = v;
                            if (linb) r = fma(p.mdt, v, linb[o]);
                            outb[o] = r;  // odd k never reaches here when symmetric
                        }
                    }
                }
            }
        }
    }
}

}  // namespace sddc
