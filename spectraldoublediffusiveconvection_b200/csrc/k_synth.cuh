// Stage 2: fused latitudinal synthesis (scaled DCT-III / DST-III as dense, parity-split DMMA GEMMs against
// L2-resident cosine/sine tables), pointwise Jacobian products on the 3/2-padded grid, the radial derivative
// Dr@(JT*omega), and the mirror fold that prepares the analysis GEMM.
//
// Reference semantics: NLIN_FX steps 2-3 (Matrix_Operators.py:776-793), NLIN_DFX (842-887), and the grid part of
// Kinetic_Energy (Main.py:117-121).  Closed forms: SURVEY.md section 9.
//
// GEMM view (per member b, per parity p of the wavenumber k = 2k'+p):
//     EO_p[row, j'] = sum_{k'} COEF[row][p][k'] * TAB[type(row)][p][j'][k'],   j' < M/2
// and with theta_{M-1-j'} = pi - theta_{j'}:
//     cosine-type rows: f(j') = E+O, f(M-1-j') = E-O ;  sine-type rows: f(j') = O+E, f(M-1-j') = O-E.
#pragma once
#include "common.cuh"

namespace sddc {

enum { EPI_FX = 0, EPI_DFX = 1, EPI_KE = 2 };

struct SynthParams {
    const double* coef;      // [B][rows][2][Khp]
    long long coef_stride;   // member stride (doubles)
    const double* tab;       // [2 types][2 par][Mhp_tab][Khp]
    int tab_Mhp;             // j' extent of the table
    const double* Dr;        // [n][n] row-major (EPI_FX / EPI_DFX)
    double* prd;             // [B][3][2][n8][Mhp]            (EPI_FX / EPI_DFX)
    const double* wr;        // [n] radial trapezoid weights     (EPI_KE)
    const double* wth;       // [tab_Mhp] w_theta(j') sin(theta_j') (EPI_KE), zero padded
    double* kepart;          // [B][gridDim.x] partial sums      (EPI_KE)
    int rows;                // nfields * n8
    unsigned type_mask;      // bit f set: field f is sine-type
    Geo g;
};

template <int NT>
struct SynthCfg {
    static constexpr int W = NT * 8;   // mirror pairs (columns) per CTA
    static constexpr int KC = 8;       // k' per pipeline stage
    static constexpr int KS = KC / 4;  // MMA k-steps per stage
    static constexpr int LDE = (NT == 4) ? W + 8 : W;  // row stride of the E/O exchange buffer
};

template <int NT>
__host__ __device__ inline size_t synth_stage_doubles(int rows) {
    return (size_t)SynthCfg<NT>::KS * 2 * rows * 4 + (size_t)SynthCfg<NT>::KS * 4 * SynthCfg<NT>::W * 4;
}
template <int NT>
__host__ __device__ inline size_t synth_epi_doubles(int rows, int n) {
    return (size_t)2 * rows * SynthCfg<NT>::LDE + (size_t)2 * n * SynthCfg<NT>::W + (size_t)n * n + 32;
}

// grid = (Mhp/W, B), block = 256 (8 warps: warps 0-3 even-k GEMM, warps 4-7 odd-k GEMM; each warp owns a
// contiguous range of 8-row tiles and all NT column tiles).
template <int NT, int MTW, int EPI>
__global__ void __launch_bounds__(256, 1) synth_kernel(SynthParams p, int nstage) {
    using C = SynthCfg<NT>;
    constexpr int W = C::W, KS = C::KS, KC = C::KC, LDE = C::LDE;
    extern __shared__ __align__(16) double smem[];
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int par = warp >> 2, q = warp & 3;
    const int rows = p.rows, TM = rows >> 3, tpw = (TM + 3) >> 2;
    const int tile0 = q * tpw;
    const int ntiles = max(0, min(tpw, TM - tile0));
    const int b = blockIdx.y, jt = blockIdx.x;
    const int Khp = g.Khp;
    const double* A = p.coef + (long long)b * p.coef_stride;
    const double* Bt = p.tab + (long long)jt * W * Khp;
    const long long tab_ps = (long long)p.tab_Mhp * Khp;  // parity stride; type stride = 2*tab_ps

    const int A_ST = KS * 2 * rows * 4;
    const int STAGE = A_ST + KS * 4 * W * 4;
    const int nchunk = Khp / KC;

    // per-warp tile types (bit mt set: sine-type table)
    unsigned long long my_types = 0;  // up to MTW = 36 tiles
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
        const int fld = (tile0 + mt) / g.nt8;
        if (fld < 32) my_types |= (unsigned long long)((p.type_mask >> fld) & 1u) << mt;
    }

    double acc[MTW][NT][2];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    auto load_stage = [&](int st, int chunk) {
        double* sA = smem + (size_t)st * STAGE;
        double* sB = sA + A_ST;
        const int k0 = chunk * KC;
        // A: (row, parity) segments of KC doubles = 4 pieces of 16 B
        for (int idx = tid; idx < rows * 2 * 4; idx += 256) {
            const int piece = idx & 3, rp = idx >> 2;
            const int pr = rp & 1, row = rp >> 1;
            const double* src = A + ((long long)row * 2 + pr) * Khp + k0 + piece * 2;
            double* dst = sA + (((piece >> 1) * 2 + pr) * rows + row) * 4 + (piece & 1) * 2;
            cp_async16(dst, src);
        }
        for (int idx = tid; idx < 4 * W * 4; idx += 256) {
            const int piece = idx & 3, r = idx >> 2;
            const int col = r % W, tp = r / W;  // tp = type*2 + par
            const double* src = Bt + (long long)tp * tab_ps + (long long)col * Khp + k0 + piece * 2;
            double* dst = sB + (((piece >> 1) * 4 + tp) * W + col) * 4 + (piece & 1) * 2;
            cp_async16(dst, src);
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
            double bf[2][NT];
#pragma unroll
            for (int ty = 0; ty < 2; ++ty)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    bf[ty][nt] = sB[(((ks * 2 + ty) * 2 + par) * W + nt * 8 + gq) * 4 + tq];
            const double* sAk = sA + ((ks * 2 + par) * rows + tile0 * 8 + gq) * 4 + tq;
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                if (mt < ntiles) {
                    const double a = sAk[mt * 32];
                    const bool sn = (my_types >> mt) & 1ull;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        mma884(acc[mt][nt][0], acc[mt][nt][1], a, sn ? bf[1][nt] : bf[0][nt]);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- epilogue: exchange E/O through shared memory ----
    double* sEO = smem;                         // [2][rows][LDE]
    double* sA1 = sEO + (size_t)2 * rows * LDE; // [2 mirror][n][W]
    double* sDr = sA1 + (size_t)2 * g.n * W;    // [n][n]
    double* red = sDr + (size_t)g.n * g.n;      // [32]
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
        if (mt < ntiles) {
            const int row = (tile0 + mt) * 8 + gq;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<double2*>(&sEO[((size_t)par * rows + row) * LDE + nt * 8 + 2 * tq]) = v;
            }
        }
    }
    const int n = g.n, n8 = g.n8;
    if (EPI != EPI_KE) {
        for (int idx = tid; idx < n * n; idx += 256) sDr[idx] = p.Dr[idx];
    }
    __syncthreads();

    const double* E = sEO;
    const double* O = sEO + (size_t)rows * LDE;
    const int npts = n * W;

    if (EPI == EPI_KE) {
        // rows: field 0 = J_theta(psi)/r (cosine), field 1 = Dr psi (sine).  Main.py:117-130.
        double part = 0.0;
        for (int pt = tid; pt < npts; pt += 256) {
            const int i = pt / W, c = pt - i * W;
            const double e0 = E[(size_t)i * LDE + c], o0 = O[(size_t)i * LDE + c];
            const double e1 = E[(size_t)(n8 + i) * LDE + c], o1 = O[(size_t)(n8 + i) * LDE + c];
            const double ja = e0 + o0, jb = e0 - o0, da = o1 + e1, db = o1 - e1;
            part += p.wr[i] * p.wth[jt * W + c] * ((ja * ja + da * da) + (jb * jb + db * db));
        }
        const double s = block_sum(part, red);
        if (tid == 0) p.kepart[(long long)b * gridDim.x + jt] = s;
        return;
    }

    // field order (Derivatives): 0 JT, 1 kDpsi, 2 komega, 3 DT, 4 DS | 5 omega, 6 Dpsi, 7 kT, 8 kS
    constexpr int PTS = 4;
    double qv[PTS][2];
    double* prd = p.prd + (long long)b * 3 * 2 * n8 * g.Mhp;
    const long long pps = (long long)n8 * g.Mhp;  // parity stride in prd; field stride = 2*pps
#pragma unroll
    for (int s = 0; s < PTS; ++s) {
        const int pt = tid + s * 256;
        qv[s][0] = qv[s][1] = 0.0;
        if (pt < npts) {
            const int i = pt / W, c = pt - i * W;
            double f0[9], f1[9];  // values at j' and at the mirror point M-1-j'
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                const double e = E[(size_t)(a * n8 + i) * LDE + c], o = O[(size_t)(a * n8 + i) * LDE + c];
                if (a < 5) { f0[a] = e + o; f1[a] = e - o; } else { f0[a] = o + e; f1[a] = o - e; }
            }
            double a1_0, a1_1, nt0, nt1, ns0, ns1;
            if (EPI == EPI_FX) {
                a1_0 = f0[0] * f0[5];                     a1_1 = f1[0] * f1[5];
                qv[s][0] = f0[1] * f0[5] + f0[6] * f0[2]; qv[s][1] = f1[1] * f1[5] + f1[6] * f1[2];
                nt0 = f0[0] * f0[3] - f0[6] * f0[7];      nt1 = f1[0] * f1[3] - f1[6] * f1[7];
                ns0 = f0[0] * f0[4] - f0[6] * f0[8];      ns1 = f1[0] * f1[4] - f1[6] * f1[8];
            } else {
                double h0[9], h1[9];  // perturbation fields (second coefficient set)
#pragma unroll
                for (int a = 0; a < 9; ++a) {
                    const double e = E[(size_t)((9 + a) * n8 + i) * LDE + c], o = O[(size_t)((9 + a) * n8 + i) * LDE + c];
                    if (a < 5) { h0[a] = e + o; h1[a] = e - o; } else { h0[a] = o + e; h1[a] = o - e; }
                }
                a1_0 = f0[0] * h0[5] + h0[0] * f0[5];
                a1_1 = f1[0] * h1[5] + h1[0] * f1[5];
                qv[s][0] = (f0[1] * h0[5] + f0[6] * h0[2]) + (h0[1] * f0[5] + h0[6] * f0[2]);
                qv[s][1] = (f1[1] * h1[5] + f1[6] * h1[2]) + (h1[1] * f1[5] + h1[6] * f1[2]);
                nt0 = (h0[0] * f0[3] - h0[6] * f0[7]) + (f0[0] * h0[3] - f0[6] * h0[7]);
                nt1 = (h1[0] * f1[3] - h1[6] * f1[7]) + (f1[0] * h1[3] - f1[6] * h1[7]);
                ns0 = (h0[0] * f0[4] - h0[6] * f0[8]) + (f0[0] * h0[4] - f0[6] * h0[8]);
                ns1 = (h1[0] * f1[4] - h1[6] * f1[8]) + (f1[0] * h1[4] - f1[6] * h1[8]);
            }
            sA1[(size_t)i * W + c] = a1_0;
            sA1[(size_t)(n + i) * W + c] = a1_1;
            // cosine-type analysis (T, S): even k uses f(j')+f(mirror), odd k uses the difference
            const long long o = (long long)i * g.Mhp + jt * W + c;
            prd[(1 * 2 + 0) * pps + o] = nt0 + nt1;
            prd[(1 * 2 + 1) * pps + o] = nt0 - nt1;
            prd[(2 * 2 + 0) * pps + o] = ns0 + ns1;
            prd[(2 * 2 + 1) * pps + o] = ns0 - ns1;
        }
    }
    __syncthreads();
    // N_psi = Dr @ (JT*omega) - (kDpsi*omega + Dpsi*komega)   (Matrix_Operators.py:791)
#pragma unroll
    for (int s = 0; s < PTS; ++s) {
        const int pt = tid + s * 256;
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
            // sine-type analysis: odd k uses the sum, even k the difference
            const long long o = (long long)i * g.Mhp + jt * W + c;
            prd[(0 * 2 + 1) * pps + o] = v0 + v1;
            prd[(0 * 2 + 0) * pps + o] = v0 - v1;
        }
    }
}

}  // namespace sddc
