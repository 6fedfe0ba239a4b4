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
//
// Operand staging: both operands live in global memory in *tile-major* order, so that everything one pipeline
// stage needs is one contiguous block per operand and is fetched by a single TMA bulk copy
// (cp.async.bulk ... mbarrier::complete_tx) issued by a dedicated producer warp:
//     COEF set : [b][chunk][ks][parity][row][4]      chunk = 8 consecutive k', ks = MMA k-step, 4 = MMA k
//     TAB      : [column tile][chunk][ks][type][parity][col][4]
// The [row][4] / [col][4] inner layout is exactly the m8n8k4 fragment order (bank-conflict free LDS.64).
#pragma once
#include "common.cuh"

namespace sddc {

enum { EPI_FX = 0, EPI_DFX = 1, EPI_KE = 2 };

struct SynthParams {
    const double* coef0;     // coefficient set of X      [B][Khp/8][2][2][RS][4]
    const double* coef1;     // coefficient set of dv (EPI_DFX)
    long long coef_stride;   // member stride (doubles)
    const double* tab;       // [Mhp_tab/W][Khp/8][2][2 types][2 par][W][4]
    const double* Dr;        // [n][n] row-major (EPI_FX / EPI_DFX)
    double* prd;             // [B][2 par][Mhp/8][2][3*n8][4]   (EPI_FX / EPI_DFX), tile-major for the analysis GEMM
    const double* wr;        // [n] radial trapezoid weights     (EPI_KE)
    const double* wth;       // [Mhp_tab] w_theta(j') sin(theta_j') (EPI_KE), zero padded
    double* kepart;          // [B][gridDim.x] partial sums      (EPI_KE)
    double* gridc;           // [B][9][2][n8][Mhp] cached base-state grid fields (k_synth_ws.cuh, JVP with cached base)
    Geo g;
};

constexpr int SYNTH_KC = 8;            // k' per pipeline stage
constexpr int SYNTH_KS = SYNTH_KC / 4; // MMA k-steps per stage
constexpr int SYNTH_MAX_STAGES = 4;

// column tiles per CTA for a warp that owns `mtw` 8-row tiles (accumulators: mtw*NT*2 doubles per thread)
__host__ __device__ constexpr int synth_nt_for(int mtw) { return mtw <= 4 ? 4 : (mtw <= 8 ? 2 : 1); }
__host__ __device__ constexpr int synth_lde(int nt) { return nt == 4 ? 40 : nt * 8; }  // E/O exchange row stride

// doubles per pipeline stage / for the epilogue, for `nset` coefficient sets of `rs` rows and NT column tiles
__host__ __device__ inline size_t synth_stage_doubles(int nset, int rs, int nt) {
    return (size_t)nset * SYNTH_KS * 2 * rs * 4 + (size_t)SYNTH_KS * 4 * (nt * 8) * 4;
}
__host__ __device__ inline size_t synth_epi_doubles(int nset, int rs, int n, int nt) {
    return (size_t)2 * nset * rs * synth_lde(nt) + (size_t)2 * n * (nt * 8) + (size_t)n * n + 32;
}

// grid = (Mhp/W, B).  Warps 0 .. 2*NF-1 are MMA consumers, one per (parity, field): the warp owns the NT8 8-row
// tiles of its field (in both coefficient sets for EPI_DFX) and all NT column tiles, so table type, tile count and
// loop bounds are compile-time.  The last warp is the TMA producer.
template <int NT8, int EPI>
__global__ void __launch_bounds__(32 * (2 * (EPI == EPI_KE ? 2 : 9) + 1), 1) synth_kernel(SynthParams p, int nstage) {
    constexpr int NSET = (EPI == EPI_DFX) ? 2 : 1;
    constexpr int NF = (EPI == EPI_KE) ? 2 : 9;  // fields per set = consumer warps per parity
    constexpr int RS = NF * NT8 * 8;             // rows per set
    constexpr int MTW = NT8 * NSET;
    constexpr int NT = synth_nt_for(MTW);
    constexpr int W = NT * 8, KS = SYNTH_KS, LDE = synth_lde(NT);
    constexpr int NCW = 2 * NF, NTHR = 32 * (NCW + 1);
    constexpr int A_SET = KS * 2 * RS * 4, A_ST = NSET * A_SET, B_ST = KS * 4 * W * 4, STAGE = A_ST + B_ST;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[SYNTH_MAX_STAGES], bar_empty[SYNTH_MAX_STAGES];
    const Geo& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, jt = blockIdx.x;
    const int nchunk = g.Khp / SYNTH_KC;

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    double acc[MTW][NT][2];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const int par = warp / NF, fw = warp - par * NF;  // consumer identity (garbage for the producer warp)

    if (warp == NCW) {
        // ---------------- producer: one lane streams the stages ----------------
        if (lane == 0) {
            const double* gA0 = p.coef0 + (long long)b * p.coef_stride;
            const double* gA1 = (NSET == 2) ? p.coef1 + (long long)b * p.coef_stride : nullptr;
            const double* gB = p.tab + (long long)jt * nchunk * B_ST;
            int st = 0, ph = 0;
            for (int c = 0; c < nchunk; ++c) {
                if (c >= nstage) mbar_wait(&bar_empty[st], ph ^ 1);
                double* sA = smem + (size_t)st * STAGE;
                mbar_expect_tx(&bar_full[st], (unsigned)(STAGE * sizeof(double)));
                bulk_g2s(sA, gA0 + (long long)c * A_SET, A_SET * sizeof(double), &bar_full[st]);
                if (NSET == 2) bulk_g2s(sA + A_SET, gA1 + (long long)c * A_SET, A_SET * sizeof(double), &bar_full[st]);
                bulk_g2s(sA + A_ST, gB + (long long)c * B_ST, B_ST * sizeof(double), &bar_full[st]);
                if (++st == nstage) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers: DMMA main loop ----------------
        const int ty = (EPI == EPI_KE) ? fw : (fw >= 5 ? 1 : 0);  // 0: cosine table, 1: sine table
        const int a_off = ((par * RS) + fw * NT8 * 8 + gq) * 4 + tq;          // + set*A_SET + ks*2*RS*4 + mt*32
        const int b_off = A_ST + ((ty * 2 + par) * W + gq) * 4 + tq;         // + ks*4*W*4 + nt*32
        int st = 0, ph = 0;
        for (int c = 0; c < nchunk; ++c) {
            mbar_wait(&bar_full[st], ph);
            const double* sS = smem + (size_t)st * STAGE;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double bf[NT], af[MTW];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[nt] = sS[b_off + ks * 4 * W * 4 + nt * 32];
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt)
                    af[mt] = sS[a_off + (mt / NT8) * A_SET + ks * 2 * RS * 4 + (mt % NT8) * 32];
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            if (++st == nstage) { st = 0; ph ^= 1; }
        }
    }
    __syncthreads();  // every stage has been consumed; the staging area is reused below

    // ---- epilogue: exchange E/O through shared memory ----
    constexpr int ROWS = NSET * RS;
    double* sEO = smem;                          // [2 par][ROWS][LDE]
    double* sA1 = sEO + (size_t)2 * ROWS * LDE;  // [2 mirror][n][W]
    double* sDr = sA1 + (size_t)2 * g.n * W;     // [n][n]
    double* red = sDr + (size_t)g.n * g.n;       // [32]
    if (warp < NCW) {
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) {
            const int row = (mt / NT8) * RS + (fw * NT8 + (mt % NT8)) * 8 + gq;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<double2*>(&sEO[((size_t)par * ROWS + row) * LDE + nt * 8 + 2 * tq]) = v;
            }
        }
    }
    const int n = g.n;
    constexpr int n8 = NT8 * 8;
    if (EPI != EPI_KE) {
        for (int idx = tid; idx < n * n; idx += NTHR) sDr[idx] = p.Dr[idx];
    }
    __syncthreads();

    const double* E = sEO;
    const double* O = sEO + (size_t)ROWS * LDE;
    const int npts = n * W;

    if (EPI == EPI_KE) {
        // rows: field 0 = J_theta(psi)/r (cosine), field 1 = Dr psi (sine).  Main.py:117-130.
        double part = 0.0;
        for (int pt = tid; pt < npts; pt += NTHR) {
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
    constexpr int PTS = (n8 * W + NTHR - 1) / NTHR;
    constexpr int ROWS3 = 3 * n8;
    double qv[PTS][2];
    // prd[b][par][chunk][ks][f*n8 + i][4]: the analysis GEMM's A operand, contraction index j' = 8 chunk + 4 ks + kk
    double* prd = p.prd + (long long)b * 2 * g.Mhp * ROWS3;
    const long long pps = (long long)g.Mhp * ROWS3;  // parity stride
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
                    const double e = E[(size_t)(RS + a * n8 + i) * LDE + c], o = O[(size_t)(RS + a * n8 + i) * LDE + c];
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
            const long long oT = prd_off(1, i, c), oS = prd_off(2, i, c);
            prd[oT] = nt0 + nt1;
            prd[pps + oT] = nt0 - nt1;
            prd[oS] = ns0 + ns1;
            prd[pps + oS] = ns0 - ns1;
        }
    }
    __syncthreads();
    // N_psi = Dr @ (JT*omega) - (kDpsi*omega + Dpsi*komega)   (Matrix_Operators.py:791)
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
            // sine-type analysis: odd k uses the sum, even k the difference
            const long long o = prd_off(0, i, c);
            prd[pps + o] = v0 + v1;
            prd[o] = v0 - v1;
        }
    }
}

}  // namespace sddc
