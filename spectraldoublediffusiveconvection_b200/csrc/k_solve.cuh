// Stage 4: implicit block back-substitutions over latitudinal modes (two decoupled parity chains per field),
// batched over members as small DMMA GEMMs  L_inv[j] (n x n) @ RHS (n x members)  with the pre-inverted
// operators streamed through shared memory.
//
// Reference semantics: A4_BSub_TSTEP_V2 (Matrix_Operators.py:1115-1194) and NAB2_BSub_TSTEP_V2 (1033-1086).
#pragma once
#include "common.cuh"

namespace sddc {

struct SolveParams {
    // right-hand sides, one of two layouts:
    //  * state layout (SM = false): g[member*g_stride + fld*g_field_off + row*n + i]   (stand-alone solves)
    //  * solve-major (SM = true):   g[((fld*K + row)*bstride + member)*LDG + i], LDG = n8+2: the 16 members a CTA
    //    owns are one contiguous tile per chain step, fetched by one TMA bulk copy (hot path; written in this
    //    order by prep_kernel / analysis_kernel)
    const double* g;
    const double* fnl;      // optional nonlinear term F(X), same layout as g: rhs = g + mdt * fnl  (Main.py:262,271)
    const double* spec;     // gather mode of the hot kernel (k_solve_hot.cuh, GATH): analysed products [B][n][4][K] of the row
                            // kernels, rows parity-split; F(X) is formed from them inside the chain (fnl unused)
    const double* DrT;      // gather mode: [n][n8], DrT[i'][i] = Dr[i][i']
    double mdt;             // -dt
    long long g_stride;     // state layout: member stride (doubles)
    long long g_field_off;  // state layout: offset between fields inside a member (N, or 0 for single-field calls)
    long long bstride;      // solve-major: members per (field, row) slab (max_batch rounded up to 16)
    double* out;
    long long out_stride;
    long long out_field_off;
    const double* sub;      // optional: out = f - sub (same layout as out) for residual / JVP
    const double* LinvA4;   // [K][2][n8][LDL] zero padded, descending-mode order (index K - j): {L_inv_j, L_inv_j @ D2}
    const double* LinvT;    // [K][n8][LDL] index K-1-j
    const double* LinvS;
    const double* D2;       // [n8][LDL]
    const double* ir2;      // [n] diag IR2 (A4 aux)
    const double* ir4;      // [n]
    Geo geo;
    int B;
    int field_mask;         // bit f: solve field f (0 psi, 1 T, 2 S)
    int field_base;         // field index of chain group 0 (for single-field calls)
    double dt_psi, dt_T, dt_S;  // Pr*dt, dt, Tau*dt
    int nsl;                    // pipeline stages in use (2 or 3)
    double* dpart;              // optional [B][6][3] (k_solve_hot.cuh, DIAG): per chain (field * 2 + parity) the partial sums
                                // of the NEW state's diagnostics: sum f^2, and for the even T / S chains the Nusselt sums at
                                // the inner and outer wall (Main.py:41-68, 292)
    const double *nu_in, *nu_out;   // [n] Nusselt weights (R_w^2 / A_T) D[w, 1:-1]
    const double* nu_w;             // [K] 1 / (1 - k^2) (Main.py:58-63; only even k are used)
    double* jj_out;             // optional [B][K+1][n]: theta-coupling brackets of the NEW stream function, exactly what
                                // scan_kernel would compute from the output (the running sum f_e of the A4 chain is that
                                // suffix sum), so the next step of a multi-step call skips its scan (k_solve_hot.cuh)
};

#ifndef SOLVE_NSL_MAX
#define SOLVE_NSL_MAX 6
#endif
constexpr int SOLVE_NSL = SOLVE_NSL_MAX;  // maximum pipeline stages (operator + right-hand-side tiles)

template <int NTB>
__host__ __device__ inline size_t solve_smem_doubles(int n8, int nsl) {
    const int LDL = n8 + 4, LDG = n8 + SDDC_SM_PAD;
    return (size_t)(2 * nsl) * n8 * LDL + (size_t)4 * (8 * NTB) * LDL + (size_t)nsl * 2 * (8 * NTB) * LDG;
}

// grid = (ceil(B / (8*NTB)), 2 chains, nfields), block = 32 * (nt8 + 1): compute warp w < nt8 owns radial rows
// 8w..8w+7, the last warp is the TMA producer.
// Every thread owns the elements (i = 8w + g, member = nt*8 + 2t + e) in MMA accumulator layout, so the
// running vectors b / f_e / bf_e of the reference live in registers.  Per chain step one thread issues TMA bulk
// copies (pre-inverted operator of the mode, right-hand-side tiles) two steps ahead into a 3-stage ring.
template <int NTB, bool SM>
__global__ void __launch_bounds__(288) solve_kernel(SolveParams p) {
    constexpr int BT = 8 * NTB, NE = 2 * NTB;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[SOLVE_NSL], bar_empty[SOLVE_NSL];
    const Geo& G = p.geo;
    const int n = G.n, n8 = G.n8, K = G.K, LDL = n8 + 4, MAT = n8 * LDL, LDG = n8 + SDDC_SM_PAD, GT = BT * LDG;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int nthr = blockDim.x - 32;          // compute threads (the last warp only streams operands)
    const int ncw = nthr >> 5;
    const bool is_producer = warp == ncw;
    const int b0 = blockIdx.x * BT, which = blockIdx.y, fld = p.field_base + blockIdx.z;
    if (!((p.field_mask >> fld) & 1)) return;
    double* sL = smem;                          // [NSL][2][n8][LDL]  (second matrix only used by the psi chains)
    const int NSL = p.nsl;
    double* sR = sL + (size_t)NSL * 2 * MAT;  // [2 step parities][2][BT][LDL]
    double* sG = sR + (size_t)4 * BT * LDL;     // [NSL][2][BT][LDG]  (solve-major right-hand-side tiles: lin, F)

    const int i = warp * 8 + gq;
    const bool row_ok = i < n;
    long long goff[NE], ooff[NE];
    bool ok[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
        ok[e] = row_ok && (b0 + m) < p.B;
        goff[e] = (long long)(b0 + m) * p.g_stride + (long long)fld * p.g_field_off + i;
        ooff[e] = (long long)(b0 + m) * p.out_stride + (long long)fld * p.out_field_off + i;
    }
    // right-hand side of chain row `row` (pipeline stage st) -> registers
    auto load_g = [&](int row, int st, double* v) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            double x = 0.0;
            if (ok[e]) {
                if (SM) {
                    const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
                    const double* t = sG + (size_t)st * 2 * GT + m * LDG + i;
                    x = t[0];
                    if (p.fnl) x = fma(p.mdt, t[GT], x);
                } else {
                    const long long o = goff[e] + (long long)row * n;
                    x = p.g[o];
                    if (p.fnl) x = fma(p.mdt, p.fnl[o], x);
                }
            }
            v[e] = x;
        }
    };
    // residual / JVP: out = f - sub.  The subtrahend is fetched at the top of the chain step so that its (scattered)
    // load latency is hidden behind the step's GEMM instead of stalling the store.
    double subv[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) subv[e] = 0.0;
    auto load_sub = [&](int row) {
        if (p.sub) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (ok[e]) subv[e] = p.sub[ooff[e] + (long long)row * n];
        }
    };
    auto store_out = [&](int row, const double* f) {
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (ok[e]) p.out[ooff[e] + (long long)row * n] = f[e] - subv[e];
    };
    auto put_rhs = [&](double* buf, const double* v) {
#pragma unroll
        for (int e = 0; e < NE; ++e) buf[((e >> 1) * 8 + 2 * tq + (e & 1)) * LDL + i] = v[e];
    };
    // C = Mat(n8 x n8, smem, row stride LDL) @ V (smem [member][i']); two independent accumulator chains
    // (even / odd k-steps) halve the dependent-MMA latency
    auto gemm = [&](const double* mat, const double* vec, double* c) {
        double c1[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) c[e] = c1[e] = 0.0;
        const double* ar = mat + (warp * 8 + gq) * LDL + tq;
        const double* br = vec + gq * LDL + tq;
#pragma unroll 2
        for (int ks = 0; ks < n8 / 4; ks += 2) {
            const double a0 = ar[ks * 4], a1 = ar[ks * 4 + 4];
#pragma unroll
            for (int nt = 0; nt < NTB; ++nt) {
                mma884(c[2 * nt], c[2 * nt + 1], a0, br[nt * 8 * LDL + ks * 4]);
                mma884(c1[2 * nt], c1[2 * nt + 1], a1, br[nt * 8 * LDL + ks * 4 + 4]);
            }
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) c[e] += c1[e];
    };
    // C = MatA @ VA + MatB @ VB as two interleaved, independent accumulator chains (one per product)
    auto gemm2 = [&](const double* matA, const double* vecA, const double* matB, const double* vecB, double* c) {
        double c1[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) c[e] = c1[e] = 0.0;
        const int ao = (warp * 8 + gq) * LDL + tq, bo = gq * LDL + tq;
#pragma unroll 2
        for (int ks = 0; ks < n8 / 4; ++ks) {
            const double a0 = matA[ao + ks * 4], a1 = matB[ao + ks * 4];
#pragma unroll
            for (int nt = 0; nt < NTB; ++nt) {
                mma884(c[2 * nt], c[2 * nt + 1], a0, vecA[bo + nt * 8 * LDL + ks * 4]);
                mma884(c1[2 * nt], c1[2 * nt + 1], a1, vecB[bo + nt * 8 * LDL + ks * 4]);
            }
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) c[e] += c1[e];
    };

    const bool is_psi = (fld == 0);
    // chain start mode: psi: j0 = K (which 0) | K-1 (which 1, skipped if symmetric)
    //                   T,S: j0 = K-2 (which 0) | K-1 (which 1, skipped if symmetric)
    const int j0 = is_psi ? (K - which) : (which == 0 ? K - 2 : K - 1);
    const int jend = is_psi ? 1 : 0;
    if (G.symmetric && which == 1) {
        if (is_producer) return;
        double z[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) z[e] = 0.0;
        for (int j = j0; j >= jend; j -= 2) {
            load_sub(is_psi ? j - 1 : j);
            store_out(is_psi ? j - 1 : j, z);
        }
        return;
    }
    const double* Lg = is_psi ? p.LinvA4 : (fld == 1 ? p.LinvT : p.LinvS);
    const int nsteps = (j0 - jend) / 2 + 1;
    if (tid == 0) {
        for (int s = 0; s < NSL; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], ncw); }
        mbar_fence_init();
    }
    // chain step `step` handles mode j = j0 - 2*step, state row (j-1 for psi, j otherwise)
    auto producer_loop = [&]() {
        const unsigned tile_bytes = (unsigned)(GT * sizeof(double));
        const unsigned mat_bytes = (unsigned)((is_psi ? 2 : 1) * MAT * sizeof(double));
        const unsigned bytes = mat_bytes + (SM ? tile_bytes * (p.fnl ? 2u : 1u) : 0u);
        int st = 0, ph = 0;
        for (int step = 0; step < nsteps; ++step) {
            const int j = j0 - 2 * step;
            const int jj = is_psi ? (K - j) : (K - 1 - j), row = is_psi ? j - 1 : j;
            if (step >= NSL) mbar_wait(&bar_empty[st], ph ^ 1);
            mbar_expect_tx(&bar_full[st], bytes);
            bulk_g2s(sL + (size_t)st * 2 * MAT, Lg + (long long)jj * (is_psi ? 2 : 1) * MAT, mat_bytes, &bar_full[st]);
            if (SM) {
                const long long o = (((long long)fld * K + row) * p.bstride + b0) * LDG;
                bulk_g2s(sG + (size_t)st * 2 * GT, p.g + o, tile_bytes, &bar_full[st]);
                if (p.fnl) bulk_g2s(sG + (size_t)st * 2 * GT + GT, p.fnl + o, tile_bytes, &bar_full[st]);
            }
            if (++st == NSL) { st = 0; ph ^= 1; }
        }
    };
    for (int idx = tid; idx < 4 * BT * LDL; idx += nthr) sR[idx] = 0.0;  // padded rows stay zero
    if (is_producer) {
        // (no shared-memory initialisation duty: threads >= nthr skip the loop above)
    }
    __syncthreads();
    if (is_producer) {
        if (lane == 0) producer_loop();
        return;
    }

    double f[NE], gv[NE];
    if (!is_psi) {
        const double dt = (fld == 1) ? p.dt_T : p.dt_S;
        double bsum[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) { bsum[e] = 0.0; f[e] = 0.0; }
        int st = 0, ph = 0;
        for (int step = 0; step < nsteps; ++step) {
            const int j = j0 - 2 * step;
            load_sub(j);
            mbar_wait(&bar_full[st], ph);
            load_g(j, st, gv);
            double rhs[NE];
            const double beta = 2.0 * dt * (j + 2.0);
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (j < K - 2) bsum[e] += beta * f[e];
                rhs[e] = gv[e] - (j == 0 ? 0.5 * bsum[e] : bsum[e]);
            }
            double* buf = sR + (size_t)(step & 1) * BT * LDL;
            put_rhs(buf, rhs);
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // compute warps only
            gemm(sL + (size_t)st * 2 * MAT, buf, f);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);   // operator + tiles of this stage are consumed
            if (++st == NSL) { st = 0; ph ^= 1; }
            store_out(j, f);
        }
    } else {
        const double dt = p.dt_psi;
        double fe[NE], bfe[NE];
        const double ir2 = row_ok ? p.ir2[i] : 0.0, ir4 = row_ok ? p.ir4[i] : 0.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) { fe[e] = 0.0; bfe[e] = 0.0; f[e] = 0.0; }
        double* bufA = sR;                        // dt*bjt*f_e   (multiplied by L_inv_j @ D2)
        double* bufB = sR + (size_t)BT * LDL;     // elementwise part of the right-hand side (multiplied by L_inv_j)
        // With L1_j = D2 + b_j IR4 (Matrix_Operators.py:1149) the reference's update
        //     f_j = L_inv_j @ ( g_j + dt*bjt*(L1_j @ f_e + IR4 @ bf_e) - bjt*IR2 @ f_e )
        // is evaluated as  L_inv_j @ rhs_elem + (L_inv_j @ D2) @ (dt*bjt*f_e)  with the product L_inv_j @ D2 formed
        // once on the host: one barrier and one (double-width) GEMM round per chain step instead of two.
        int st = 0, ph = 0;
        for (int step = 0; step < nsteps; ++step) {
            const int j = j0 - 2 * step;
            const double bj = -(double)j * (j + 1.0), bjt = -2.0 * j;
            double rhs[NE], sfe[NE];
            load_sub(j - 1);
            mbar_wait(&bar_full[st], ph);
            load_g(j - 1, st, gv);
            if (step == 0) {
#pragma unroll
                for (int e = 0; e < NE; ++e) { rhs[e] = gv[e]; sfe[e] = 0.0; }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    fe[e] += f[e];
                    rhs[e] = gv[e] + (dt * bjt * (bj * (ir4 * fe[e]) + ir4 * bfe[e]) - bjt * (ir2 * fe[e]));
                    sfe[e] = (dt * bjt) * fe[e];
                }
            }
            double* bA = bufA + (size_t)(step & 1) * 2 * BT * LDL;
            double* bB = bufB + (size_t)(step & 1) * 2 * BT * LDL;
            put_rhs(bA, sfe);
            put_rhs(bB, rhs);
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // compute warps only
            const double* Lm = sL + (size_t)st * 2 * MAT;
            gemm2(Lm, bB, Lm + MAT, bA, f);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            if (++st == NSL) { st = 0; ph ^= 1; }
            store_out(j - 1, f);
#pragma unroll
            for (int e = 0; e < NE; ++e) bfe[e] += (step == 0) ? bj * f[e] : (bj * f[e] + bjt * fe[e]);
        }
    }
}

}  // namespace sddc
