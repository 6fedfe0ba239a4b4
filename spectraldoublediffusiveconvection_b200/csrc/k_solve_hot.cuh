// Stage 4, hot path: the back-substitution chains of k_solve.cuh specialised for the solve-major layout with every
// size a compile-time constant.  One chain step is a serial dependency (f_j needs f_{j+2}), and a CTA has only one
// warp per scheduler, so the step time is the *instruction count* of the loop body times the issue latency: here
// all shared-memory offsets are immediates, the output pointers are carried and decremented, and the special cases
// of the recurrences (first step, mode 0) are folded into data instead of branches.
//
// The stream-function chains do twice the tensor work per step (L_inv_j and L_inv_j @ D2); they run on narrower
// member tiles (SOLVE_NTB_PSI) than the scalar chains (SOLVE_NTB_TS) so that every CTA of the launch has about the
// same step time.
//
// Reference semantics: A4_BSub_TSTEP_V2 (Matrix_Operators.py:1115-1194) and NAB2_BSub_TSTEP_V2 (1033-1086).
#pragma once
#include "k_solve.cuh"

namespace sddc {

#ifndef SOLVE_NTB_PSI
#define SOLVE_NTB_PSI 1
#endif
#ifndef SOLVE_NTB_TS
#define SOLVE_NTB_TS 2
#endif

// GATH -- EXPERIMENTAL, compiled in but only selected by plans of builds with -DSDDC_EXPERIMENTAL_GATHER (its results were
// not bit-reproducible run to run and the race was not found: DESIGN.md section 4) -- (FFT formulation, n8 <= 32, K a
// multiple of 8; from 128 members on): the nonlinear term is not read as a
// solve-major tile that a separate kernel (post_kernel) transposed; the chain gathers the analysed products of the row
// kernel itself.  spec[b][i][field][K] holds every row parity-split (fft_core.h: spec_pos), so the four coefficients a
// chain needs in four consecutive steps are one 32-byte sector: a producer warp of its own (the warp behind the TMA
// producer) fetches a GROUP (four chain steps x members x radial rows x fields) with two adjacent 16-byte cp.async per
// (member, row, field) -- a lane pair covers the sector -- completion on an mbarrier (cp.async.mbarrier.arrive.noinc),
// SOLVE_NSF groups in flight; the consumers hand a slot back through a second mbarrier.  The stream-function chains apply
// Dr @ to the first product with the operator's A fragments held in registers (F_psi = Dr @ DST(JT*om) - DST(..),
// Matrix_Operators.py:776-802) one chain step AHEAD: the operands are loaded early in the previous step, the MMAs issued
// behind that step's own, so that none of it sits on the dependent chain.
// Shared-memory layout of a group: [16-byte half][member][row][2 steps] (half 1 = the first two of the four steps, each
// half [later step, earlier step]); the Dr @ operand is read as MMA B fragments (member = lane group, row = lane in
// group: row pitch n8 + 4), everything else in accumulator layout (member pair = lane in group, row = lane group:
// pitch n8 + 1).  What else was measured (ring depths, TMA bulk copies, auxiliary warps, pre-tiled spectra): DESIGN.md.
constexpr int SOLVE_NSF = 2;      // gather groups in flight
constexpr int SOLVE_GSTEPS = 4;   // chain steps per group (a 32-byte sector per member, row and field)

template <int NTB, bool PSI>
__host__ __device__ constexpr size_t solve_gath_slot_doubles(int n8) {
    return (size_t)(8 * NTB) * ((PSI ? (n8 + 4) : 0) + (n8 + 1)) * SOLVE_GSTEPS;
}
template <int NTB, bool PSI>
__host__ __device__ constexpr size_t solve_gath_doubles(int n8, int nsl) {
    const int LDL = n8 + 4, LDG = n8 + SDDC_SM_PAD, NM = PSI ? 2 : 1;
    return (size_t)nsl * NM * n8 * LDL + (size_t)2 * NM * (8 * NTB) * LDL + (size_t)nsl * (8 * NTB) * LDG +
           (size_t)SOLVE_NSF * solve_gath_slot_doubles<NTB, PSI>(n8);
}

template <int NTB, bool PSI>
__host__ __device__ constexpr size_t solve_hot_doubles(int n8, int nsl) {
    const int LDL = n8 + 4, LDG = n8 + SDDC_SM_PAD, NM = PSI ? 2 : 1;
    return (size_t)nsl * NM * n8 * LDL + (size_t)2 * NM * (8 * NTB) * LDL + (size_t)nsl * 2 * (8 * NTB) * LDG;
}
// nsl = pipeline stages (operator + right-hand-side tiles): 3, or 2 when three do not fit an SM
__host__ inline size_t solve_hot_smem_bytes(int n8, int nsl) {
    const size_t a = solve_hot_doubles<SOLVE_NTB_PSI, true>(n8, nsl), b = solve_hot_doubles<SOLVE_NTB_TS, false>(n8, nsl);
    return sizeof(double) * (a > b ? a : b);
}

__host__ inline size_t solve_gath_smem_bytes(int n8) {
    const size_t a = solve_gath_doubles<SOLVE_NTB_PSI, true>(n8, 3), b = solve_gath_doubles<SOLVE_NTB_TS, false>(n8, 3);
    return sizeof(double) * (a > b ? a : b);
}

// SUB: the call carries a subtrahend (residual / JVP); a compile-time switch so that the plain step keeps its registers
// DIAG: the chain also accumulates its share of ||X'||^2 and of the Nusselt sums of the state it produces (p.dpart)
// GATH: the nonlinear term comes from p.spec (see above) instead of the solve-major tile p.fnl
template <int NT8, int NSL, int NTB, bool PSI, bool SUB = true, bool DIAG = false, bool GATH = false>
__device__ __forceinline__ void solve_chain_hot(const SolveParams& p, double* smem, uint64_t* bar_full,
                                                uint64_t* bar_empty, uint64_t* bar_gfull, uint64_t* bar_gempty, int fld,
                                                int which, int b0) {
    constexpr int n8 = 8 * NT8, LDL = n8 + 4, MAT = n8 * LDL, LDG = n8 + SDDC_SM_PAD, BT = 8 * NTB, GT = BT * LDG, NE = 2 * NTB;
    constexpr int NM = PSI ? 2 : 1, NTHR = 32 * NT8;
    constexpr int NCH = NTB >= 2 ? 1 : 2;   // accumulator chains per product and member tile (k-steps interleaved)
    constexpr int NGT = GATH ? 1 : 2;       // right-hand-side tiles per ring stage (lin [, F])
    constexpr int LDF1 = n8 + 4, LDF2 = n8 + 1, GS = SOLVE_GSTEPS;
    constexpr int F2OFF = PSI ? BT * LDF1 * 2 : 0, SLOT = (int)solve_gath_slot_doubles<NTB, PSI>(n8), HSZ = SLOT / 2;
    static_assert(GS == 4 && SOLVE_NSF == 2, "group geometry");
    static_assert(!GATH || (NSL == 3 && n8 <= 32), "gather mode: three ring stages, n8 <= 32");
    const Geo& G = p.geo;
    const int n = G.n, K = G.K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const bool is_producer = warp >= NT8;   // warp NT8: operator / right-hand-side ring (TMA); warp NT8 + 1 (GATH): gathers
    double* sL = smem;                              // [NSL][NM][n8][LDL]
    double* sR = sL + (size_t)NSL * NM * MAT;       // [2 step parities][NM][BT][LDL]
    double* sG = sR + (size_t)2 * NM * BT * LDL;    // [NSL][NGT][BT][LDG]   (lin [, F])
    double* sF = sG + (size_t)NSL * NGT * GT;       // GATH: [SOLVE_NSF][2 halves][{P1: [BT][LDF1][2]}, [BT][LDF2][2]]
    const int i = warp * 8 + gq;
    const bool row_ok = i < n;
    const bool has_f = !GATH && p.fnl != nullptr, has_sub = SUB && p.sub != nullptr;

    const int j0 = PSI ? (K - which) : (which == 0 ? K - 2 : K - 1);
    const int jend = PSI ? 1 : 0;
    const int row0 = PSI ? j0 - 1 : j0;
    const int nsteps = (j0 - jend) / 2 + 1;
    // GATH: chain step s >= so0 reads the parity-split position ptop - (s - so0); the block K-1 of the stream function
    // (first step of the even chain) has no nonlinear term (Matrix_Operators.py:802)
    const int so0 = (PSI && which == 0) ? 1 : 0, ptop = which * (K >> 1) + (K >> 1) - 1;
    const int ngroups = (nsteps - so0 + GS - 1) / GS;

    double* outp[NE];
    bool ok[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
        ok[e] = row_ok && (b0 + m) < p.B;
        outp[e] = p.out + (long long)(b0 + m) * p.out_stride + (long long)fld * p.out_field_off + (long long)row0 * n + i;
    }
    const long long sub_delta = has_sub ? (p.sub - p.out) : 0;
    const long long rstep = -2LL * n;
    // JJ[b][m][i] of the new stream function (k_prep.cuh, scan_kernel), sine mode m = j of this chain
    const bool want_jj = PSI && p.jj_out != nullptr;
    double* jjp[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
        jjp[e] = want_jj ? p.jj_out + ((long long)(b0 + m) * (K + 1) + j0) * n + i : nullptr;
    }

    if (G.symmetric && which == 1) {   // these modes are identically zero under the equatorial symmetry
        pdl_wait();
        if (DIAG) {
            for (int m = tid; m < 3 * BT; m += blockDim.x)
                if (b0 + m / 3 < p.B) p.dpart[((long long)(b0 + m / 3) * 6 + fld * 2 + which) * 3 + m % 3] = 0.0;
        }
        if (is_producer) return;
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (ok[e]) outp[e][0] = has_sub ? -outp[e][sub_delta] : 0.0;
                outp[e] += rstep;
                if (want_jj) {
                    if (ok[e]) jjp[e][0] = 0.0;
                    jjp[e] += rstep;
                }
            }
        }
        return;
    }
    if (tid == 0) {
        for (int s = 0; s < NSL; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NT8); }
        if (GATH)
            for (int s = 0; s < SOLVE_NSF; ++s) { mbar_init(&bar_gfull[s], 32); mbar_init(&bar_gempty[s], NT8); }
        mbar_fence_init();
    }
    for (int idx = tid; idx < 2 * NM * BT * LDL; idx += blockDim.x) sR[idx] = 0.0;   // padded rows stay zero
    if (GATH)
        for (int idx = tid; idx < SOLVE_NSF * SLOT; idx += blockDim.x) sF[idx] = 0.0;   // padded rows / absent members stay zero
    __syncthreads();
    pdl_wait();   // right-hand sides from prep_kernel / post_kernel

    if (is_producer && warp == NT8) {
      if (lane != 0) return;
      const double* Lg = PSI ? p.LinvA4 : (fld == 1 ? p.LinvT : p.LinvS);
      constexpr unsigned tile_bytes = GT * sizeof(double), mat_bytes = NM * MAT * sizeof(double);
      const unsigned bytes = mat_bytes + tile_bytes * (has_f ? 2u : 1u);
      int st = 0, ph = 0;
      for (int step = 0; step < nsteps; ++step) {
          const int j = j0 - 2 * step;
          const int jj = PSI ? (K - j) : (K - 1 - j), row = PSI ? j - 1 : j;
          if (step >= NSL) mbar_wait(&bar_empty[st], ph ^ 1);
          mbar_expect_tx(&bar_full[st], bytes);
          bulk_g2s(sL + (size_t)st * NM * MAT, Lg + (long long)jj * NM * MAT, mat_bytes, &bar_full[st]);
          const long long o = (((long long)fld * K + row) * p.bstride + b0) * LDG;
          bulk_g2s(sG + (size_t)st * NGT * GT, p.g + o, tile_bytes, &bar_full[st]);
          if (has_f) bulk_g2s(sG + (size_t)st * NGT * GT + GT, p.fnl + o, tile_bytes, &bar_full[st]);
          if (++st == NSL) { st = 0; ph ^= 1; }
      }
      return;
    }
    if (is_producer) {
      // GATH, second producer warp: group g = positions ptop - 4g - 3 .. ptop - 4g of every (member, row, field) of the
      // tile; lane = (row & 15, 16-byte half); a slot is refilled as soon as the consumers have released it
      constexpr int RPP = 16;   // rows per pass of the warp
      const int gh = lane & 1, gi = lane >> 1;
      const long long SP = spec_pitch(K);
      const double* gsrc = p.spec + (long long)b0 * n * SP + (PSI ? 0 : fld + 1) * K + (ptop - (GS - 1) + 2 * gh);
      for (int g = 0; g < ngroups; ++g) {
          const int slot = g & (SOLVE_NSF - 1);
          if (g >= SOLVE_NSF) mbar_wait(&bar_gempty[slot], ((g / SOLVE_NSF) - 1) & 1);
          double* dst0 = sF + slot * SLOT + gh * HSZ;
          const double* src0 = gsrc - GS * g;
#pragma unroll
          for (int q = 0; q < NM; ++q)
#pragma unroll 4
              for (int m = 0; m < BT; ++m) {
                  if (b0 + m < p.B) {
#pragma unroll
                      for (int ip = 0; ip < (n8 + RPP - 1) / RPP; ++ip) {
                          const int ii = gi + RPP * ip;
                          if (ii < n)
                              cp_async16(dst0 + (q == 0 && PSI ? 0 : F2OFF) + (m * (q == 0 && PSI ? LDF1 : LDF2) + ii) * 2,
                                         src0 + (long long)(m * n + ii) * SP + q * K);
                      }
                  }
              }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_gfull[slot])) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      return;
    }

    // per-thread shared-memory bases; everything below is base + stage * constant + immediate
    const double* gT = sG + (2 * tq) * LDG + i;          // right-hand-side tile element (member 2tq, row i)
    double* rW = sR + (2 * tq) * LDL + i;                // GEMM operand element (member 2tq, row i)
    const double* aP = sL + (warp * 8 + gq) * LDL + tq;  // A fragment: L_inv[8w+g][4ks+t]
    const double* bP = sR + gq * LDL + tq;               // B fragment: rhs[4ks+t][member g]

    double f[NE], subv[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) { f[e] = 0.0; subv[e] = 0.0; }
    double s1[NE], s2[NE];   // T,S: s1 = running sum b;   psi: s1 = f_e, s2 = bf_e
#pragma unroll
    for (int e = 0; e < NE; ++e) { s1[e] = 0.0; s2[e] = 0.0; }
    const double dt = PSI ? p.dt_psi : (fld == 1 ? p.dt_T : p.dt_S);
    const double ir2 = (PSI && row_ok) ? p.ir2[i] : 0.0, ir4 = (PSI && row_ok) ? p.ir4[i] : 0.0;

    // subtrahend of the residual / JVP (state layout, scattered 8-byte loads from HBM): fetched three chain steps ahead --
    // ncu showed the consumers stalled on these loads (long scoreboard 6.5 per issue against 1.6 without a subtrahend)
    double subq[3][NE];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int e = 0; e < NE; ++e) subq[d][e] = (has_sub && ok[e] && d < nsteps) ? outp[e][sub_delta + d * rstep] : 0.0;
    double dn2[NE], dni[NE], dno[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) { dn2[e] = 0.0; dni[e] = 0.0; dno[e] = 0.0; }
    const bool nu_chain = DIAG && !PSI && which == 0;   // cosine modes 0, 2, 4, ... of T / S (Main.py:58-63)
    const double nu_i = (nu_chain && row_ok) ? p.nu_in[i] : 0.0, nu_o = (nu_chain && row_ok) ? p.nu_out[i] : 0.0;
    // GATH: A fragments of Dr (rows 8w..8w+7 of this warp), and the nonlinear term of chain step s from the gathered group
    double aDr[GATH && PSI ? n8 / 4 : 1];
    if (GATH && PSI) {
#pragma unroll
        for (int ks = 0; ks < n8 / 4; ++ks)
            aDr[ks] = (row_ok && 4 * ks + tq < n) ? p.DrT[(4 * ks + tq) * n8 + i] : 0.0;
    }
    // fetch_fn(s): the operands of F at chain step s from the gathered group into registers (issued early in the
    // previous step, so that the shared-memory latency is covered by the right-hand-side arithmetic and the barrier);
    // finish_fn(): fn = Dr @ P1 - P2 (stream function) or the product itself, after the chain step's MMAs are issued
    double fn[NE], fb[GATH && PSI ? (n8 / 4) * NTB : 1], fp[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) { fn[e] = 0.0; fp[e] = 0.0; }
    bool fn_ok = false;
    auto fetch_fn = [&](int s) {
        const int d = s - so0;
        fn_ok = d >= 0 && s < nsteps;
        if (!fn_ok) return;
        const int g = d / GS, r = d & (GS - 1), slot = g & (SOLVE_NSF - 1);
        if (r == 0) mbar_wait(&bar_gfull[slot], (g / SOLVE_NSF) & 1);
        // element GS - 1 - r of the group: 16-byte half (r < 2 ? 1 : 0 when there are two), second double of it when r is even
        const double* fs = sF + slot * SLOT + (r < GS - 2 ? HSZ : 0) + ((r & 1) ? 0 : 1);
        const double* f2 = fs + F2OFF + ((2 * tq) * LDF2 + i) * 2;
        if (PSI) {
            const double* f1 = fs + (gq * LDF1 + tq) * 2;
#pragma unroll
            for (int ks = 0; ks < n8 / 4; ++ks)
#pragma unroll
                for (int nt = 0; nt < NTB; ++nt) fb[ks * NTB + nt] = f1[(nt * 8 * LDF1 + 4 * ks) * 2];
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) fp[e] = f2[((e >> 1) * 8 + (e & 1)) * LDF2 * 2];
        if (r == GS - 1 || s == nsteps - 1) {   // last use of this group
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gempty[slot]);
        }
    };
    auto finish_fn = [&]() {
        if (!fn_ok) {
#pragma unroll
            for (int e = 0; e < NE; ++e) fn[e] = 0.0;
            return;
        }
        if (PSI) {
            double dacc[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) dacc[e] = 0.0;
#pragma unroll
            for (int ks = 0; ks < n8 / 4; ++ks)
#pragma unroll
                for (int nt = 0; nt < NTB; ++nt) mma884(dacc[2 * nt], dacc[2 * nt + 1], aDr[ks], fb[ks * NTB + nt]);
#pragma unroll
            for (int e = 0; e < NE; ++e) fn[e] = dacc[e] - fp[e];
        } else {
#pragma unroll
            for (int e = 0; e < NE; ++e) fn[e] = fp[e];
        }
    };
    if (GATH) { fetch_fn(0); finish_fn(); }
    int st = 0, ph = 0;
    for (int step = 0; step < nsteps; ++step) {
        const int j = j0 - 2 * step;
        if (has_sub) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                subv[e] = subq[0][e];
                subq[0][e] = subq[1][e];
                subq[1][e] = subq[2][e];
                if (ok[e] && step + 3 < nsteps) subq[2][e] = outp[e][sub_delta + 3 * rstep];
            }
        }
        mbar_wait(&bar_full[st], ph);
        const double* gt = gT + st * (NGT * GT);
        double gv[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int off = ((e >> 1) * 8 + (e & 1)) * LDG;
            gv[e] = gt[off];
            if (GATH) gv[e] = fma(p.mdt, fn[e], gv[e]);
            else if (has_f) gv[e] = fma(p.mdt, gt[GT + off], gv[e]);
        }
        if (GATH) fetch_fn(step + 1);
        double* buf = rW + (step & 1) * (NM * BT * LDL);
        if (!PSI) {
            // b += 2 dt (j+2) f_{j+2};  rhs = g_j - b   (halved for mode 0)       (Matrix_Operators.py:1063-1079)
            const double beta = 2.0 * dt * (j + 2.0), scale = (j == 0) ? -0.5 : -1.0;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                s1[e] = fma(beta, f[e], s1[e]);
                const double rhs = fma(scale, s1[e], gv[e]);
                if (row_ok) buf[((e >> 1) * 8 + (e & 1)) * LDL] = rhs;
            }
        } else {
            // f_j = L_inv_j ( g_j + dt bjt (L1_j f_e + IR4 bf_e) - bjt IR2 f_e ),  L1_j = D2 + b_j IR4
            //     = L_inv_j ( g_j + c1 f_e + c2 bf_e ) + (L_inv_j D2) (dt bjt f_e)   (Matrix_Operators.py:1149-1180)
            const double bj = -(double)j * (j + 1.0), bjt = -2.0 * j;
            const double c3 = dt * bjt, c2 = c3 * ir4, c1 = bjt * (dt * bj * ir4 - ir2);
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                s1[e] += f[e];
                const double rhs = fma(c2, s2[e], fma(c1, s1[e], gv[e]));
                if (row_ok) {
                    buf[((e >> 1) * 8 + (e & 1)) * LDL] = rhs;
                    buf[BT * LDL + ((e >> 1) * 8 + (e & 1)) * LDL] = c3 * s1[e];
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");   // compute warps only
        {
            const double* a = aP + st * (NM * MAT);
            const double* b = bP + (step & 1) * (NM * BT * LDL);
            double c[NM][NCH][NE];
#pragma unroll
            for (int q = 0; q < NM; ++q)
#pragma unroll
                for (int h = 0; h < NCH; ++h)
#pragma unroll
                    for (int e = 0; e < NE; ++e) c[q][h][e] = 0.0;
#pragma unroll
            for (int ks = 0; ks < n8 / 4; ++ks) {
#pragma unroll
                for (int q = 0; q < NM; ++q) {
                    const double av = a[q * MAT + ks * 4];
#pragma unroll
                    for (int nt = 0; nt < NTB; ++nt)
                        mma884(c[q][ks % NCH][2 * nt], c[q][ks % NCH][2 * nt + 1], av,
                               b[q * BT * LDL + nt * 8 * LDL + ks * 4]);
                }
            }
            if (GATH) finish_fn();   // independent of this step's result: fills the MMA latency
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < NM; ++q)
#pragma unroll
                    for (int h = 0; h < NCH; ++h) s += c[q][h][e];
                f[e] = s;
            }
        }
        if (DIAG) {
            const double wk = nu_chain ? __ldg(p.nu_w + j) : 0.0;   // 1 / (1 - j^2), tabulated: a division per chain step showed
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                dn2[e] = fma(f[e], f[e], dn2[e]);
                if (nu_chain) {
                    const double t = f[e] * wk;
                    dni[e] = fma(nu_i, t, dni[e]);
                    dno[e] = fma(nu_o, t, dno[e]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);   // operator + tiles of this stage are consumed
        if (++st == NSL) { st = 0; ph ^= 1; }
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (ok[e]) outp[e][0] = f[e] - subv[e];
            outp[e] += rstep;
        }
        if (want_jj) {
            // (j+1) psi^(j) + 2 S[j] with S[j] = f_e, rounded exactly like scan_kernel does it (product first)
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (ok[e]) jjp[e][0] = __fma_rn(s1[e], 2.0, __dmul_rn((double)j + 1.0, f[e]));
                jjp[e] += rstep;
            }
        }
        if (PSI) {
            const double bj = -(double)j * (j + 1.0), bjt = -2.0 * j;
#pragma unroll
            for (int e = 0; e < NE; ++e) s2[e] += fma(bj, f[e], bjt * s1[e]);
        }
    }
    if (want_jj && which == 0) {
        // JJ[0] = S[0] = sum of all even sine modes: the chain ended at j = 2 and the pointers now sit at m = 0
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (ok[e]) jjp[e][0] = s1[e] + f[e];
    }
    if (DIAG) {
        // sum over the radial rows: the eight row groups of a warp by shuffles, the warps through shared memory in a
        // fixed order; the lanes 0..3 of warp 0 then own the members 2 tq, 2 tq + 1 (+ 8 per member tile)
        // (scratch: the GEMM operand buffers sR, free once every warp has left the last chain step)
        static_assert(2 * NM * LDL >= NT8 * 3, "reduction scratch does not fit into the operand buffers");
        double (*s_dred)[BT][3] = reinterpret_cast<double (*)[BT][3]>(sR);
        asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            double v0 = dn2[e], v1 = dni[e], v2 = dno[e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            if (gq == 0) {
                const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
                s_dred[warp][m][0] = v0; s_dred[warp][m][1] = v1; s_dred[warp][m][2] = v2;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");
        if (warp == 0 && gq == 0) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int m = (e >> 1) * 8 + 2 * tq + (e & 1);
                if (b0 + m < p.B) {
                    double* d = p.dpart + ((long long)(b0 + m) * 6 + fld * 2 + which) * 3;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        double t = 0.0;
                        for (int w2 = 0; w2 < NT8; ++w2) t += s_dred[w2][m][q];
                        d[q] = t;
                    }
                }
            }
        }
    }
}

// grid = 2 * (psi member tiles) + 4 * (T,S member tiles) CTAs, stream-function chains first; block = 32 * (NT8 + 1):
// compute warp w owns radial rows 8w..8w+7 in MMA accumulator layout, the last warp is the TMA producer.
template <int NT8, int NSL, bool SUB = false, bool DIAG = false, bool GATH = false>
__global__ void __launch_bounds__(32 * (NT8 + (GATH ? 2 : 1))) solve_hot_kernel(SolveParams p, int npsi_tiles) {
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t bar_full[NSL], bar_empty[NSL], bar_gfull[SOLVE_NSF], bar_gempty[SOLVE_NSF];
    pdl_launch_dependents();
    const int bid = blockIdx.x;
    if (bid < 2 * npsi_tiles) {
        solve_chain_hot<NT8, NSL, SOLVE_NTB_PSI, true, SUB, DIAG, GATH>(p, smem, bar_full, bar_empty, bar_gfull, bar_gempty, 0,
                                                                        bid & 1, (bid >> 1) * 8 * SOLVE_NTB_PSI);
    } else {
        const int r = bid - 2 * npsi_tiles;
        solve_chain_hot<NT8, NSL, SOLVE_NTB_TS, false, SUB, DIAG, GATH>(p, smem, bar_full, bar_empty, bar_gfull, bar_gempty,
                                                                        1 + ((r >> 1) & 1), r & 1, (r >> 2) * 8 * SOLVE_NTB_TS);
    }
}

}  // namespace sddc
