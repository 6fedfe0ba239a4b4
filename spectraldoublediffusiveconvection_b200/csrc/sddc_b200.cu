// Host side of the C ABI (include/sddc_b200.h): plan construction, operator upload / padding, kernel dispatch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/sddc_b200.h"
#include "common.cuh"
#include "k_analysis.cuh"
#include "k_misc.cuh"
#include "k_prep.cuh"
#include "k_solve.cuh"
#include "k_solve_hot.cuh"
#include "k_synth.cuh"
#include "k_synth_wsq.cuh"
#include "k_nlin_fft.cuh"
#include "k_krylov.cuh"

using namespace sddc;

namespace {

thread_local std::string g_create_error;
constexpr size_t SMEM_LIMIT = 227 * 1024 - 1024;  // dynamic shared memory we opt in to (1 KB left for static barriers)
#ifndef KE_NW768
#define KE_NW768 8   // (6: 0.093 ms, 8: 0.084 ms, 10: 0.093 ms at (30,256), 512 members)  // workers per CTA of the kinetic-energy transform at M = 768
#endif
#ifndef KE_NW1536
#define KE_NW1536 4  // 40 KB of planes + rows per worker at M = 1536 (N_fm = 512)
#endif
#ifndef SOLVE_NTB
#define SOLVE_NTB 2  // members per back-substitution CTA = 8 * SOLVE_NTB (1 and 4 measured slower)
#endif

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct DevBuf {
    double* p = nullptr;
    size_t n = 0;
};

}  // namespace

struct sddc_plan {
    sddc_config cfg{};
    Geo g{};
    int device = 0;
    int LDL = 0;
    int Mh3p = 0;  // padded mirror pairs of the 3K kinetic-energy grid
    int nke = 0;   // kinetic-energy partial sums per member
    double ke_scale = 0.0;
    long long launches = 0;
    std::string err;
    std::vector<void*> allocs;
    std::vector<double> h_LinvA4, h_D2;  // host copies (unpadded) to rebuild the fused A4 stack
    // operators
    double *DrT = nullptr, *D2rT = nullptr, *DsqT = nullptr, *Dr = nullptr, *D2p = nullptr;
    double *DrP = nullptr, *D2rP = nullptr, *DsqP = nullptr;  // [n8][n8+4] zero-padded row-major
    double *LA4 = nullptr, *LT = nullptr, *LS = nullptr;
    double *ir2 = nullptr, *ir4 = nullptr, *r2 = nullptr, *dT0 = nullptr, *gb = nullptr, *a4_ir2 = nullptr,
           *a4_ir4 = nullptr, *ir = nullptr, *nu_in = nullptr, *nu_out = nullptr, *wr = nullptr;
    double *tab1 = nullptr, *tab1d = nullptr, *tab2 = nullptr, *tab3 = nullptr, *wth = nullptr;
    double *tab1q = nullptr, *tab2q = nullptr;  // second-mirror-level tables (k_synth_wsq.cuh)
    bool quarter = false;
    // scratch
    double *JJ = nullptr, *coef = nullptr, *prd = nullptr, *lin = nullptr, *rhs = nullptr, *xtmp = nullptr,
           *kepart = nullptr, *zeroRa = nullptr;
    double* coef1 = nullptr;
    double* nu_w = nullptr;    // [K] 1 / (1 - k^2)
    double* dpart = nullptr;   // [max_batch][6][3] diagnostics partial sums written by the back-substitution (DIAG)
    long long coef_member_stride = 0;
    double *lin_sm = nullptr, *f_sm = nullptr;  // solve-major [3][K][bstride][n8+2]
    double *gridc = nullptr, *xbase = nullptr;  // cached base state of sddc_jvp_set_base (lazily allocated)
    // FFT formulation of the nonlinear term (k_nlin_fft.cuh): available for N_fm = 128, 256, 512
    int fft_M = 0;              // 3 N_fm / 2 when the row pipeline (prep -> row kernel -> post) is active, else 0
    bool fft_dfx = false;       // two-state (JVP) products through the row pipeline
    bool fft_direct = false;    // row kernel = nlin_direct_kernel (N_fm = 2 mod 4) instead of the FFT kernels
    bool dfx_direct = false;    // only the two-state products take the (direct) row pipeline: dense shapes beyond N_r = 41
    int ke_M = 0;               // 3 N_fm when the kinetic-energy synthesis runs as an FFT (N_fm = 128, 256), else 0
    double *ke_tab = nullptr, *ke_Wn = nullptr;
    int* fft_row = nullptr;     // [2] row counter / finished-CTA counter of the row kernels (self-resetting: k_nlin_fft.cuh)
    double *coef7 = nullptr, *coef7b = nullptr, *coef7base = nullptr, *spec4 = nullptr, *fft_tab = nullptr;
    double* grid7 = nullptr;    // [max_batch n][7][M] grid fields of the base state of sddc_jvp_set_base (FFT kernels; lazily allocated)
    int base_B = 0;
    long long bstride = 0;
    // host-API staging
    double *hX0 = nullptr, *hX1 = nullptr, *hX2 = nullptr, *hRa = nullptr, *hRas = nullptr, *hDiag = nullptr;
    double* hHist = nullptr;  // [hist_cap][max_batch][6] diagnostics history of sddc_time_step_host
    int hist_cap = 0;
    int ckpt_phase = 0;       // sddc_plan_set_ckpt_phase: checkpoints after the steps phase, phase + every, ... (0: every, 2 every, ...)
    cudaEvent_t ev_ckpt = nullptr;
    // kinetic energy of a record under the next step's back-substitution (sddc_time_step*): side stream + events
    cudaStream_t ke_stream = nullptr;
    cudaEvent_t ev_rows = nullptr, ev_ke = nullptr;
    cudaStream_t own_stream = nullptr, in_stream = nullptr, out_stream = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;
    // kernel configuration
    bool dfx_ok = true;
    bool ws_ok = false;       // persistent warp-specialised quarter-wave synthesis (k_synth_wsq.cuh) available for this shape
    size_t ws_smem = 0;
    int num_sms = 148;
    int synth_nt_fx = 0, synth_nt_dfx = 0, synth_nt_ke = 0, synth_stage_fx = 0, synth_stage_dfx = 0, synth_stage_ke = 0;
    size_t synth_smem_fx = 0, synth_smem_dfx = 0, synth_smem_ke = 0;
    int ana_nt = 0, ana_stage = 0;
    size_t ana_smem = 0;
    size_t solve_smem = 0, solve_hot_smem = 0, solve_gath_smem = 0;
    bool solve_gath = false;   // the hot back-substitution gathers the nonlinear term from spec4 (no post_kernel launch)
    int solve_nsl = 3, solve_hot_nsl = 3;
    double dt_psi = 0, dt_T = 0, dt_S = 0;  // effective time steps of the three operator stacks
    // optional per-stage CUDA-event timing (sddc_profile_begin / sddc_profile_end)
    bool profiling = false;
    struct Ev { int stage; cudaEvent_t a, b; };
    std::vector<Ev> events;
};

namespace {
struct StageTimer {
    sddc_plan* pl; cudaStream_t st; int idx = -1;
    StageTimer(sddc_plan* p, int stage, cudaStream_t s) : pl(p), st(s) {
        if (!pl->profiling) return;
        sddc_plan::Ev e{stage, nullptr, nullptr};
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        pl->events.push_back(e);
        idx = (int)pl->events.size() - 1;
    }
    ~StageTimer() { if (idx >= 0) cudaEventRecord(pl->events[idx].b, st); }
};
}  // namespace

namespace {

#define PLAN_CUDA(plan, expr)                                                                          \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            (plan)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                         \
            return SDDC_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

// Launch with programmatic stream serialization (common.cuh, pdl_*): the kernels of a member-step overlap their
// prologues with the tail of their predecessor.
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

int dev_alloc(sddc_plan* pl, double** out, size_t count, bool zero) {
    void* p = nullptr;
    PLAN_CUDA(pl, cudaMalloc(&p, count * sizeof(double)));
    pl->allocs.push_back(p);
    if (zero) PLAN_CUDA(pl, cudaMemset(p, 0, count * sizeof(double)));
    *out = static_cast<double*>(p);
    return SDDC_OK;
}

int upload(sddc_plan* pl, double** out, const std::vector<double>& h) {
    int rc = dev_alloc(pl, out, h.size(), false);
    if (rc) return rc;
    PLAN_CUDA(pl, cudaMemcpy(*out, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return SDDC_OK;
}

// M[i][i'] (n x n) -> M^T padded: out[i'][i], row length n8
std::vector<double> transpose_pad(const double* M, int n, int n8) {
    std::vector<double> o((size_t)n * n8, 0.0);
    for (int i = 0; i < n; ++i)
        for (int ip = 0; ip < n; ++ip) o[(size_t)ip * n8 + i] = M[(size_t)i * n + ip];
    return o;
}

// stack of nmat (n x n) matrices -> [nmat][n8][LDL], zero padded
std::vector<double> pad_stack(const double* M, int nmat, int n, int n8, int LDL) {
    std::vector<double> o((size_t)nmat * n8 * LDL, 0.0);
    for (int m = 0; m < nmat; ++m)
        for (int i = 0; i < n; ++i)
            std::memcpy(&o[((size_t)m * n8 + i) * LDL], &M[((size_t)m * n + i) * n], sizeof(double) * n);
    return o;
}

// A4 operator stack [K][2][n8][LDL]: for every mode the pre-inverted operator L_inv_j and the product L_inv_j @ D2
// (k_solve.cuh), zero padded
std::vector<double> build_a4_stack(const double* Linv, const double* D2, int K, int n, int n8, int LDL) {
    std::vector<double> o((size_t)K * 2 * n8 * LDL, 0.0);
    std::vector<double> P((size_t)n * n);
    for (int m = 0; m < K; ++m) {
        const double* L = Linv + (size_t)m * n * n;
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < n; ++c) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc = std::fma(L[(size_t)i * n + k], D2[(size_t)k * n + c], acc);
                P[(size_t)i * n + c] = acc;
            }
        for (int i = 0; i < n; ++i) {
            std::memcpy(&o[(((size_t)m * 2 + 0) * n8 + i) * LDL], &L[(size_t)i * n], sizeof(double) * n);
            std::memcpy(&o[(((size_t)m * 2 + 1) * n8 + i) * LDL], &P[(size_t)i * n], sizeof(double) * n);
        }
    }
    return o;
}

// The opt-in limit is a property of the kernel function, shared by every plan in the process: always raise it to
// the architectural maximum so that plans of different shapes cannot lower each other's limit.
template <typename Kern>
int set_smem(sddc_plan* pl, Kern kern, size_t bytes) {
    if (bytes > SMEM_LIMIT) { pl->err = "kernel needs more shared memory than an SM has"; return SDDC_ERR_UNSUPPORTED; }
    PLAN_CUDA(pl, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    return SDDC_OK;
}

// stages and dynamic shared memory of the synthesis kernel for `nset` coefficient sets of `nf` fields
int pick_synth(const Geo& g, int nset, int nf, int* nt, int* nstage, size_t* smem) {
    const int rs = nf * g.n8, mtw = g.nt8 * nset;
    if (mtw > 16) return SDDC_ERR_UNSUPPORTED;
    *nt = synth_nt_for(mtw);
    if (g.n8 * (*nt) * 8 > 4 * 32 * (2 * nf + 1) && nf == 9) return SDDC_ERR_UNSUPPORTED;
    const size_t st = synth_stage_doubles(nset, rs, *nt) * sizeof(double);
    const size_t ep = synth_epi_doubles(nset, rs, g.n, *nt) * sizeof(double);
    int ns = SYNTH_MAX_STAGES;
    while (ns > 2 && ns * st > SMEM_LIMIT) --ns;
    if (ns * st > SMEM_LIMIT || ep > SMEM_LIMIT) return SDDC_ERR_UNSUPPORTED;
    *nstage = ns;
    *smem = std::max(ns * st, ep);
    return SDDC_OK;
}

template <int NT8, int EPI>
void launch_synth_inst(const SynthParams& sp, int nstage, size_t smem, dim3 grid, cudaStream_t st, bool set_attr) {
    constexpr int NF = (EPI == EPI_KE) ? 2 : 9;
    if (set_attr) {
        cudaFuncSetAttribute(synth_kernel<NT8, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        return;
    }
    synth_kernel<NT8, EPI><<<grid, 32 * (2 * NF + 1), smem, st>>>(sp, nstage);
}

template <int EPI>
int launch_synth(sddc_plan* pl, const SynthParams& sp, int nstage, size_t smem, int ncol_tiles, int B,
                 cudaStream_t st, bool set_attr = false) {
    dim3 grid(ncol_tiles, B);
    switch (pl->g.nt8) {
        case 3: launch_synth_inst<3, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        case 4: launch_synth_inst<4, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        case 5: launch_synth_inst<5, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        case 6: launch_synth_inst<6, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        case 7: launch_synth_inst<7, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        case 8: launch_synth_inst<8, EPI>(sp, nstage, smem, grid, st, set_attr); break;
        default: pl->err = "unsupported radial tile count"; return SDDC_ERR_UNSUPPORTED;
    }
    if (!set_attr) pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

template <int NT8>
void launch_ana_inst(const AnaParams& ap, int nstage, size_t smem, dim3 grid, cudaStream_t st, bool set_attr) {
    if (set_attr) {
        cudaFuncSetAttribute(analysis_kernel<NT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        return;
    }
    analysis_kernel<NT8><<<grid, 416, smem, st>>>(ap, nstage);
}

int launch_analysis(sddc_plan* pl, const AnaParams& ap, int B, cudaStream_t st, bool set_attr = false) {
    const int KT3 = 32 * ana_nt_for(pl->g.nt8);
    dim3 grid(pl->g.Khp2 / KT3, 2, B);
    switch (pl->g.nt8) {
        case 3: launch_ana_inst<3>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        case 4: launch_ana_inst<4>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        case 5: launch_ana_inst<5>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        case 6: launch_ana_inst<6>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        case 7: launch_ana_inst<7>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        case 8: launch_ana_inst<8>(ap, pl->ana_stage, pl->ana_smem, grid, st, set_attr); break;
        default: pl->err = "unsupported radial tile count"; return SDDC_ERR_UNSUPPORTED;
    }
    if (!set_attr) pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

int check_batch(sddc_plan* pl, int B) {
    if (!pl) return SDDC_ERR_INVALID;
    if (B < 1 || B > pl->cfg.max_batch) {
        pl->err = "batch size " + std::to_string(B) + " outside [1, max_batch=" + std::to_string(pl->cfg.max_batch) + "]";
        return SDDC_ERR_INVALID;
    }
    PLAN_CUDA(pl, cudaSetDevice(pl->device));
    return SDDC_OK;
}

int run_scan(sddc_plan* pl, const double* X, long long stride, int B, cudaStream_t st) {
    const long long tot = (long long)B * 2 * pl->g.n;
    StageTimer tm(pl, SDDC_STAGE_SCAN, st);
    scan_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(X, stride, pl->JJ, pl->g, B);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

// scan + prep of state X into coefficient set `set` (and optionally the linear right-hand side)
int run_prep(sddc_plan* pl, const double* X, int set, bool want_coef, double* lin, const double* Ra,
             const double* Ras, int B, cudaStream_t st, double* coef7 = nullptr, bool have_jj = false) {
    // have_jj: pl->JJ already holds the brackets of X (written by the previous step's back-substitution)
    int rc = have_jj ? SDDC_OK : run_scan(pl, X, 3LL * pl->g.N, B, st);
    if (rc) return rc;
    PrepParams pp{};
    pp.X = X; pp.x_stride = 3LL * pl->g.N; pp.JJ = pl->JJ;
    const bool fftl = coef7 != nullptr;  // seven spectral rows per radial point for the FFT formulation
    pp.coef = want_coef ? (fftl ? coef7 : (set == 0 ? pl->coef : pl->coef1)) : nullptr;
    pp.coef_stride = pl->coef_member_stride;
    pp.lin = lin; pp.bstride = pl->bstride; pp.Ra = Ra; pp.Ras = Ras;
    pp.DrP = pl->DrP; pp.D2rP = pl->D2rP; pp.DsqP = pl->DsqP;
    pp.ir2 = pl->ir2; pp.ir4 = pl->ir4; pp.r2 = pl->r2; pp.dT0 = pl->dT0; pp.gb = pl->gb;
    pp.g = pl->g; pp.B = B;
    const int ntiles = ((pl->g.K + PREP_TC - 1) / PREP_TC) * B;
    pp.nstage = (pl->g.nt8 > 4 && prep_smem_bytes(pl->g.n8, 2) <= SMEM_LIMIT) ? 2 : 1;
    const size_t smem = prep_smem_bytes(pl->g.n8, pp.nstage);
    const int grid = pl->g.nt8 <= 4 ? ntiles : std::min(ntiles, pl->num_sms);   // one CTA per tile, or persistent
    StageTimer tm(pl, SDDC_STAGE_PREP, st);
    if (fftl) {
        switch (pl->g.nt8) {
            case 3: launch_pdl(prep_kernel<3, true>, dim3(grid), dim3(192), smem, st, pp, ntiles); break;
            case 4: launch_pdl(prep_kernel<4, true>, dim3(grid), dim3(256), smem, st, pp, ntiles); break;
            case 5: launch_pdl(prep_kernel<5, true>, dim3(grid), dim3(320), smem, st, pp, ntiles); break;
            case 6: launch_pdl(prep_kernel<6, true>, dim3(grid), dim3(384), smem, st, pp, ntiles); break;
            case 7: launch_pdl(prep_kernel<7, true>, dim3(grid), dim3(448), smem, st, pp, ntiles); break;
            default: launch_pdl(prep_kernel<8, true>, dim3(grid), dim3(512), smem, st, pp, ntiles); break;
        }
    } else {
        switch (pl->g.nt8) {
            case 3: launch_pdl(prep_kernel<3, false>, dim3(grid), dim3(192), smem, st, pp, ntiles); break;
            case 4: launch_pdl(prep_kernel<4, false>, dim3(grid), dim3(256), smem, st, pp, ntiles); break;
            case 5: launch_pdl(prep_kernel<5, false>, dim3(grid), dim3(320), smem, st, pp, ntiles); break;
            case 6: launch_pdl(prep_kernel<6, false>, dim3(grid), dim3(384), smem, st, pp, ntiles); break;
            case 7: launch_pdl(prep_kernel<7, false>, dim3(grid), dim3(448), smem, st, pp, ntiles); break;
            default: launch_pdl(prep_kernel<8, false>, dim3(grid), dim3(512), smem, st, pp, ntiles); break;
        }
    }
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

int run_synth_nl(sddc_plan* pl, bool dfx, int B, cudaStream_t st) {
    if (dfx && !pl->dfx_ok) {
        pl->err = "NLIN_DFX / JVP is not available for this N_r (two-state synthesis exceeds shared memory; N_r <= 41)";
        return SDDC_ERR_UNSUPPORTED;
    }
    SynthParams sp{};
    sp.coef0 = pl->coef; sp.coef1 = pl->coef1; sp.coef_stride = pl->coef_member_stride;
    sp.tab = dfx ? pl->tab1d : pl->tab1; sp.Dr = pl->Dr; sp.prd = pl->prd;
    sp.g = pl->g;
    const int nt = dfx ? pl->synth_nt_dfx : pl->synth_nt_fx;
    const int tiles = pl->g.Mhp / (8 * nt);
    StageTimer tm(pl, SDDC_STAGE_SYNTH, st);
    if (!dfx && pl->ws_ok) {
        sp.tab = pl->tab1q;
        const int ntj = pl->g.Mhp / SWS_W, nwork = ntj * B;
        const int grid = std::min(nwork, pl->num_sms);
        synth_wsq_kernel<4, SWS_FX><<<grid, SWS_NTHR, pl->ws_smem, st>>>(sp, ntj, nwork);
        pl->launches++;
        PLAN_CUDA(pl, cudaGetLastError());
        return SDDC_OK;
    }
    if (dfx) return launch_synth<EPI_DFX>(pl, sp, pl->synth_stage_dfx, pl->synth_smem_dfx, tiles, B, st);
    return launch_synth<EPI_FX>(pl, sp, pl->synth_stage_fx, pl->synth_smem_fx, tiles, B, st);
}

int run_analysis(sddc_plan* pl, double* out, bool solve_major, int B, cudaStream_t st, bool quarter) {
    AnaParams ap{};
    ap.prd = pl->prd; ap.tab2 = quarter ? pl->tab2q : pl->tab2; ap.quarter = quarter ? 1 : 0;
    ap.out = out; ap.bstride = solve_major ? pl->bstride : 0; ap.g = pl->g;
    StageTimer tm(pl, SDDC_STAGE_ANALYSIS, st);
    return launch_analysis(pl, ap, B, st);
}

// FFT formulation: coefficient rows -> analysed products spec4 (k_nlin_fft.cuh), then post_kernel
// workers per CTA: 8 plane pairs... the one-state worker holds 4 plane pairs (8 M doubles), the two-state one 7
template <int M>
constexpr int nlin_fft_nw(bool dfx) { return dfx ? (M == 768 ? 2 : (M == 384 ? 5 : 8)) : (M == 768 ? 4 : NLIN_FFT_NW); }

// mode: 0 products of c0 (and c1: two-state kernel); 1 grid fields of c0 -> pl->grid7; 2 products of c1 with the cached grid
template <int M>
int launch_nlin_fft(sddc_plan* pl, NlinFftParams& np, double* out, long long bstride, bool dfx, cudaStream_t st, bool set_attr,
                    int mode = 0) {
    constexpr int NW = nlin_fft_nw<M>(false), NWD = nlin_fft_nw<M>(true);
#ifndef NLIN_JVPC_NW768
#define NLIN_JVPC_NW768 3   // 4 workers (512 threads) cap the registers at 128: 360 bytes of spills, 0.71 ms against 0.61 ms at (40,512)
#endif
    constexpr int NWJ = M == 768 ? NLIN_JVPC_NW768 : NW;   // workers of the cached-base product kernel
    constexpr int NT = M == 768 ? 128 : 64;   // threads per worker = columns of the radix-6 pass at M = 768
    constexpr size_t smem = nlin_fft_smem_bytes<M, false>(NW), smem_d = nlin_fft_smem_bytes<M, true>(NWD);
    static_assert(smem <= SMEM_LIMIT && smem_d <= SMEM_LIMIT, "workers do not fit into shared memory");
    if (set_attr) {
        PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_kernel<M, false, NW, NT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_kernel<M, true, NWD, NT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_kernel<M, false, NW, NT, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_kernel<M, false, NWJ, NT, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        if (M == 384) {
            PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_staged_kernel<384, NLIN_FFT_STAGED_NW>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
            PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_staged_kernel<384, NLIN_FFT_STAGED_NW, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
            PLAN_CUDA(pl, cudaFuncSetAttribute((nlin_fft_staged_kernel<384, NLIN_FFT_STAGED_NW, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        }
        return SDDC_OK;
    }
    const int n = pl->g.n, n8 = pl->g.n8;
    np.next_row = pl->fft_row;
    {
        StageTimer tm(pl, SDDC_STAGE_SYNTH, st);
        if (dfx && mode == 0) {
            const int grid = std::min((np.nrows + NWD - 1) / NWD, pl->num_sms);
            launch_pdl(nlin_fft_kernel<M, true, NWD, NT, 0>, dim3(grid), dim3(NT * NWD), smem_d, st, np);
        } else if (M == 384) {
            // headline shape: per-warp ownership of the transforms, coefficient rows staged by TMA (k_nlin_fft.cuh)
            constexpr int NWS = NLIN_FFT_STAGED_NW;
            static_assert(nlin_fft_staged_smem_bytes<384>(NWS) <= SMEM_LIMIT, "staged workers do not fit into shared memory");
            const int grid = std::min((np.nrows + NWS - 1) / NWS, pl->num_sms);
            const size_t ssm = nlin_fft_staged_smem_bytes<384>(NWS);
            if (mode == 1) launch_pdl(nlin_fft_staged_kernel<384, NWS, 1>, dim3(grid), dim3(64 * NWS), ssm, st, np);
            else if (mode == 2) launch_pdl(nlin_fft_staged_kernel<384, NWS, 2>, dim3(grid), dim3(64 * NWS), ssm, st, np);
            else launch_pdl(nlin_fft_staged_kernel<384, NWS, 0>, dim3(grid), dim3(64 * NWS), ssm, st, np);
        } else {
            const int grid = std::min((np.nrows + NW - 1) / NW, pl->num_sms);
            if (mode == 1) launch_pdl(nlin_fft_kernel<M, false, NW, NT, 1>, dim3(grid), dim3(NT * NW), smem, st, np);
            else if (mode == 2) launch_pdl(nlin_fft_kernel<M, false, NWJ, NT, 2>, dim3(std::min((np.nrows + NWJ - 1) / NWJ, pl->num_sms)), dim3(NT * NWJ), nlin_fft_smem_bytes<M, false>(NWJ), st, np);
            else launch_pdl(nlin_fft_kernel<M, false, NW, NT, 0>, dim3(grid), dim3(NT * NW), smem, st, np);
        }
    }
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    if (out) {
        PostParams pp{};
        pp.spec = pl->spec4; pp.DrT = pl->DrT; pp.out = out; pp.bstride = bstride; pp.g = pl->g;
        const int ntiles = ((pl->g.K + POST_TC - 1) / POST_TC) * (np.nrows / n);
        const size_t psm = post_smem_bytes(n, n8);
        const int per_sm = std::max(1, std::min(4, (int)((SMEM_LIMIT + 1024) / (psm + 1024))));
        StageTimer tm(pl, SDDC_STAGE_ANALYSIS, st);
        launch_pdl(post_kernel, dim3(std::min(ntiles, per_sm * pl->num_sms)), dim3(256), psm, st, pp, ntiles);
        pl->launches++;
        PLAN_CUDA(pl, cudaGetLastError());
    }
    return SDDC_OK;
}

// nonlinear term of the rows c0 (and, two-state, c1) into `out` (state layout or solve-major)
int run_nlin_fft(sddc_plan* pl, const double* c0, const double* c1, double* out, bool solve_major, int B, cudaStream_t st,
                 bool set_attr = false, int mode = 0) {
    NlinFftParams np{};
    np.coef0 = c0; np.coef1 = c1; np.spec = pl->spec4; np.tab = pl->fft_tab; np.nrows = B * pl->g.n; np.grid = pl->grid7;
    const long long bstride = solve_major ? pl->bstride : 0;
    const bool dfx = c1 != nullptr;
    if (pl->fft_direct || (dfx && pl->dfx_direct)) {
        if (set_attr) return SDDC_OK;
        NlinDirectParams dp{};
        dp.coef0 = c0; dp.coef1 = c1; dp.spec = pl->spec4; dp.K = pl->g.K; dp.M = pl->g.M;
        const size_t dsm = sizeof(double) * ((size_t)(dfx ? 14 : 7) * pl->g.K + 4 * (size_t)pl->g.M);
        {
            StageTimer tm(pl, SDDC_STAGE_SYNTH, st);
            nlin_direct_kernel<<<np.nrows, 128, dsm, st>>>(dp);
        }
        pl->launches++;
        PLAN_CUDA(pl, cudaGetLastError());
        if (out) {
            PostParams pp{};
            pp.spec = pl->spec4; pp.DrT = pl->DrT; pp.out = out; pp.bstride = bstride; pp.g = pl->g;
            const int ntiles = ((pl->g.K + POST_TC - 1) / POST_TC) * B;
            const size_t psm = post_smem_bytes(pl->g.n, pl->g.n8);
            const int per_sm = std::max(1, std::min(4, (int)((SMEM_LIMIT + 1024) / (psm + 1024))));
            StageTimer tm(pl, SDDC_STAGE_ANALYSIS, st);
            launch_pdl(post_kernel, dim3(std::min(ntiles, per_sm * pl->num_sms)), dim3(256), psm, st, pp, ntiles);
            pl->launches++;
            PLAN_CUDA(pl, cudaGetLastError());
        }
        return SDDC_OK;
    }
    switch (pl->fft_M) {
        case 192: return launch_nlin_fft<192>(pl, np, out, bstride, dfx, st, set_attr, mode);
        case 384: return launch_nlin_fft<384>(pl, np, out, bstride, dfx, st, set_attr, mode);
        case 768: return launch_nlin_fft<768>(pl, np, out, bstride, dfx, st, set_attr, mode);
        default: pl->err = "FFT path not available for this N_fm"; return SDDC_ERR_UNSUPPORTED;
    }
}

// gs < 0 selects the solve-major layout for g / fnl
int run_solve(sddc_plan* pl, const double* g, const double* fnl, long long gs, long long gf, double* out, long long os,
              long long of, const double* sub, int field_base, int nfields, int B, cudaStream_t st, double* jj_out = nullptr,
              double* dpart = nullptr, const double* spec = nullptr) {
    SolveParams sp{};
    sp.spec = spec; sp.DrT = pl->DrT;
    sp.jj_out = jj_out;
    sp.dpart = dpart; sp.nu_in = pl->nu_in; sp.nu_out = pl->nu_out; sp.nu_w = pl->nu_w;
    const bool sm = gs < 0;
    sp.bstride = pl->bstride;
    sp.g = g; sp.fnl = fnl; sp.mdt = -pl->g.dt; sp.g_stride = gs; sp.g_field_off = gf; sp.out = out; sp.out_stride = os; sp.out_field_off = of;
    sp.sub = sub; sp.LinvA4 = pl->LA4; sp.LinvT = pl->LT; sp.LinvS = pl->LS; sp.D2 = pl->D2p;
    sp.ir2 = pl->a4_ir2; sp.ir4 = pl->a4_ir4; sp.geo = pl->g; sp.B = B;
    sp.field_mask = 7; sp.field_base = field_base;
    sp.dt_psi = pl->dt_psi; sp.dt_T = pl->dt_T; sp.dt_S = pl->dt_S; sp.nsl = pl->solve_nsl;
    // single-field calls pass field offsets of 0; the operator stack follows field_base
    dim3 grid((B + 8 * SOLVE_NTB - 1) / (8 * SOLVE_NTB), 2, nfields);
    StageTimer tm(pl, SDDC_STAGE_SOLVE, st);
    if (sm) {
        // hot path: all three fields from the solve-major buffers (k_solve_hot.cuh)
        const int npsi = (B + 8 * SOLVE_NTB_PSI - 1) / (8 * SOLVE_NTB_PSI), nts = (B + 8 * SOLVE_NTB_TS - 1) / (8 * SOLVE_NTB_TS);
        const int nthr = 32 * (pl->g.nt8 + 1);
        const size_t smb = spec ? pl->solve_gath_smem : pl->solve_hot_smem;
        const bool n3 = pl->solve_hot_nsl == 3;
        const int nblk = 2 * npsi + 4 * nts;
        if (sub && dpart) { pl->err = "diagnostics partial sums are only produced by plain steps"; return SDDC_ERR_INVALID; }
        if (spec) {
            if (!pl->solve_gath || fnl) { pl->err = "gather mode of the back-substitution is not available for this plan"; return SDDC_ERR_INVALID; }
#define SDDC_LAUNCH_SOLVE_GATH(NT)                                                                                  \
    if (sub) launch_pdl(solve_hot_kernel<NT, 3, true, false, true>, dim3(nblk), dim3(nthr), smb, st, sp, npsi);   \
    else if (dpart) launch_pdl(solve_hot_kernel<NT, 3, false, true, true>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); \
    else launch_pdl(solve_hot_kernel<NT, 3, false, false, true>, dim3(nblk), dim3(nthr), smb, st, sp, npsi);
            const int nthr = 32 * (pl->g.nt8 + 2);   // + the gather warp
            if (pl->g.nt8 == 3) { SDDC_LAUNCH_SOLVE_GATH(3) } else { SDDC_LAUNCH_SOLVE_GATH(4) }
#undef SDDC_LAUNCH_SOLVE_GATH
            pl->launches++;
            PLAN_CUDA(pl, cudaGetLastError());
            return SDDC_OK;
        }
#define SDDC_LAUNCH_SOLVE_HOT(NT)                                                                                 \
    if (sub) { if (n3) launch_pdl(solve_hot_kernel<NT, 3, true, false>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); else launch_pdl(solve_hot_kernel<NT, 2, true, false>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); } \
    else if (dpart) { if (n3) launch_pdl(solve_hot_kernel<NT, 3, false, true>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); else launch_pdl(solve_hot_kernel<NT, 2, false, true>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); } \
    else { if (n3) launch_pdl(solve_hot_kernel<NT, 3, false, false>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); else launch_pdl(solve_hot_kernel<NT, 2, false, false>, dim3(nblk), dim3(nthr), smb, st, sp, npsi); }
        switch (pl->g.nt8) {
            case 3: SDDC_LAUNCH_SOLVE_HOT(3) break;
            case 4: SDDC_LAUNCH_SOLVE_HOT(4) break;
            case 5: SDDC_LAUNCH_SOLVE_HOT(5) break;
            case 6: SDDC_LAUNCH_SOLVE_HOT(6) break;
            case 7: SDDC_LAUNCH_SOLVE_HOT(7) break;
            case 8: SDDC_LAUNCH_SOLVE_HOT(8) break;
            default: pl->err = "unsupported radial tile count"; return SDDC_ERR_UNSUPPORTED;
        }
#undef SDDC_LAUNCH_SOLVE_HOT
    } else solve_kernel<SOLVE_NTB, false><<<grid, 32 * (pl->g.nt8 + 1), pl->solve_smem, st>>>(sp);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

// row kernels on (c0, c1, mode) followed by the hot back-substitution: through the gather mode where the plan has it (the
// chain reads the analysed products itself), else through post_kernel and the solve-major F(X) buffer.
// The gather mode trades a throughput-bound kernel (post_kernel: 0.075 ms per 512 members at (30,256)) for a longer chain
// step of the latency-bound one (0.099 -> 0.134 ms at 512 members, 0.080 -> 0.103 ms for a few members): it pays from
// 128 members on (tools/batch_scaling.py on both builds; 64 concurrent Newton solves ran 20 % slower
// through it).
#ifndef SOLVE_GATH_MIN_MEMBERS
#define SOLVE_GATH_MIN_MEMBERS 128   // compile-time only (tools/batch_scaling.py compares builds)
#endif
constexpr int SOLVE_GATH_MIN_B = SOLVE_GATH_MIN_MEMBERS;
int run_rows_and_solve(sddc_plan* pl, const double* c0, const double* c1, int mode, double* out, const double* sub, int B,
                       cudaStream_t st, double* jj_out = nullptr, double* dpart = nullptr) {
    const long long N3 = 3LL * pl->g.N;
    int rc;
    if (pl->solve_gath && B >= SOLVE_GATH_MIN_B) {
        if ((rc = run_nlin_fft(pl, c0, c1, nullptr, true, B, st, false, mode))) return rc;
        return run_solve(pl, pl->lin_sm, nullptr, -1, 0, out, N3, pl->g.N, sub, 0, 3, B, st, jj_out, dpart, pl->spec4);
    }
    if ((rc = run_nlin_fft(pl, c0, c1, pl->f_sm, true, B, st, false, mode))) return rc;
    return run_solve(pl, pl->lin_sm, pl->f_sm, -1, 0, out, N3, pl->g.N, sub, 0, 3, B, st, jj_out, dpart);
}

// kinetic energy by FFT from coefficient rows + the remaining diagnostics (norm, Nusselt numbers)
int run_ke_fft(sddc_plan* pl, const double* X, const double* rows, long long row_stride, int b_off, const double* ascale,
               double* out, int B, cudaStream_t st, const double* dpart = nullptr, bool side = false) {
    const Geo& g = pl->g;
    KeFftParams kf{};
    kf.rows = rows; kf.row_stride = row_stride; kf.b_off = b_off; kf.ascale = ascale;
    kf.tab = pl->ke_tab; kf.Wn = pl->ke_Wn; kf.wr = pl->wr; kf.kepart = pl->kepart;
    kf.nrows = B * g.n; kf.n = g.n;
    {
        StageTimer tm(pl, SDDC_STAGE_KE_SYNTH, st);
        if (side && pl->ke_M == 768) ke_fft_kernel<768, 2><<<std::min((kf.nrows + 1) / 2, 3 * pl->num_sms), 128, ke_fft_smem_bytes<768>(2), st>>>(kf);
        else if (pl->ke_M == 384) ke_fft_kernel<384, 8><<<std::min((kf.nrows + 7) / 8, pl->num_sms), 512, ke_fft_smem_bytes<384>(8), st>>>(kf);
        else if (pl->ke_M == 1536) ke_fft_kernel<1536, KE_NW1536><<<std::min((kf.nrows + KE_NW1536 - 1) / KE_NW1536, pl->num_sms), 64 * KE_NW1536, ke_fft_smem_bytes<1536>(KE_NW1536), st>>>(kf);
        else ke_fft_kernel<768, KE_NW768><<<std::min((kf.nrows + KE_NW768 - 1) / KE_NW768, pl->num_sms), 64 * KE_NW768, ke_fft_smem_bytes<768>(KE_NW768), st>>>(kf);
    }
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    StageTimer tm2(pl, SDDC_STAGE_DIAG, st);
    // norm and Nusselt numbers: from the partial sums the back-substitution that produced X left behind, else from X
    if (dpart) diag_finish_kernel<<<(B + 3) / 4, 128, 0, st>>>(dpart, pl->kepart, g.n, pl->ke_scale, B, out);
    else diag_kernel<<<B, 256, 0, st>>>(X, pl->kepart, g.n, pl->nu_in, pl->nu_out, pl->ke_scale, g, out);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

// Step(X) [- sub]: the shared composition of Step_Python / PFX (Main.py:255-283, 473-496)
// have_jj: the previous member-step of this call left the theta-coupling brackets of X in pl->JJ (emit_jj below)
int run_step_prep(sddc_plan* pl, const double* X, const double* Ra, const double* Ras, int B, bool linear, cudaStream_t st,
                  bool have_jj = false) {
    const bool fft = pl->fft_M != 0 && !linear;
    return run_prep(pl, X, 0, !linear, pl->lin_sm, Ra, Ras, B, st, fft ? pl->coef7 : nullptr, have_jj);
}

int run_step_rest(sddc_plan* pl, double* out, const double* sub, int B, bool linear, cudaStream_t st, bool emit_jj = false,
                  double* emit_diag = nullptr) {
    const long long N3 = 3LL * pl->g.N;
    const bool fft = pl->fft_M != 0 && !linear;
    int rc;
    const double* fnl = nullptr;
    if (fft) return run_rows_and_solve(pl, pl->coef7, nullptr, 0, out, sub, B, st, emit_jj ? pl->JJ : nullptr, emit_diag);
    if (!linear) {
        if ((rc = run_synth_nl(pl, false, B, st))) return rc;
        if ((rc = run_analysis(pl, pl->f_sm, true, B, st, pl->quarter))) return rc;
        fnl = pl->f_sm;  // F(X); the solve kernel forms lin - dt * F
    }
    return run_solve(pl, pl->lin_sm, fnl, -1, 0, out, N3, pl->g.N, sub, 0, 3, B, st, emit_jj ? pl->JJ : nullptr,
                     emit_diag);
}

// have_jj / emit_jj chain the steps of a multi-step call: the A4 back-substitution of step s writes the suffix-sum
// brackets of its output, step s+1 starts without its scan launch
int run_member_step(sddc_plan* pl, const double* X, double* out, const double* sub, const double* Ra,
                    const double* Ras, int B, bool linear, cudaStream_t st, bool have_jj = false, bool emit_jj = false) {
    int rc = run_step_prep(pl, X, Ra, Ras, B, linear, st, have_jj);
    if (rc) return rc;
    return run_step_rest(pl, out, sub, B, linear, st, emit_jj);
}

int ensure_host_staging(sddc_plan* pl) {
    if (pl->hX0) return SDDC_OK;
    const size_t cnt = (size_t)pl->cfg.max_batch * 3 * pl->g.N;
    int rc;
    if ((rc = dev_alloc(pl, &pl->hX0, cnt, false))) return rc;
    if ((rc = dev_alloc(pl, &pl->hX1, cnt, false))) return rc;
    if ((rc = dev_alloc(pl, &pl->hX2, cnt, false))) return rc;
    if ((rc = dev_alloc(pl, &pl->hRa, pl->cfg.max_batch, false))) return rc;
    if ((rc = dev_alloc(pl, &pl->hRas, pl->cfg.max_batch, false))) return rc;
    if ((rc = dev_alloc(pl, &pl->hDiag, (size_t)pl->cfg.max_batch * 6, false))) return rc;
    PLAN_CUDA(pl, cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking));
    PLAN_CUDA(pl, cudaStreamCreateWithFlags(&pl->in_stream, cudaStreamNonBlocking));
    PLAN_CUDA(pl, cudaStreamCreateWithFlags(&pl->out_stream, cudaStreamNonBlocking));
    return SDDC_OK;
}

}  // namespace

extern "C" {

int sddc_version(void) { return 100; }

int sddc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* sddc_last_error(const sddc_plan* plan) { return plan ? plan->err.c_str() : g_create_error.c_str(); }

long long sddc_launch_count(const sddc_plan* plan) { return plan ? plan->launches : 0; }

int sddc_plan_info(const sddc_plan* plan, int what) {
    if (!plan) return -1;
    switch (what) {
        case 0: return plan->quarter ? 1 : 0;       // second mirror level active on the hot path
        case 1: return plan->ws_ok ? 1 : 0;         // dense path: 0 generic synthesis kernel, 1 persistent warp-specialised
        case 2: return plan->dfx_ok ? 1 : 0;        // two-state JVP synthesis available
        case 3: return plan->g.n8;
        case 4: return plan->fft_direct ? 0 : plan->fft_M;   // grid size of the FFT formulation of the nonlinear term (0: not active)
        case 5: return plan->fft_dfx ? 1 : 0;       // FFT formulation also used for the two-state (JVP) products
        case 6: return plan->ke_M;                  // grid size of the kinetic-energy FFT (0: dense synthesis)
        case 7: return plan->fft_direct ? 1 : (plan->dfx_direct ? 2 : 0);   // direct-summation row kernel: 1 every product, 2 two-state products only
        case 8: return plan->solve_gath ? 1 : 0;    // hot back-substitution gathers the analysed products itself (no post_kernel)
        default: return -1;
    }
}

void sddc_plan_destroy(sddc_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    for (void* p : plan->allocs) cudaFree(p);
    if (plan->hHist) cudaFree(plan->hHist);
    if (plan->own_stream) cudaStreamDestroy(plan->own_stream);
    if (plan->in_stream) cudaStreamDestroy(plan->in_stream);
    if (plan->out_stream) cudaStreamDestroy(plan->out_stream);
    if (plan->ev_ckpt) cudaEventDestroy(plan->ev_ckpt);
    if (plan->ke_stream) cudaStreamDestroy(plan->ke_stream);
    if (plan->ev_rows) cudaEventDestroy(plan->ev_rows);
    if (plan->ev_ke) cudaEventDestroy(plan->ev_ke);
    for (auto e : plan->ev_in) cudaEventDestroy(e);
    for (auto e : plan->ev_done) cudaEventDestroy(e);
    delete plan;
}

int sddc_plan_create(sddc_plan** out, const sddc_config* cfg, const sddc_operators* ops) {
    if (!out || !cfg || !ops) { g_create_error = "null argument"; return SDDC_ERR_INVALID; }
    *out = nullptr;
    if (cfg->N_fm % 2 != 0) {
        // same condition the reference raises on (Matrix_Operators.py:758-759)
        g_create_error = "The number of Fourier modes is not even " + std::to_string(cfg->N_fm);
        return SDDC_ERR_INVALID;
    }
    if (cfg->N_fm < 8) { g_create_error = "N_fm must be >= 8"; return SDDC_ERR_UNSUPPORTED; }
    if (cfg->N_r < 4 || cfg->N_r - 1 > 64) { g_create_error = "N_r must satisfy 4 <= N_r <= 65"; return SDDC_ERR_UNSUPPORTED; }
    if (cfg->max_batch < 1) { g_create_error = "max_batch must be >= 1"; return SDDC_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_error = "no CUDA device"; return SDDC_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return SDDC_ERR_INVALID; }

    sddc_plan* pl = new (std::nothrow) sddc_plan();
    if (!pl) { g_create_error = "out of host memory"; return SDDC_ERR_INVALID; }
    pl->cfg = *cfg;
    pl->device = cfg->device;
    Geo& g = pl->g;
    g.n = cfg->N_r - 1; g.n8 = std::max(24, round_up(g.n, 8)); g.nt8 = g.n8 / 8;  // kernels are instantiated for nt8 = 3..8
    g.K = cfg->N_fm; g.Kh = g.K / 2; g.Khp = round_up(g.Kh, 8); g.Khp2 = round_up(g.Kh, 128);
    g.M = 3 * g.K / 2; g.Mh = g.M / 2; g.Mhp = round_up(g.Mh, 32);
    g.N = g.n * g.K; g.symmetric = cfg->symmetric ? 1 : 0;
    g.dt = cfg->dt; g.Pr = cfg->Pr; g.Tau = cfg->Tau;
    pl->LDL = g.n8 + 4;
    pl->dt_psi = g.Pr * g.dt; pl->dt_T = g.dt; pl->dt_S = g.Tau * g.dt;
    const int n = g.n, n8 = g.n8, K = g.K;
    const int M3 = 3 * K;
    pl->Mh3p = round_up(M3 / 2, 32);

    auto fail = [&](int rc) {
        g_create_error = pl->err;
        sddc_plan_destroy(pl);
        return rc;
    };
#define TRY(x) do { int rc__ = (x); if (rc__) return fail(rc__); } while (0)
#define TRYC(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { pl->err = std::string(#x) + ": " + cudaGetErrorString(e__); return fail(SDDC_ERR_CUDA); } } while (0)

    TRYC(cudaSetDevice(pl->device));
    // ---- operators ----
    TRY(upload(pl, &pl->DrT, transpose_pad(ops->Dr, n, n8)));
    TRY(upload(pl, &pl->D2rT, transpose_pad(ops->D2r, n, n8)));
    TRY(upload(pl, &pl->DsqT, transpose_pad(ops->Dsq, n, n8)));
    TRY(upload(pl, &pl->Dr, std::vector<double>(ops->Dr, ops->Dr + (size_t)n * n)));
    TRY(upload(pl, &pl->D2p, pad_stack(ops->D2, 1, n, n8, pl->LDL)));
    TRY(upload(pl, &pl->DrP, pad_stack(ops->Dr, 1, n, n8, pl->LDL)));
    TRY(upload(pl, &pl->D2rP, pad_stack(ops->D2r, 1, n, n8, pl->LDL)));
    TRY(upload(pl, &pl->DsqP, pad_stack(ops->Dsq, 1, n, n8, pl->LDL)));
    pl->h_LinvA4.assign(ops->Linv_A4, ops->Linv_A4 + (size_t)K * n * n);
    pl->h_D2.assign(ops->D2, ops->D2 + (size_t)n * n);
    TRY(upload(pl, &pl->LA4, build_a4_stack(ops->Linv_A4, ops->D2, K, n, n8, pl->LDL)));
    TRY(upload(pl, &pl->LT, pad_stack(ops->Linv_T, K, n, n8, pl->LDL)));
    TRY(upload(pl, &pl->LS, pad_stack(ops->Linv_S, K, n, n8, pl->LDL)));
    auto vec = [&](const double* v) { return std::vector<double>(v, v + n); };
    TRY(upload(pl, &pl->ir2, vec(ops->ir2)));
    TRY(upload(pl, &pl->ir4, vec(ops->ir4)));
    TRY(upload(pl, &pl->r2, vec(ops->r2)));
    TRY(upload(pl, &pl->dT0, vec(ops->dT0)));
    TRY(upload(pl, &pl->gb, vec(ops->gbuoy)));
    TRY(upload(pl, &pl->a4_ir2, vec(ops->a4_ir2)));
    TRY(upload(pl, &pl->a4_ir4, vec(ops->a4_ir4)));
    TRY(upload(pl, &pl->ir, vec(ops->ir)));
    TRY(upload(pl, &pl->nu_in, vec(ops->nu_in)));
    TRY(upload(pl, &pl->nu_out, vec(ops->nu_out)));
    {
        // np.trapz(.., x=R[1:-1], axis=0): interior nodes only (Main.py:129)
        std::vector<double> wr(n, 0.0);
        for (int i = 0; i + 1 < n; ++i) {
            const double h = 0.5 * (ops->r[i + 1] - ops->r[i]);
            wr[i] += h; wr[i + 1] += h;
        }
        TRY(upload(pl, &pl->wr, wr));
        const double V = (2.0 / 3.0) * (ops->R_out * ops->R_out * ops->R_out - ops->R_in * ops->R_in * ops->R_in);
        pl->ke_scale = 0.5 / V;
    }
    // ---- kernel configuration (tile shapes decide the table layouts) ----
    if (pick_synth(g, 1, 9, &pl->synth_nt_fx, &pl->synth_stage_fx, &pl->synth_smem_fx) ||
        pick_synth(g, 1, 2, &pl->synth_nt_ke, &pl->synth_stage_ke, &pl->synth_smem_ke)) {
        pl->err = "N_r too large for the instantiated synthesis tiles";
        return fail(SDDC_ERR_UNSUPPORTED);
    }
    // the two-state (JVP) synthesis needs twice the staging; beyond N_r = 41 it does not fit in shared memory
    pl->dfx_ok = pick_synth(g, 2, 9, &pl->synth_nt_dfx, &pl->synth_stage_dfx, &pl->synth_smem_dfx) == SDDC_OK;
    if (!pl->dfx_ok) { pl->synth_nt_dfx = pl->synth_nt_fx; pl->synth_smem_dfx = 0; }
    pl->ana_nt = ana_nt_for(g.nt8);
    // ---- trigonometric tables (tile-major, L2-resident) ----
    {
        const int nch = g.Khp / 8, Wfx = 8 * pl->synth_nt_fx, Wd = 8 * pl->synth_nt_dfx, Wke = 8 * pl->synth_nt_ke;
        const int KT3 = 32 * pl->ana_nt;
        const size_t t1 = 4ull * g.Mhp * g.Khp, t2 = 4ull * g.Khp2 * g.Mhp, t3 = 4ull * pl->Mh3p * g.Khp;
        TRY(dev_alloc(pl, &pl->tab1, t1, false));
        TRY(dev_alloc(pl, &pl->tab1d, t1, false));
        TRY(dev_alloc(pl, &pl->tab2, t2, false));
        TRY(dev_alloc(pl, &pl->tab3, t3, false));
        TRY(dev_alloc(pl, &pl->wth, pl->Mh3p, false));
        fill_table_kernel<<<296, 256>>>(pl->tab1, 0, g.M, g.Kh, g.Mh, Wfx, g.Mhp / Wfx, nch);
        fill_table_kernel<<<296, 256>>>(pl->tab1d, 0, g.M, g.Kh, g.Mh, Wd, g.Mhp / Wd, nch);
        fill_table_kernel<<<296, 256>>>(pl->tab2, 1, g.M, g.Kh, g.Mh, KT3, g.Khp2 / KT3, g.Mhp / 8);
        fill_table_kernel<<<296, 256>>>(pl->tab3, 0, M3, g.Kh, M3 / 2, Wke, pl->Mh3p / Wke, nch);
        fill_ke_weights_kernel<<<(pl->Mh3p + 127) / 128, 128>>>(pl->wth, M3, pl->Mh3p);
        pl->launches += 5;
        TRYC(cudaGetLastError());
    }
    // ---- scratch ----
    const size_t Bm = (size_t)cfg->max_batch;
    pl->coef_member_stride = 9LL * n8 * 2 * g.Khp;
    TRY(dev_alloc(pl, &pl->JJ, Bm * (K + 1) * n, false));
    TRY(dev_alloc(pl, &pl->coef, Bm * pl->coef_member_stride, true));   // padded rows / columns stay zero
    TRY(dev_alloc(pl, &pl->coef1, Bm * pl->coef_member_stride, true));  // second set: dv (JVP) or KE rows
    TRY(dev_alloc(pl, &pl->prd, Bm * 3 * 2 * n8 * g.Mhp, true));
    pl->bstride = (long long)round_up(cfg->max_batch, 32);   // a cluster pair of 16-member tiles never leaves its slab
    TRY(dev_alloc(pl, &pl->lin_sm, (size_t)3 * K * pl->bstride * (n8 + SDDC_SM_PAD), true));
    TRY(dev_alloc(pl, &pl->f_sm, (size_t)3 * K * pl->bstride * (n8 + SDDC_SM_PAD), true));
    TRY(dev_alloc(pl, &pl->lin, Bm * 3 * g.N, false));
    TRY(dev_alloc(pl, &pl->rhs, Bm * 3 * g.N, false));
    TRY(dev_alloc(pl, &pl->xtmp, Bm * 3 * g.N, false));
    pl->nke = pl->Mh3p / (8 * pl->synth_nt_ke);
    TRY(dev_alloc(pl, &pl->kepart, Bm * std::max(pl->nke, n), true));
    TRY(dev_alloc(pl, &pl->zeroRa, Bm, true));
    TRY(dev_alloc(pl, &pl->dpart, 2 * Bm * 18, true));   // two records in flight (sddc_time_step*)
    {
        std::vector<double> w(K);
        for (int k = 0; k < K; ++k) w[k] = k == 1 ? 0.0 : 1.0 / (1.0 - (double)k * (double)k);
        TRY(upload(pl, &pl->nu_w, w));
    }
    // ---- opt in to large dynamic shared memory ----
    {
        SynthParams sp{};
        TRY(launch_synth<EPI_FX>(pl, sp, 0, pl->synth_smem_fx, 1, 1, nullptr, true));
        if (pl->dfx_ok) TRY(launch_synth<EPI_DFX>(pl, sp, 0, pl->synth_smem_dfx, 1, 1, nullptr, true));
        TRY(launch_synth<EPI_KE>(pl, sp, 0, pl->synth_smem_ke, 1, 1, nullptr, true));
        pl->ana_stage = ANA_MAX_STAGES;
        pl->ana_smem = ANA_MAX_STAGES * ana_stage_doubles(g.nt8) * sizeof(double);
        AnaParams ap{};
        TRY(launch_analysis(pl, ap, 1, nullptr, true));
    }
    {
        cudaDeviceProp prop;
        TRYC(cudaGetDeviceProperties(&prop, pl->device));
        pl->num_sms = prop.multiProcessorCount;
        pl->ws_smem = synth_wsq_smem_doubles(n8) * sizeof(double);
        // persistent quarter-wave synthesis: 24 < n <= 32 and a grid the second mirror level divides
        pl->ws_ok = g.nt8 == 4 && pl->synth_nt_dfx == SWS_NT && pl->dfx_ok && pl->ws_smem <= SMEM_LIMIT &&
                    g.Kh % 128 == 0 && g.Mh % 16 == 0 && g.Mhp == g.Mh;
        pl->quarter = pl->ws_ok;
        if (pl->quarter) {
            const int nch = g.Khp / 8, KT3q = 32 * pl->ana_nt;
            TRY(dev_alloc(pl, &pl->tab1q, 4ull * g.Mhp * g.Khp, false));
            TRY(dev_alloc(pl, &pl->tab2q, 4ull * g.Khp2 * g.Mhp, false));
            fill_table_quarter_kernel<<<296, 256>>>(pl->tab1q, 2, g.M, g.Kh, g.Mh, 16, (g.Mh / 2) / 8, nch);
            fill_table_quarter_kernel<<<296, 256>>>(pl->tab2q, 3, g.M, g.Kh, g.Mh, KT3q, g.Khp2 / KT3q, g.Mhp / 8);
            pl->launches += 2;
            TRYC(cudaGetLastError());
            TRY(set_smem(pl, (synth_wsq_kernel<4, SWS_FX>), pl->ws_smem));
            TRY(set_smem(pl, (synth_wsq_kernel<4, SWS_GRID>), pl->ws_smem));
            TRY(set_smem(pl, (synth_wsq_kernel<4, SWS_JVPC>), pl->ws_smem));
        }
    }
    {
        // FFT formulation of the nonlinear term (k_nlin_fft.cuh); SDDC_FLAG_DENSE_TRANSFORMS keeps the dense DMMA
        // transforms (validation of the path every other N_fm takes, at the headline shape)
        const bool want = !(cfg->flags & SDDC_FLAG_DENSE_TRANSFORMS);
        if (want && (K == 128 || K == 256 || K == 512)) {
            pl->fft_M = g.M;
            pl->fft_dfx = true;
            std::vector<double> tab;
            switch (g.M) {
                case 192: tab.resize(fftp::tab_doubles<192>()); fftp::fill_tables<192>(tab.data()); break;
                case 384: tab.resize(fftp::tab_doubles<384>()); fftp::fill_tables<384>(tab.data()); break;
                default: tab.resize(fftp::tab_doubles<768>()); fftp::fill_tables<768>(tab.data()); break;
            }
            TRY(upload(pl, &pl->fft_tab, tab));
            TRY(dev_alloc(pl, &pl->coef7, Bm * 7 * g.N, false));
            if (pl->fft_dfx) TRY(dev_alloc(pl, &pl->coef7b, Bm * 7 * g.N, false));
            TRY(dev_alloc(pl, &pl->spec4, Bm * g.n * (size_t)spec_pitch(K), false));
            {
                double* cnt = nullptr;
                TRY(dev_alloc(pl, &cnt, 2, true));
                pl->fft_row = reinterpret_cast<int*>(cnt);
            }
            TRY(run_nlin_fft(pl, nullptr, nullptr, nullptr, false, 1, nullptr, true));
            TRY(set_smem(pl, post_kernel, post_smem_bytes(n, n8)));
            {
                // kinetic energy on the 3K grid with the same transform code (M = 384, 768 or 1536)
                pl->ke_M = 3 * K;
                std::vector<double> kt, kw(pl->ke_M);
                if (pl->ke_M == 384) {
                    kt.resize(fftp::tab_doubles<384>()); fftp::fill_tables<384>(kt.data()); fftp::fill_ke_weights<384>(kw.data());
                    TRY(set_smem(pl, (ke_fft_kernel<384, 8>), ke_fft_smem_bytes<384>(8)));
                } else if (pl->ke_M == 1536) {
                    kt.resize(fftp::tab_doubles<1536>()); fftp::fill_tables<1536>(kt.data()); fftp::fill_ke_weights<1536>(kw.data());
                    static_assert(ke_fft_smem_bytes<1536>(KE_NW1536) <= SMEM_LIMIT, "KE workers do not fit into shared memory");
                    TRY(set_smem(pl, (ke_fft_kernel<1536, KE_NW1536>), ke_fft_smem_bytes<1536>(KE_NW1536)));
                } else {
                    kt.resize(fftp::tab_doubles<768>()); fftp::fill_tables<768>(kt.data()); fftp::fill_ke_weights<768>(kw.data());
                    TRY(set_smem(pl, (ke_fft_kernel<768, KE_NW768>), ke_fft_smem_bytes<768>(KE_NW768)));
                }
                TRY(upload(pl, &pl->ke_tab, kt));
                TRY(upload(pl, &pl->ke_Wn, kw));
                if (pl->ke_M == 768) {
                    // side-stream variant: two workers per CTA (65 KB), so that it fits next to the back-substitution's CTAs
                    TRY(set_smem(pl, (ke_fft_kernel<768, 2>), ke_fft_smem_bytes<768>(2)));
                    int lo = 0, hi = 0;
                    TRYC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    TRYC(cudaStreamCreateWithPriority(&pl->ke_stream, cudaStreamNonBlocking, lo));
                    TRYC(cudaEventCreateWithFlags(&pl->ev_rows, cudaEventDisableTiming));
                    TRYC(cudaEventCreateWithFlags(&pl->ev_ke, cudaEventDisableTiming));
                }
            }
            TRY(set_smem(pl, (prep_kernel<3, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<4, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<5, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<6, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<7, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<8, true>), prep_smem_bytes(n8, 1)));
        }
    }
    {
        // shapes neither the FFT kernels nor the mirror-split DMMA transforms cover go through the same row pipeline with
        // the direct-summation row kernel (k_nlin_fft.cuh): every product for N_fm = 2 (mod 4), the two-state products only
        // where the dense two-state synthesis does not fit into shared memory (N_r > 41)
        const bool all_direct = K % 4 != 0;
        const bool jvp_direct = !all_direct && pl->fft_M == 0 && !pl->dfx_ok;
        if (all_direct || jvp_direct) {
            if ((size_t)(14 * K + 4 * g.M) * sizeof(double) > 48 * 1024) {
                pl->err = "direct-summation path: N_fm too large";
                return fail(SDDC_ERR_UNSUPPORTED);
            }
            pl->fft_direct = all_direct; pl->dfx_direct = jvp_direct;
            if (all_direct) pl->fft_M = g.M;
            pl->fft_dfx = true;
            TRY(dev_alloc(pl, &pl->coef7, Bm * 7 * g.N, false));
            TRY(dev_alloc(pl, &pl->coef7b, Bm * 7 * g.N, false));
            TRY(dev_alloc(pl, &pl->spec4, Bm * g.n * (size_t)spec_pitch(K), false));
            TRY(set_smem(pl, post_kernel, post_smem_bytes(n, n8)));
            TRY(set_smem(pl, (prep_kernel<3, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<4, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<5, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<6, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<7, true>), prep_smem_bytes(n8, 1)));
            TRY(set_smem(pl, (prep_kernel<8, true>), prep_smem_bytes(n8, 1)));
        }
    }
    pl->solve_nsl = SOLVE_NSL;
    while (pl->solve_nsl > 2 && solve_smem_doubles<SOLVE_NTB>(n8, pl->solve_nsl) * sizeof(double) > SMEM_LIMIT) --pl->solve_nsl;
    pl->solve_smem = solve_smem_doubles<SOLVE_NTB>(n8, pl->solve_nsl) * sizeof(double);
    pl->solve_hot_nsl = solve_hot_smem_bytes(n8, 3) <= SMEM_LIMIT ? 3 : 2;
    pl->solve_hot_smem = solve_hot_smem_bytes(n8, pl->solve_hot_nsl);
    // gather mode (k_solve_hot.cuh): FFT formulation, sector-aligned parity halves, at most four radial row tiles
    // The gather mode is NOT part of the shipped configuration (compile with -DSDDC_EXPERIMENTAL_GATHER to get it): it is
    // 5 % faster per step, but its results were not bit-reproducible from run to run -- whole T / S member tiles off by
    // 1e-8 in a few runs out of forty at 512 members, in every run of one experimental schedule on equatorially symmetric
    // plans (tools/determinism_probe.py) -- and the race was not found.  The four-kernel path is
    // bit-reproducible in every probe.
#ifdef SDDC_EXPERIMENTAL_GATHER
    pl->solve_gath = pl->fft_M != 0 && !pl->fft_direct && n8 <= 32 && K % 8 == 0 && pl->solve_hot_nsl == 3;
#else
    pl->solve_gath = false;
#endif
    if (pl->solve_gath) {
        pl->solve_gath_smem = solve_gath_smem_bytes(n8);
        if (n8 == 24) {
            TRY(set_smem(pl, (solve_hot_kernel<3, 3, false, false, true>), pl->solve_gath_smem));
            TRY(set_smem(pl, (solve_hot_kernel<3, 3, true, false, true>), pl->solve_gath_smem));
            TRY(set_smem(pl, (solve_hot_kernel<3, 3, false, true, true>), pl->solve_gath_smem));
        } else {
            TRY(set_smem(pl, (solve_hot_kernel<4, 3, false, false, true>), pl->solve_gath_smem));
            TRY(set_smem(pl, (solve_hot_kernel<4, 3, true, false, true>), pl->solve_gath_smem));
            TRY(set_smem(pl, (solve_hot_kernel<4, 3, false, true, true>), pl->solve_gath_smem));
        }
    }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<3, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<3, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<3, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<3, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<3, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<3, 2, false, true>), pl->solve_hot_smem)); }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<4, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<4, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<4, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<4, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<4, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<4, 2, false, true>), pl->solve_hot_smem)); }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<5, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<5, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<5, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<5, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<5, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<5, 2, false, true>), pl->solve_hot_smem)); }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<6, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<6, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<6, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<6, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<6, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<6, 2, false, true>), pl->solve_hot_smem)); }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<7, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<7, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<7, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<7, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<7, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<7, 2, false, true>), pl->solve_hot_smem)); }
    if (pl->solve_hot_nsl == 3) { TRY(set_smem(pl, (solve_hot_kernel<8, 3, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<8, 3, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<8, 3, false, true>), pl->solve_hot_smem)); }
    else { TRY(set_smem(pl, (solve_hot_kernel<8, 2, false>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<8, 2, true>), pl->solve_hot_smem)); TRY(set_smem(pl, (solve_hot_kernel<8, 2, false, true>), pl->solve_hot_smem)); }
    TRY(set_smem(pl, solve_kernel<SOLVE_NTB, false>, pl->solve_smem));
    TRY(set_smem(pl, prep_kernel<3>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, prep_kernel<4>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, prep_kernel<5>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, prep_kernel<6>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, prep_kernel<7>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, prep_kernel<8>, prep_smem_bytes(n8, 1)));
    TRY(set_smem(pl, linop_kernel, sizeof(double) * ((size_t)PREP_TC * n + (size_t)n * n8)));
    TRY(set_smem(pl, ke_prep_kernel, sizeof(double) * ((size_t)32 * n + (size_t)n * n8)));
    TRYC(cudaDeviceSynchronize());
#undef TRY
#undef TRYC
    *out = pl;
    return SDDC_OK;
}

int sddc_plan_set_linv(sddc_plan* pl, int which, const double* Linv, double dt_eff) {
    if (!pl || !Linv || which < 0 || which > 2) return SDDC_ERR_INVALID;
    PLAN_CUDA(pl, cudaSetDevice(pl->device));
    const Geo& g = pl->g;
    std::vector<double> h;
    if (which == 0) {
        pl->h_LinvA4.assign(Linv, Linv + (size_t)g.K * g.n * g.n);
        h = build_a4_stack(Linv, pl->h_D2.data(), g.K, g.n, g.n8, pl->LDL);
    } else {
        h = pad_stack(Linv, g.K, g.n, g.n8, pl->LDL);
    }
    double* dst = which == 0 ? pl->LA4 : (which == 1 ? pl->LT : pl->LS);
    PLAN_CUDA(pl, cudaDeviceSynchronize());
    PLAN_CUDA(pl, cudaMemcpy(dst, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    (which == 0 ? pl->dt_psi : (which == 1 ? pl->dt_T : pl->dt_S)) = dt_eff;
    return SDDC_OK;
}

int sddc_plan_set_a4_aux(sddc_plan* pl, const double* D2, const double* ir2, const double* ir4) {
    if (!pl || !D2 || !ir2 || !ir4) return SDDC_ERR_INVALID;
    PLAN_CUDA(pl, cudaSetDevice(pl->device));
    const Geo& g = pl->g;
    pl->h_D2.assign(D2, D2 + (size_t)g.n * g.n);
    std::vector<double> h = build_a4_stack(pl->h_LinvA4.data(), D2, g.K, g.n, g.n8, pl->LDL);
    PLAN_CUDA(pl, cudaDeviceSynchronize());
    PLAN_CUDA(pl, cudaMemcpy(pl->LA4, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    PLAN_CUDA(pl, cudaMemcpy(pl->a4_ir2, ir2, g.n * sizeof(double), cudaMemcpyHostToDevice));
    PLAN_CUDA(pl, cudaMemcpy(pl->a4_ir4, ir4, g.n * sizeof(double), cudaMemcpyHostToDevice));
    return SDDC_OK;
}

int sddc_nlin_fx(sddc_plan* pl, const double* X, double* F, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pl->fft_M) {
        if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st, pl->coef7))) return rc;
        return run_nlin_fft(pl, pl->coef7, nullptr, F, false, B, st);
    }
    if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st))) return rc;
    if ((rc = run_synth_nl(pl, false, B, st))) return rc;
    return run_analysis(pl, F, false, B, st, pl->quarter);
}

int sddc_nlin_dfx(sddc_plan* pl, const double* dv, const double* X, double* F, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pl->fft_dfx) {
        if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st, pl->coef7))) return rc;
        if ((rc = run_prep(pl, dv, 1, true, nullptr, nullptr, nullptr, B, st, pl->coef7b))) return rc;
        return run_nlin_fft(pl, pl->coef7, pl->coef7b, F, false, B, st);
    }
    if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st))) return rc;
    if ((rc = run_prep(pl, dv, 1, true, nullptr, nullptr, nullptr, B, st))) return rc;
    if ((rc = run_synth_nl(pl, true, B, st))) return rc;
    return run_analysis(pl, F, false, B, st, false);
}

int sddc_linear_op(sddc_plan* pl, int op, const double* in, double* out, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (op < 0 || op > 5) { pl->err = "unknown linear operator code"; return SDDC_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LinopParams lp{};
    lp.in = in; lp.out = out; lp.g = pl->g; lp.B = B; lp.op = op;
    if (op <= 3) {
        if ((rc = run_scan(pl, in, pl->g.N, B, st))) return rc;
        lp.JJ = pl->JJ;
    }
    switch (op) {
        case SDDC_OP_DT0_THETA: lp.vec = pl->dT0; break;
        case SDDC_OP_A2_SINE: lp.vec = pl->ir2; lp.matT = pl->DsqT; break;
        case SDDC_OP_A2_SINE_R2: lp.vec = pl->ir4; lp.matT = pl->D2rT; break;
        case SDDC_OP_KGR: lp.vec = pl->gb; break;
        case SDDC_OP_R2: lp.vec = pl->r2; break;
        default: break;
    }
    dim3 grid((pl->g.K + PREP_TC - 1) / PREP_TC, B);
    const size_t smem = sizeof(double) * ((size_t)PREP_TC * pl->g.n + (size_t)pl->g.n * pl->g.n8);
    linop_kernel<<<grid, 256, smem, st>>>(lp);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

int sddc_solve_a4(sddc_plan* pl, const double* g, double* f, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    return run_solve(pl, g, nullptr, pl->g.N, 0, f, pl->g.N, 0, nullptr, 0, 1, B, static_cast<cudaStream_t>(stream));
}

int sddc_solve_nab2(sddc_plan* pl, int which, const double* g, double* f, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (which != 0 && which != 1) { pl->err = "which must be 0 (T) or 1 (S)"; return SDDC_ERR_INVALID; }
    return run_solve(pl, g, nullptr, pl->g.N, 0, f, pl->g.N, 0, nullptr, 1 + which, 1, B, static_cast<cudaStream_t>(stream));
}

int sddc_step(sddc_plan* pl, const double* Xin, double* Xout, const double* Ra, const double* Ras, int B, int nsteps,
              int linear, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (nsteps < 1 || Xin == Xout) { pl->err = "nsteps must be >= 1 and Xout must not alias Xin"; return SDDC_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double* src = Xin;
    for (int s = 0; s < nsteps; ++s) {
        double* dst = ((nsteps - s) & 1) ? Xout : pl->xtmp;  // the last step lands in Xout
        if ((rc = run_member_step(pl, src, dst, nullptr, Ra, Ras, B, linear != 0, st, s > 0, s + 1 < nsteps))) return rc;
        src = dst;
    }
    return SDDC_OK;
}

int sddc_time_step(sddc_plan* pl, const double* Xin, double* Xout, const double* Ra, const double* Ras, int B, int nsteps,
                   int linear, int diag_every, double* diag_hist, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (nsteps < 1 || Xin == Xout || diag_every < 0 || (diag_every > 0 && !diag_hist)) {
        pl->err = "sddc_time_step: nsteps >= 1, Xout must not alias Xin, diag_hist required when diag_every > 0";
        return SDDC_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool share = pl->fft_M != 0 && pl->ke_M != 0 && !linear;   // see sddc_time_step_host
    const double* src = Xin;
    int pending = -1;
    // The kinetic-energy transform of record r only needs the spectral rows of the prep stage that follows it; on a
    // low-priority side stream it runs under the latency-bound back-substitution of the same step instead of in front
    // of it (two-worker CTAs fit next to the back-substitution's).  The compute stream waits for it before the next
    // prep stage overwrites the rows.
    const bool side = share && pl->ke_stream != nullptr;
    bool ke_inflight = false;
    const size_t dstride = (size_t)pl->cfg.max_batch * 18;
    for (int s = 1; s <= nsteps; ++s) {
        double* dst = ((nsteps - s + 1) & 1) ? Xout : pl->xtmp;  // the last step lands in Xout
        if (ke_inflight) { PLAN_CUDA(pl, cudaStreamWaitEvent(st, pl->ev_ke, 0)); ke_inflight = false; }
        if ((rc = run_step_prep(pl, src, Ra, Ras, B, linear != 0, st, s > 1))) return rc;
        double* rec = pending >= 0 ? diag_hist + (size_t)pending * B * 6 : nullptr;
        const double* dp = pl->dpart + ((s - 1) & 1) * dstride;     // written by the back-substitution of step s - 1
        if (rec && side) PLAN_CUDA(pl, cudaEventRecord(pl->ev_rows, st));
        if (rec && !side && (rc = run_ke_fft(pl, src, pl->coef7, 7LL * pl->g.K, 3 * pl->g.K, pl->ir, rec, B, st, dp))) return rc;
        // the back-substitution also leaves ||X_s||^2 and the Nusselt sums when X_s gets a shared-prep record
        const bool rec_shared = share && diag_every && s % diag_every == 0 && s < nsteps;
        if ((rc = run_step_rest(pl, dst, nullptr, B, linear != 0, st, s < nsteps, rec_shared ? pl->dpart + (s & 1) * dstride : nullptr))) return rc;
        if (rec && side) {
            // submitted after the step's own kernels: with equal stream priorities the transform then takes what the
            // row kernel, the finishing stage and the back-substitution leave free
            PLAN_CUDA(pl, cudaStreamWaitEvent(pl->ke_stream, pl->ev_rows, 0));
            if ((rc = run_ke_fft(pl, src, pl->coef7, 7LL * pl->g.K, 3 * pl->g.K, pl->ir, rec, B, pl->ke_stream, dp, true))) return rc;
            PLAN_CUDA(pl, cudaEventRecord(pl->ev_ke, pl->ke_stream));
            ke_inflight = true;
        }
        pending = -1;
        src = dst;
        if (diag_every && s % diag_every == 0) {
            const int r = s / diag_every - 1;
            if (share && s < nsteps) pending = r;
            else {
                if (ke_inflight) { PLAN_CUDA(pl, cudaStreamWaitEvent(st, pl->ev_ke, 0)); ke_inflight = false; }   // kepart is shared
                if ((rc = sddc_diagnostics(pl, src, diag_hist + (size_t)r * B * 6, B, stream))) return rc;
            }
        }
    }
    if (ke_inflight) PLAN_CUDA(pl, cudaStreamWaitEvent(st, pl->ev_ke, 0));
    return SDDC_OK;
}

int sddc_residual(sddc_plan* pl, const double* X, double* out, const double* Ra, const double* Ras, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (X == out) { pl->err = "out must not alias X"; return SDDC_ERR_INVALID; }
    return run_member_step(pl, X, out, X, Ra, Ras, B, false, static_cast<cudaStream_t>(stream));
}

int sddc_jvp(sddc_plan* pl, const double* dv, const double* X, double* out, const double* Ra, const double* Ras, int B,
             void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (dv == out || X == out) { pl->err = "out must not alias dv or X"; return SDDC_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N3 = 3LL * pl->g.N;
    if (pl->fft_dfx) {
        if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st, pl->coef7))) return rc;
        if ((rc = run_prep(pl, dv, 1, true, pl->lin_sm, Ra, Ras, B, st, pl->coef7b))) return rc;
        return run_rows_and_solve(pl, pl->coef7, pl->coef7b, 0, out, dv, B, st);
    }
    if ((rc = run_prep(pl, X, 0, true, nullptr, nullptr, nullptr, B, st))) return rc;
    if ((rc = run_prep(pl, dv, 1, true, pl->lin_sm, Ra, Ras, B, st))) return rc;
    if ((rc = run_synth_nl(pl, true, B, st))) return rc;
    if ((rc = run_analysis(pl, pl->f_sm, true, B, st, false))) return rc;
    return run_solve(pl, pl->lin_sm, pl->f_sm, -1, 0, out, N3, pl->g.N, dv, 0, 3, B, st);
}

// launch the persistent warp-specialised synthesis in one of its cached-base modes
static int run_synth_ws_mode(sddc_plan* pl, int mode, const double* coef, int B, cudaStream_t st) {
    SynthParams sp{};
    sp.coef0 = coef; sp.coef_stride = pl->coef_member_stride; sp.tab = pl->tab1d; sp.Dr = pl->Dr; sp.prd = pl->prd;
    sp.gridc = pl->gridc; sp.g = pl->g;
    const int ntj = pl->g.Mhp / SWS_W, nwork = ntj * B;
    const int grid = std::min(nwork, pl->num_sms);
    StageTimer tm(pl, SDDC_STAGE_SYNTH, st);
    sp.tab = pl->tab1q;
    if (mode == SWS_GRID) synth_wsq_kernel<4, SWS_GRID><<<grid, SWS_NTHR, pl->ws_smem, st>>>(sp, ntj, nwork);
    else synth_wsq_kernel<4, SWS_JVPC><<<grid, SWS_NTHR, pl->ws_smem, st>>>(sp, ntj, nwork);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

int sddc_jvp_set_base(sddc_plan* pl, const double* X, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geo& g = pl->g;
    if (!pl->xbase) {
        if ((rc = dev_alloc(pl, &pl->xbase, (size_t)pl->cfg.max_batch * 3 * g.N, false))) return rc;
        if (pl->fft_dfx && pl->fft_M != 0 && !pl->fft_direct) {
            if ((rc = dev_alloc(pl, &pl->grid7, (size_t)pl->cfg.max_batch * g.n * 7 * g.M, false))) return rc;
        } else if (pl->fft_dfx) {
            if ((rc = dev_alloc(pl, &pl->coef7base, (size_t)pl->cfg.max_batch * 7 * g.N, false))) return rc;
        } else if (pl->ws_ok && (rc = dev_alloc(pl, &pl->gridc, (size_t)pl->cfg.max_batch * 9 * 2 * g.n8 * g.Mhp, true))) return rc;
    }
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->xbase, X, sizeof(double) * (size_t)B * 3 * g.N, cudaMemcpyDeviceToDevice, st));
    pl->base_B = B;
    // FFT kernels: the seven grid fields of the base state are synthesised once and cached in the order the product phase
    // reads them (fft_fused.h, i3f1_grid / i3f1_jvpc): every product then needs the four inverse transforms of the
    // perturbation only, not the seven of a pair of states
    if (pl->grid7) {
        if ((rc = run_prep(pl, pl->xbase, 0, true, nullptr, nullptr, nullptr, B, st, pl->coef7))) return rc;
        return run_nlin_fft(pl, pl->coef7, nullptr, nullptr, false, B, st, false, 1);
    }
    // direct-summation rows: the base state is kept as its seven spectral rows
    if (pl->fft_dfx) return run_prep(pl, pl->xbase, 0, true, nullptr, nullptr, nullptr, B, st, pl->coef7base);
    if (pl->ws_ok) {
        if ((rc = run_prep(pl, pl->xbase, 0, true, nullptr, nullptr, nullptr, B, st))) return rc;
        return run_synth_ws_mode(pl, SWS_GRID, pl->coef, B, st);
    }
    return SDDC_OK;
}

static int jvp_apply_impl(sddc_plan* pl, const double* dv, double* out, const double* Ra, const double* Ras, int B, void* stream,
                          bool plus_identity) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (!pl->xbase || B != pl->base_B) { pl->err = "sddc_jvp_apply: call sddc_jvp_set_base with the same batch first"; return SDDC_ERR_INVALID; }
    if (dv == out) { pl->err = "out must not alias dv"; return SDDC_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N3 = 3LL * pl->g.N;
    const double* sub = plus_identity ? nullptr : dv;   // PDFX subtracts dv at the end (Main.py:511-519)
    if (pl->grid7) {
        if ((rc = run_prep(pl, dv, 1, true, pl->lin_sm, Ra, Ras, B, st, pl->coef7b))) return rc;
        return run_rows_and_solve(pl, nullptr, pl->coef7b, 2, out, sub, B, st);
    }
    if (pl->fft_dfx) {
        if ((rc = run_prep(pl, dv, 1, true, pl->lin_sm, Ra, Ras, B, st, pl->coef7b))) return rc;
        return run_rows_and_solve(pl, pl->coef7base, pl->coef7b, 0, out, sub, B, st);
    }
    if (!pl->ws_ok) {
        if (plus_identity) { pl->err = "sddc_jvp_apply_plus is not available on this path"; return SDDC_ERR_UNSUPPORTED; }
        return sddc_jvp(pl, dv, pl->xbase, out, Ra, Ras, B, stream);
    }
    if ((rc = run_prep(pl, dv, 1, true, pl->lin_sm, Ra, Ras, B, st))) return rc;
    if ((rc = run_synth_ws_mode(pl, SWS_JVPC, pl->coef1, B, st))) return rc;
    if ((rc = run_analysis(pl, pl->f_sm, true, B, st, pl->quarter))) return rc;
    return run_solve(pl, pl->lin_sm, pl->f_sm, -1, 0, out, N3, pl->g.N, sub, 0, 3, B, st);
}

int sddc_jvp_apply(sddc_plan* pl, const double* dv, double* out, const double* Ra, const double* Ras, int B, void* stream) {
    return jvp_apply_impl(pl, dv, out, Ra, Ras, B, stream, false);
}

int sddc_jvp_apply_plus(sddc_plan* pl, const double* dv, double* out, const double* Ra, const double* Ras, int B, void* stream) {
    return jvp_apply_impl(pl, dv, out, Ra, Ras, B, stream, true);
}

int sddc_dF_dRa(sddc_plan* pl, const double* X, double* out, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N = pl->g.N, N3 = 3 * N;
    // f_mu = dt Pr G(T) (Main.py:835): kGR operator on the T field, scaled
    LinopParams lp{};
    lp.in = X + N; lp.out = pl->rhs; lp.g = pl->g; lp.B = B; lp.op = SDDC_OP_KGR; lp.vec = pl->gb;
    // the T field of member b sits at X + b*3N + N: run member by member through strides = use a strided variant
    // (linop_kernel assumes stride N), so gather through cudaMemcpy2DAsync first.
    PLAN_CUDA(pl, cudaMemcpy2DAsync(pl->lin, N * sizeof(double), X + N, N3 * sizeof(double), N * sizeof(double), B,
                                    cudaMemcpyDeviceToDevice, st));
    lp.in = pl->lin;
    dim3 grid((pl->g.K + PREP_TC - 1) / PREP_TC, B);
    const size_t smem = sizeof(double) * ((size_t)PREP_TC * pl->g.n + (size_t)pl->g.n * pl->g.n8);
    linop_kernel<<<grid, 256, smem, st>>>(lp);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    const double s = pl->g.dt * pl->g.Pr;
    axpby_kernel<<<296, 256, 0, st>>>(pl->rhs, pl->rhs, pl->rhs, s, 0.0, (long long)B * N);
    pl->launches++;
    PLAN_CUDA(pl, cudaMemsetAsync(out, 0, sizeof(double) * B * N3, st));
    return run_solve(pl, pl->rhs, nullptr, N, 0, out, N3, N, nullptr, 0, 1, B, st);
}

int sddc_diagnostics(sddc_plan* pl, const double* X, double* out, int B, void* stream) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geo& g = pl->g;
    if ((rc = run_scan(pl, X, 3LL * g.N, B, st))) return rc;
    KEPrepParams kp{};
    kp.X = X; kp.x_stride = 3LL * g.N; kp.JJ = pl->JJ; kp.coef = pl->coef1; kp.coef_stride = pl->coef_member_stride;
    kp.DrT = pl->DrT; kp.ir = pl->ir; kp.g = g;
    dim3 grid((g.K + 31) / 32, B);
    if (pl->ke_M) {
        // FFT formulation: two row-major coefficient rows per radial point (scratch: coef7), one transform per row
        kp.coef = pl->coef7; kp.rows = 1;
        {
            StageTimer tm(pl, SDDC_STAGE_KE_PREP, st);
            ke_prep_kernel<<<grid, 256, sizeof(double) * ((size_t)32 * g.n + (size_t)g.n * g.n8), st>>>(kp);
        }
        pl->launches++;
        PLAN_CUDA(pl, cudaGetLastError());
        if ((rc = run_ke_fft(pl, X, pl->coef7, 2LL * g.K, g.K, nullptr, out, B, st))) return rc;
        return SDDC_OK;
    }
    {
        StageTimer tm(pl, SDDC_STAGE_KE_PREP, st);
        ke_prep_kernel<<<grid, 256, sizeof(double) * ((size_t)32 * g.n + (size_t)g.n * g.n8), st>>>(kp);
    }
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    SynthParams sp{};
    sp.coef0 = pl->coef1; sp.coef_stride = pl->coef_member_stride; sp.tab = pl->tab3;
    sp.wr = pl->wr; sp.wth = pl->wth; sp.kepart = pl->kepart; sp.g = g;
    {
        StageTimer tm(pl, SDDC_STAGE_KE_SYNTH, st);
        if ((rc = launch_synth<EPI_KE>(pl, sp, pl->synth_stage_ke, pl->synth_smem_ke, pl->Mh3p / (8 * pl->synth_nt_ke), B, st))) return rc;
    }
    StageTimer tm2(pl, SDDC_STAGE_DIAG, st);
    diag_kernel<<<B, 256, 0, st>>>(X, pl->kepart, pl->nke, pl->nu_in, pl->nu_out, pl->ke_scale, g, out);
    pl->launches++;
    PLAN_CUDA(pl, cudaGetLastError());
    return SDDC_OK;
}

int sddc_transform(int kind, const double* in, double* out, int rows, int n_in, int n_out, void* stream) {
    if (kind < 0 || kind > 3 || rows < 1 || n_in < 1 || n_out < 1) return SDDC_ERR_INVALID;
    if (kind >= 2 && n_out > n_in) return SDDC_ERR_INVALID;
    const long long tot = (long long)rows * n_out;
    transform_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(kind, in, out, rows, n_in, n_out);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

int sddc_interp_radial(const double* in, double* out, const double* W, long long rows, int nr_o, int nr_n, void* stream) {
    if (!in || !out || !W || rows < 1 || nr_o < 1 || nr_n < 1 || in == out) return SDDC_ERR_INVALID;
    const size_t smem = sizeof(double) * (size_t)nr_o * nr_n;
    if (smem > 48 * 1024) return SDDC_ERR_UNSUPPORTED;
    const long long tot = rows * nr_n;
    const int grid = (int)std::min<long long>((tot + 255) / 256, 148 * 16);
    interp_radial_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(in, out, W, rows, nr_o, nr_n);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

int sddc_interp_thetas(const double* in, double* out, int B, int K_o, int K_n, int nr, void* stream) {
    if (!in || !out || B < 1 || K_o < 1 || K_n < 1 || nr < 1 || in == out) return SDDC_ERR_INVALID;
    const long long tot = 3LL * K_n * nr * B;
    const int grid = (int)std::min<long long>((tot + 255) / 256, 148 * 16);
    interp_thetas_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, B, K_o, K_n, nr);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

// ---- batched Arnoldi orthogonalisation (k_krylov.cuh); stateless like sddc_transform ----
static int gs_fill(GsParams& gp, const double* V, long long member_stride, int n, int nvec, double* w, int ldp, int B) {
    if (!V || !w || n < 1 || nvec < 1 || B < 1 || ldp < nvec + 1 || member_stride < (long long)nvec * n) return SDDC_ERR_INVALID;
    gp.V = V; gp.member_stride = member_stride; gp.n = n; gp.nvec = nvec; gp.w = w;
    gp.nchunk = (n + GS_CHUNK - 1) / GS_CHUNK; gp.ldp = ldp;
    return SDDC_OK;
}

int sddc_gs_chunks(int n) { return n < 1 ? 0 : (n + GS_CHUNK - 1) / GS_CHUNK; }

int sddc_gs_dots(const double* V, long long member_stride, int n, int nvec, const double* w, double* part, int ldp, int B,
                 void* stream) {
    GsParams gp{};
    int rc = gs_fill(gp, V, member_stride, n, nvec, const_cast<double*>(w), ldp, B);
    if (rc || !part) return SDDC_ERR_INVALID;
    gp.part_out = part;
    gs_dots_kernel<<<dim3(gp.nchunk, B), GS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(gp);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

int sddc_gs_update(const double* V, long long member_stride, int n, int nvec, double* w, const double* part_in,
                   double* h_out, double* part_out, int ldp, int want_dots, const int* member_mask, int B, void* stream) {
    GsParams gp{};
    int rc = gs_fill(gp, V, member_stride, n, nvec, w, ldp, B);
    if (rc || !part_in || !h_out || !part_out || part_in == part_out) return SDDC_ERR_INVALID;
    gp.part_in = part_in; gp.part_out = part_out; gp.h_out = h_out; gp.member_mask = member_mask;
    const size_t smem = sizeof(double) * ((size_t)GS_CHUNK + nvec);
    if (smem > 48 * 1024) return SDDC_ERR_UNSUPPORTED;   // nvec <= 5120: far beyond any Krylov space in use
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (want_dots) gs_update_kernel<true><<<dim3(gp.nchunk, B), GS_THREADS, smem, st>>>(gp);
    else gs_update_kernel<false><<<dim3(gp.nchunk, B), GS_THREADS, smem, st>>>(gp);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

int sddc_gmres_column(const double* h, int ldh, const double* hn, double* H, double* cs, double* sn, double* g, double* resid,
                      const double* tol, int* live, int* any_live, int B, int j, int m, int shifted, void* stream) {
    if (!h || !hn || !H || !cs || !sn || !g || !resid || !tol || !live || !any_live || B < 1 || j < 0 || j >= m || ldh < j + 1)
        return SDDC_ERR_INVALID;
    GmresColParams gp{};
    gp.h = h; gp.hn = hn; gp.H = H; gp.cs = cs; gp.sn = sn; gp.g = g; gp.resid = resid; gp.tol = tol; gp.live = live;
    gp.any_live = any_live; gp.B = B; gp.j = j; gp.m = m; gp.ldh = ldh; gp.shifted = shifted;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(any_live, 0, sizeof(int), st) != cudaSuccess) return SDDC_ERR_CUDA;
    gmres_column_kernel<<<(B + 127) / 128, 128, 0, st>>>(gp);
    return cudaGetLastError() == cudaSuccess ? SDDC_OK : SDDC_ERR_CUDA;
}

int sddc_profile_begin(sddc_plan* pl) {
    if (!pl) return SDDC_ERR_INVALID;
    pl->profiling = true;
    return SDDC_OK;
}

int sddc_profile_end(sddc_plan* pl, double* ms, int* counts) {
    if (!pl || !ms || !counts) return SDDC_ERR_INVALID;
    pl->profiling = false;
    for (int i = 0; i < SDDC_STAGE_COUNT; ++i) { ms[i] = 0.0; counts[i] = 0; }
    PLAN_CUDA(pl, cudaSetDevice(pl->device));
    PLAN_CUDA(pl, cudaDeviceSynchronize());
    for (auto& e : pl->events) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) { ms[e.stage] += t; counts[e.stage]++; }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    pl->events.clear();
    return SDDC_OK;
}

// Host-buffer member-steps.  The batch is cut into chunks of members that flow through a three-stream pipeline
// (H2D copy | kernels | D2H copy), so PCIe transfers in both directions overlap each other and the compute.
int sddc_step_host(sddc_plan* pl, const double* Xin, double* Xout, const double* Ra, const double* Ras, int B,
                   int nsteps, int linear, double* diag_out) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if ((rc = ensure_host_staging(pl))) return rc;
    const size_t W = 3 * (size_t)pl->g.N;
    const int chunk = std::min(B, std::max(16, (B + 7) / 8));
    const int nchunks = (B + chunk - 1) / chunk;
    while ((int)pl->ev_in.size() < nchunks) {
        cudaEvent_t a, b2;
        PLAN_CUDA(pl, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        PLAN_CUDA(pl, cudaEventCreateWithFlags(&b2, cudaEventDisableTiming));
        pl->ev_in.push_back(a); pl->ev_done.push_back(b2);
    }
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRa, Ra, sizeof(double) * B, cudaMemcpyHostToDevice, pl->in_stream));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRas, Ras, sizeof(double) * B, cudaMemcpyHostToDevice, pl->in_stream));
    for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * chunk, nb = std::min(chunk, B - b0);
        const size_t off = (size_t)b0 * W, bytes = sizeof(double) * (size_t)nb * W;
        PLAN_CUDA(pl, cudaMemcpyAsync(pl->hX0 + off, Xin + off, bytes, cudaMemcpyHostToDevice, pl->in_stream));
        PLAN_CUDA(pl, cudaEventRecord(pl->ev_in[c], pl->in_stream));
        PLAN_CUDA(pl, cudaStreamWaitEvent(pl->own_stream, pl->ev_in[c], 0));
        // xtmp is indexed from member 0 inside sddc_step; chunks run back to back on own_stream, so it is free
        if ((rc = sddc_step(pl, pl->hX0 + off, pl->hX1 + off, pl->hRa + b0, pl->hRas + b0, nb, nsteps, linear,
                            pl->own_stream)))
            return rc;
        if (diag_out && (rc = sddc_diagnostics(pl, pl->hX1 + off, pl->hDiag + (size_t)b0 * 6, nb, pl->own_stream)))
            return rc;
        PLAN_CUDA(pl, cudaEventRecord(pl->ev_done[c], pl->own_stream));
        PLAN_CUDA(pl, cudaStreamWaitEvent(pl->out_stream, pl->ev_done[c], 0));
        PLAN_CUDA(pl, cudaMemcpyAsync(Xout + off, pl->hX1 + off, bytes, cudaMemcpyDeviceToHost, pl->out_stream));
        if (diag_out)
            PLAN_CUDA(pl, cudaMemcpyAsync(diag_out + (size_t)b0 * 6, pl->hDiag + (size_t)b0 * 6, sizeof(double) * nb * 6,
                                          cudaMemcpyDeviceToHost, pl->out_stream));
    }
    PLAN_CUDA(pl, cudaStreamSynchronize(pl->out_stream));
    PLAN_CUDA(pl, cudaStreamSynchronize(pl->own_stream));
    return SDDC_OK;
}

// The loop of Main._Time_Step (Main.py:286-329) for an ensemble, driven from HOST buffers: the state goes to the
// device once, every `diag_every` steps the diagnostics of all members are computed and copied back to
// diag_hist[record][B][6] (asynchronously, overlapped with the following steps), every `ckpt_every` steps the state
// is copied back to ckpt[record][B][3N] (the reference's X_DATA checkpoints), and the final state lands in Xout.
int sddc_time_step_host(sddc_plan* pl, const double* Xin, double* Xout, const double* Ra, const double* Ras, int B,
                        int nsteps, int linear, int diag_every, double* diag_hist, int ckpt_every, double* ckpt) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if (nsteps < 1 || diag_every < 0 || ckpt_every < 0 || (diag_every > 0 && !diag_hist) || (ckpt_every > 0 && !ckpt)) {
        pl->err = "bad arguments to sddc_time_step_host";
        return SDDC_ERR_INVALID;
    }
    if ((rc = ensure_host_staging(pl))) return rc;
    const size_t W = 3 * (size_t)pl->g.N, bytes = sizeof(double) * (size_t)B * W;
    const int nrec = diag_every ? nsteps / diag_every : 0;
    if (nrec > pl->hist_cap) {
        PLAN_CUDA(pl, cudaDeviceSynchronize());
        if (pl->hHist) cudaFree(pl->hHist);  // stays registered in allocs only once: allocate fresh, track manually
        void* q = nullptr;
        PLAN_CUDA(pl, cudaMalloc(&q, sizeof(double) * (size_t)nrec * pl->cfg.max_batch * 6));
        pl->hHist = static_cast<double*>(q);
        pl->hist_cap = nrec;
    }
    if (!pl->ev_ckpt) PLAN_CUDA(pl, cudaEventCreateWithFlags(&pl->ev_ckpt, cudaEventDisableTiming));
    if (pl->ev_in.empty()) {
        cudaEvent_t a, b2;
        PLAN_CUDA(pl, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        PLAN_CUDA(pl, cudaEventCreateWithFlags(&b2, cudaEventDisableTiming));
        pl->ev_in.push_back(a); pl->ev_done.push_back(b2);
    }
    cudaStream_t cs = pl->own_stream;
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRa, Ra, sizeof(double) * B, cudaMemcpyHostToDevice, cs));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRas, Ras, sizeof(double) * B, cudaMemcpyHostToDevice, cs));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hX0, Xin, bytes, cudaMemcpyHostToDevice, cs));
    double* cur = pl->hX0;
    double* nxt = pl->hX1;
    bool ckpt_pending = false;
    // With the FFT formulation the prep stage of step s+1 already holds J_theta(psi) and Dr psi of X_s as spectral rows:
    // the kinetic energy of X_s is taken from them (no second scan / derivative pass); only the last record needs the
    // stand-alone diagnostics.  (Symmetric runs: stepped states are masked already, so the masked prep loads are exact.)
    const bool share = pl->fft_M != 0 && pl->ke_M != 0 && !linear;
    int pending = -1;   // diagnostics record of the current state still to be produced
    auto ship_record = [&](int r, cudaStream_t from) -> int {
        double* drec = pl->hHist + (size_t)r * B * 6;
        PLAN_CUDA(pl, cudaEventRecord(pl->ev_done[0], from));
        PLAN_CUDA(pl, cudaStreamWaitEvent(pl->out_stream, pl->ev_done[0], 0));
        PLAN_CUDA(pl, cudaMemcpyAsync(diag_hist + (size_t)r * B * 6, drec, sizeof(double) * B * 6,
                                      cudaMemcpyDeviceToHost, pl->out_stream));
        return SDDC_OK;
    };
    // kinetic energy of a record on the side stream, under the back-substitution of the following step (see sddc_time_step)
    const bool side = share && pl->ke_stream != nullptr;
    bool ke_inflight = false;
    const size_t dstride = (size_t)pl->cfg.max_batch * 18;
    for (int s = 1; s <= nsteps; ++s) {
        if (ke_inflight) { PLAN_CUDA(pl, cudaStreamWaitEvent(cs, pl->ev_ke, 0)); ke_inflight = false; }
        if ((rc = run_step_prep(pl, cur, pl->hRa, pl->hRas, B, linear != 0, cs, s > 1))) return rc;
        double* rec = pending >= 0 ? pl->hHist + (size_t)pending * B * 6 : nullptr;
        const double* dp = pl->dpart + ((s - 1) & 1) * dstride;
        if (rec && side) PLAN_CUDA(pl, cudaEventRecord(pl->ev_rows, cs));
        if (rec && !side) {
            if ((rc = run_ke_fft(pl, cur, pl->coef7, 7LL * pl->g.K, 3 * pl->g.K, pl->ir, rec, B, cs, dp))) return rc;
            if ((rc = ship_record(pending, cs))) return rc;
        }
        const bool rec_shared = share && diag_every && s % diag_every == 0 && s < nsteps;
        if ((rc = run_step_rest(pl, nxt, nullptr, B, linear != 0, cs, s < nsteps, rec_shared ? pl->dpart + (s & 1) * dstride : nullptr))) return rc;
        if (rec && side) {
            PLAN_CUDA(pl, cudaStreamWaitEvent(pl->ke_stream, pl->ev_rows, 0));
            if ((rc = run_ke_fft(pl, cur, pl->coef7, 7LL * pl->g.K, 3 * pl->g.K, pl->ir, rec, B, pl->ke_stream, dp, true))) return rc;
            PLAN_CUDA(pl, cudaEventRecord(pl->ev_ke, pl->ke_stream));
            ke_inflight = true;
            if ((rc = ship_record(pending, pl->ke_stream))) return rc;
        }
        pending = -1;
        std::swap(cur, nxt);
        if (diag_every && s % diag_every == 0) {
            const int r = s / diag_every - 1;
            if (share && s < nsteps) {
                pending = r;
            } else {
                if (ke_inflight) { PLAN_CUDA(pl, cudaStreamWaitEvent(cs, pl->ev_ke, 0)); ke_inflight = false; }   // kepart is shared
                if ((rc = sddc_diagnostics(pl, cur, pl->hHist + (size_t)r * B * 6, B, cs))) return rc;
                if ((rc = ship_record(r, cs))) return rc;
            }
        }
        const int cph = pl->ckpt_phase > 0 ? pl->ckpt_phase : ckpt_every;   // step of the first checkpoint
        if (ckpt_every && s >= cph && (s - cph) % ckpt_every == 0) {
            const int r = (s - cph) / ckpt_every;
            if (ckpt_pending) PLAN_CUDA(pl, cudaStreamWaitEvent(cs, pl->ev_ckpt, 0));  // staging buffer free again
            PLAN_CUDA(pl, cudaMemcpyAsync(pl->hX2, cur, bytes, cudaMemcpyDeviceToDevice, cs));
            PLAN_CUDA(pl, cudaEventRecord(pl->ev_in[0], cs));
            PLAN_CUDA(pl, cudaStreamWaitEvent(pl->out_stream, pl->ev_in[0], 0));
            PLAN_CUDA(pl, cudaMemcpyAsync(ckpt + (size_t)r * B * W, pl->hX2, bytes, cudaMemcpyDeviceToHost, pl->out_stream));
            PLAN_CUDA(pl, cudaEventRecord(pl->ev_ckpt, pl->out_stream));
            ckpt_pending = true;
        }
    }
    PLAN_CUDA(pl, cudaMemcpyAsync(Xout, cur, bytes, cudaMemcpyDeviceToHost, cs));
    PLAN_CUDA(pl, cudaStreamSynchronize(cs));
    if (pl->ke_stream) PLAN_CUDA(pl, cudaStreamSynchronize(pl->ke_stream));
    PLAN_CUDA(pl, cudaStreamSynchronize(pl->out_stream));
    return SDDC_OK;
}

int sddc_plan_set_ckpt_phase(sddc_plan* pl, int first_step) {
    if (!pl || first_step < 0) return SDDC_ERR_INVALID;
    pl->ckpt_phase = first_step;
    return SDDC_OK;
}

int sddc_jvp_host(sddc_plan* pl, const double* dv, const double* X, double* out, const double* Ra, const double* Ras,
                  int B) {
    int rc = check_batch(pl, B);
    if (rc) return rc;
    if ((rc = ensure_host_staging(pl))) return rc;
    cudaStream_t st = pl->own_stream;
    const size_t bytes = sizeof(double) * (size_t)B * 3 * pl->g.N;
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hX0, dv, bytes, cudaMemcpyHostToDevice, st));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hX2, X, bytes, cudaMemcpyHostToDevice, st));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRa, Ra, sizeof(double) * B, cudaMemcpyHostToDevice, st));
    PLAN_CUDA(pl, cudaMemcpyAsync(pl->hRas, Ras, sizeof(double) * B, cudaMemcpyHostToDevice, st));
    if ((rc = sddc_jvp(pl, pl->hX0, pl->hX2, pl->hX1, pl->hRa, pl->hRas, B, st))) return rc;
    PLAN_CUDA(pl, cudaMemcpyAsync(out, pl->hX1, bytes, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(pl, cudaStreamSynchronize(st));
    return SDDC_OK;
}

}  // extern "C"
