// CPU emulation of one worker of nlin_fft_kernel (csrc/fft_core.h): the per-thread phase functions are the ones the
// device kernel runs; here they are executed thread by thread with the barriers of the kernel as loop boundaries.
// Test infrastructure only (tests/test_fft_core_cpu.py builds it with g++ and compares against the oracle).
#include <complex>
#include <cstdio>
#include <vector>

#include "../spectraldoublediffusiveconvection_b200/csrc/fft_core.h"
#include "../spectraldoublediffusiveconvection_b200/csrc/fft_fused.h"
#include <cstring>

using namespace sddc::fftp;

template <int M, bool DFX>
static void run_rows(const double* coef0, const double* coef1, double* out, int nrows) {
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NF = DFX ? 7 : 4, NT = M == 768 ? 128 : 64;   // threads per worker as launched
    std::vector<double> tab(tab_doubles<M>());
    fill_tables<M>(tab.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)pairs_doubles<M>(NF));
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;  // poison: every position that is read must have been written
        const double* r0 = coef0 + (size_t)row * 7 * K;
        const double* r1 = coef1 + (size_t)row * 7 * K;
        for (int t = 0; t < NT; ++t) {
            if (DFX) bc_inv_dfx<M, NT>(t, r0, r1, buf.data(), tb);
            else bc_inv_fx<M, NT>(t, r0, buf.data(), tb);
        }
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, NF, +1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) {
            if (DFX) i3f1_dfx<M, NT>(t, buf.data(), tb);
            else i3f1_fx<M, NT>(t, buf.data(), tb);
        }
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 2, -1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) cp_fwd<M, NT>(t, buf.data(), out + (size_t)row * 4 * K, tb);
    }
}

// cached base state: grid fields of rows0 by the GRID mode, then the bilinear products of (rows0, rows1) by the JVPC mode
template <int M>
static void run_rows_cached(const double* coef0, const double* coef1, double* out, int nrows) {
    constexpr int K = Cfg<M>::K, NT = M == 768 ? 128 : 64;
    std::vector<double> tab(tab_doubles<M>());
    fill_tables<M>(tab.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)pairs_doubles<M>(4)), grid((size_t)7 * M);
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;
        for (auto& v : grid) v = 1e300;
        for (int t = 0; t < NT; ++t) bc_inv_fx<M, NT>(t, coef0 + (size_t)row * 7 * K, buf.data(), tb);
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 4, +1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) i3f1_grid<M, NT>(t, buf.data(), tb, grid.data());
        for (auto& v : buf) v = 1e300;
        for (int t = 0; t < NT; ++t) bc_inv_fx<M, NT>(t, coef1 + (size_t)row * 7 * K, buf.data(), tb);
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 4, +1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) i3f1_jvpc<M, NT>(t, buf.data(), tb, grid.data());
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 2, -1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) cp_fwd<M, NT>(t, buf.data(), out + (size_t)row * 4 * K, tb);
    }
}

// staged one-state schedule (nlin_fft_staged_kernel): the two warps of a worker, each running ahead of the other as far
// as the two worker barriers allow (warp 1 completely before warp 0 in every barrier interval)
template <int M>
static void run_rows_staged(const double* coef0, double* out, int nrows) {
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NT = 64;
    std::vector<double> tab(tab_doubles<M>());
    fill_tables<M>(tab.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)8 * PL + 3 * K);
    for (auto& v : buf) v = 1e300;
    auto fetch = [&](int row, int warp) {   // the bulk copies lane 0 of `warp` issues
        const double* src = coef0 + (size_t)row * 7 * K;
        std::memcpy(buf.data() + (4 + 2 * warp) * PL, src + 2 * warp * K, sizeof(double) * 2 * K);
        if (warp == 0) std::memcpy(buf.data() + 8 * PL, src + 4 * K, sizeof(double) * 2 * K);
        else std::memcpy(buf.data() + 8 * PL + 2 * K, src + 6 * K, sizeof(double) * K);
    };
    fetch(0, 0); fetch(0, 1);
    for (int row = 0; row < nrows; ++row) {
        for (int wi = 0; wi < 2; ++wi) {
            const int warp = 1 - wi;
            for (int stage = 0; stage < 3; ++stage)
                for (int lane = 0; lane < 32; ++lane) { C tw[Cfg<M>::RD]; load_tw<M>(32 * warp + lane, tb, tw); staged_pack<M>(warp, lane, buf.data(), tb, tw, stage); }
        }
        for (int t = 0; t < NT; ++t) i3f1_fx<M, NT>(t, buf.data(), tb);
        for (int wi = 0; wi < 2; ++wi) {
            const int warp = 1 - wi;
            if (row + 1 < nrows) fetch(row + 1, warp);
            for (int stage = 0; stage < 2; ++stage)
                for (int lane = 0; lane < 32; ++lane) { C tw[Cfg<M>::RD]; load_tw<M>(32 * warp + lane, tb, tw); staged_unpack<M>(warp, lane, buf.data(), out + (size_t)row * 4 * K, tb, tw, stage); }
        }
    }
}

template <int R, int SIGN>
static double check_dft() {
    C x[R], y[R];
    for (int j = 0; j < R; ++j) x[j] = C{0.3 + 0.7 * j - 0.05 * j * j, -0.2 + 0.11 * j * j};
    Dft<R, SIGN>::run(x, y);
    double err = 0;
    for (int k = 0; k < R; ++k) {
        std::complex<double> s = 0;
        for (int j = 0; j < R; ++j) s += std::complex<double>(x[j].r, x[j].i) * std::polar(1.0, SIGN * 2 * M_PI * ((j * k) % R) / R);
        err = std::max(err, std::abs(s - std::complex<double>(y[k].r, y[k].i)));
    }
    return err;
}
template <int SIGN>
static double check_dft6() {
    C x[6], y[6];
    for (int j = 0; j < 6; ++j) x[j] = C{0.3 + 0.7 * j - 0.05 * j * j, -0.2 + 0.11 * j * j};
    dft6<SIGN>(x, y);
    double err = 0;
    for (int k = 0; k < 6; ++k) {
        std::complex<double> s = 0;
        for (int j = 0; j < 6; ++j) s += std::complex<double>(x[j].r, x[j].i) * std::polar(1.0, SIGN * 2 * M_PI * ((j * k) % 6) / 6);
        err = std::max(err, std::abs(s - std::complex<double>(y[k].r, y[k].i)));
    }
    return err;
}

template <int M>
static void run_ke_rows(const double* rows, double* out, int nrows) {
    constexpr int Kc = M / 3, PL = Cfg<M>::PL;
    std::vector<double> tab(tab_doubles<M>()), Wn(M);
    fill_tables<M>(tab.data());
    fill_ke_weights<M>(Wn.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)2 * PL), srow((size_t)2 * Cfg<M>::K);
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;
        for (auto& v : srow) v = 1e300;
        const double* r = rows + (size_t)row * 2 * Kc;
        for (int t = 0; t < NTW; ++t) ke_stage<M>(t, r, r + Kc, 1.0, srow.data());
        for (int t = 0; t < NTW; ++t) ke_pack<M>(t, srow.data(), buf.data(), tb);
        if constexpr (Cfg<M>::RD == 32) {
            for (int t = 0; t < NTW; ++t) pass_d32_a<M>(t, buf.data(), tb);
            for (int t = 0; t < NTW; ++t) pass_d32_b<M>(t, buf.data());
        } else {
            for (int t = 0; t < NTW; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 1, +1>(t, buf.data(), tw); }
        }
        double s = 0.0;
        for (int t = 0; t < NTW; ++t) s += ke6<M>(t, buf.data(), tb, Wn.data());
        out[row] = s;
    }
}

extern "C" {

// rows: [nrows][2][M/3] (cosine-type, sine-type coefficients); out[row] = sum_j w_j sin(th_j) (f_j^2 + g_j^2)
int fft_emul_ke_rows(int M, const double* rows, double* out, int nrows) {
    switch (M) {
        case 384: run_ke_rows<384>(rows, out, nrows); return 0;
        case 768: run_ke_rows<768>(rows, out, nrows); return 0;
        case 1536: run_ke_rows<1536>(rows, out, nrows); return 0;
        default: return -1;
    }
}

// worst absolute error of the register butterflies against a direct DFT
double fft_emul_butterfly_error() {
    double e = 0;
    e = std::max(e, check_dft<4, +1>());
    e = std::max(e, check_dft<4, -1>());
    e = std::max(e, check_dft<8, +1>());
    e = std::max(e, check_dft<8, -1>());
    e = std::max(e, check_dft<16, +1>());
    e = std::max(e, check_dft<16, -1>());
    e = std::max(e, check_dft6<+1>());
    e = std::max(e, check_dft6<-1>());
    return e;
}

// staged schedule of the one-state kernel (M = 384)
int fft_emul_rows_staged(int M, const double* coef0, double* out, int nrows) {
    if (M != 384) return -1;
    run_rows_staged<384>(coef0, out, nrows);
    return 0;
}

// coef0 / coef1: [nrows][7][K]; out: [nrows][4][K].  Returns 0, or -1 for an unsupported grid size.
int fft_emul_rows(int M, int dfx, const double* coef0, const double* coef1, double* out, int nrows) {
    if (dfx == 2) {   // two-state products through the cached base grid (GRID + JVPC modes)
        switch (M) {
            case 192: run_rows_cached<192>(coef0, coef1, out, nrows); return 0;
            case 384: run_rows_cached<384>(coef0, coef1, out, nrows); return 0;
            case 768: run_rows_cached<768>(coef0, coef1, out, nrows); return 0;
            default: return -1;
        }
    }
    switch (M) {
        case 192: dfx ? run_rows<192, true>(coef0, coef1, out, nrows) : run_rows<192, false>(coef0, coef1, out, nrows); return 0;
        case 384: dfx ? run_rows<384, true>(coef0, coef1, out, nrows) : run_rows<384, false>(coef0, coef1, out, nrows); return 0;
        case 768: dfx ? run_rows<768, true>(coef0, coef1, out, nrows) : run_rows<768, false>(coef0, coef1, out, nrows); return 0;
        default: return -1;
    }
}
}
