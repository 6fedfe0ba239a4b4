// CPU emulation of one worker of nlin_fft_kernel (csrc/fft_core.h): the per-thread phase functions are the ones the
// device kernel runs; here they are executed thread by thread with the barriers of the kernel as loop boundaries.
// Test infrastructure only (tests/test_fft_core_cpu.py builds it with g++ and compares against the oracle).
#include <complex>
#include <cstdio>
#include <vector>

#include "../spectraldoublediffusiveconvection_b200/csrc/fft_core.h"

using namespace sddc::fftp;

template <int M, bool DFX>
static void run_rows(const double* coef0, const double* coef1, double* out, int nrows) {
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NF = DFX ? 9 : 5, NT = M == 768 ? 128 : 64;   // threads per worker as launched
    std::vector<double> tab(tab_doubles<M>());
    fill_tables<M>(tab.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)2 * NF * PL);
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;  // poison: every position that is read must have been written
        for (int t = 0; t < NT; ++t) {
            if (DFX) {
                build<M, 1, NT>(t, coef0 + (size_t)row * 7 * K, buf.data(), tb, coef1 + (size_t)row * 7 * K);
                build<M, 2, NT>(t, coef1 + (size_t)row * 7 * K, buf.data() + 10 * PL, tb);
            } else {
                build<M, 0, NT>(t, coef0 + (size_t)row * 7 * K, buf.data(), tb);
            }
        }
        for (int t = 0; t < NT; ++t) pass_c<M, NF, +1, NT>(t, buf.data());
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, NF, +1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) i3f1<M, DFX, NT>(t, buf.data(), tb);
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 2, -1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) pass_c<M, 2, -1, NT>(t, buf.data());
        for (int t = 0; t < NT; ++t) post<M, NT>(t, buf.data(), out + (size_t)row * 4 * K, tb);
    }
}

// the two-state kernel that transforms the perturbation two fields at a time (14 planes per worker)
template <int M>
static void run_rows_dfx2(const double* coef0, const double* coef1, double* out, int nrows) {
    constexpr int K = Cfg<M>::K, PL = Cfg<M>::PL, NT = Cfg<M>::L;   // one radix-6 column per thread
    std::vector<double> tab(tab_doubles<M>());
    fill_tables<M>(tab.data());
    const Tables tb = make_tables<M>(tab.data());
    std::vector<double> buf((size_t)14 * PL);
    std::vector<Dfx2State> st(NT);
    auto twd = [&](int t, C (&tw)[Cfg<M>::RD]) { load_tw<M>(t, tb, tw); };
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;
        const double* r0 = coef0 + (size_t)row * 7 * K;
        const double* r1 = coef1 + (size_t)row * 7 * K;
        double* pp = buf.data() + 10 * PL;
        for (int t = 0; t < NT; ++t) { build<M, 1, NT>(t, r0, buf.data(), tb, r1); build<M, 3, NT>(t, r1, pp, tb); }
        for (int t = 0; t < NT; ++t) pass_c<M, 7, +1, NT>(t, buf.data());
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; twd(t, tw); pass_d<M, 7, +1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) dfx2_first<M>(t, buf.data(), tb, st[t]);
        for (int t = 0; t < NT; ++t) build<M, 4, NT>(t, r1, pp, tb);
        for (int t = 0; t < NT; ++t) pass_c<M, 2, +1, NT>(t, pp);
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; twd(t, tw); pass_d<M, 2, +1, NT>(t, pp, tw); }
        for (int t = 0; t < NT; ++t) dfx2_second<M>(t, buf.data(), tb, st[t]);
        for (int t = 0; t < NT; ++t) { C tw[Cfg<M>::RD]; twd(t, tw); pass_d<M, 2, -1, NT>(t, buf.data(), tw); }
        for (int t = 0; t < NT; ++t) pass_c<M, 2, -1, NT>(t, buf.data());
        for (int t = 0; t < NT; ++t) post<M, NT>(t, buf.data(), out + (size_t)row * 4 * K, tb);
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
    std::vector<double> buf((size_t)2 * PL);
    for (int row = 0; row < nrows; ++row) {
        for (auto& v : buf) v = 1e300;
        const double* r = rows + (size_t)row * 2 * Kc;
        for (int t = 0; t < NTW; ++t) build_ke<M>(t, r, r + Kc, 1.0, buf.data(), tb);
        for (int t = 0; t < NTW; ++t) pass_c<M, 1, +1>(t, buf.data());
        for (int t = 0; t < NTW; ++t) { C tw[Cfg<M>::RD]; load_tw<M>(t, tb, tw); pass_d<M, 1, +1>(t, buf.data(), tw); }
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

// two-state products, perturbation transformed two fields at a time (M = 384, 768)
int fft_emul_rows_dfx2(int M, const double* coef0, const double* coef1, double* out, int nrows) {
    switch (M) {
        case 384: run_rows_dfx2<384>(coef0, coef1, out, nrows); return 0;
        case 768: run_rows_dfx2<768>(coef0, coef1, out, nrows); return 0;
        default: return -1;
    }
}

// coef0 / coef1: [nrows][7][K]; out: [nrows][4][K].  Returns 0, or -1 for an unsupported grid size.
int fft_emul_rows(int M, int dfx, const double* coef0, const double* coef1, double* out, int nrows) {
    switch (M) {
        case 192: dfx ? run_rows<192, true>(coef0, coef1, out, nrows) : run_rows<192, false>(coef0, coef1, out, nrows); return 0;
        case 384: dfx ? run_rows<384, true>(coef0, coef1, out, nrows) : run_rows<384, false>(coef0, coef1, out, nrows); return 0;
        case 768: dfx ? run_rows<768, true>(coef0, coef1, out, nrows) : run_rows<768, false>(coef0, coef1, out, nrows); return 0;
        default: return -1;
    }
}
}
