// Table generation, diagnostics reductions and the generic (drop-in) transforms.
#pragma once
#include "common.cuh"

namespace sddc {

// cos / sin of k * theta_j, theta_j = pi (2j+1) / (2M), with exact integer argument reduction.
__device__ __forceinline__ void trig_kj(long long k, long long j, long long M, double& c, double& s) {
    const long long q = (k * (2 * j + 1)) % (4 * M);  // angle = pi * q / (2M)
    sincospi((double)q / (double)(2 * M), &s, &c);
}

// Tile-major trigonometric tables (one contiguous block per GEMM pipeline stage):
// mode 0 (synthesis):  tab[jt][chunk][ks][type][par][col (W)][kk]   k' = 8 chunk + 4 ks + kk,  j' = jt W + col
//                      value = trig((2k'+par) theta_j')
// mode 1 (analysis):   tab[par][kt][chunk][ks][type][col (W)][kk]   j' = 8 chunk + 4 ks + kk,  k' = kt W + col
//                      value = scale_k * trig((2k'+par) theta_j'),  scale 2/M (1/M for the cosine k = 0 row)
// entries with k' >= Kh or j' >= Mh are zero; the sine k = 0 row is zero.  nA = number of tiles of the tiled
// index (jt or kt), nchunk = chunks of the contraction index.
__global__ void fill_table_kernel(double* tab, int mode, int M, int Kh, int Mh, int W, int nA, int nchunk) {
    const long long total = 8LL * nA * nchunk * W * 4;  // (2 ks * 2 types * 2 par) * tiles * chunks * W * 4
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int kk = (int)(r & 3); r >>= 2;
        const int col = (int)(r % W); r /= W;
        int type, par, ks, chunk, tile;
        if (mode == 0) {
            par = (int)(r & 1); r >>= 1; type = (int)(r & 1); r >>= 1; ks = (int)(r & 1); r >>= 1;
            chunk = (int)(r % nchunk); tile = (int)(r / nchunk);
        } else {
            type = (int)(r & 1); r >>= 1; ks = (int)(r & 1); r >>= 1;
            chunk = (int)(r % nchunk); r /= nchunk; tile = (int)(r % nA); par = (int)(r / nA);
        }
        const int inner = 8 * chunk + 4 * ks + kk, outer = tile * W + col;
        // synthesis tables follow the coefficient order inside a chunk (chunk_pos)
        const int kp = (mode == 0) ? chunk_pos_inv(inner, par) : outer, jp = (mode == 0) ? outer : inner;
        double v = 0.0;
        if (kp < Kh && jp < Mh) {
            const int k = 2 * kp + par;
            double c, s;
            trig_kj(k, jp, M, c, s);
            v = type == 0 ? c : s;
            if (k == 0 && type == 1) v = 0.0;
            if (mode == 1) v *= (type == 0 && k == 0) ? (1.0 / M) : (2.0 / M);
        }
        tab[idx] = v;
    }
}

// Tables of the second mirror level ("quarter-wave" split, k_synth_ws.cuh).  Mirror pairs j' come in orbits
// (L = j'', R = Mh-1-j''), j'' < Mq = Mh/2, with theta_R = pi/2 - theta_L.
// mode 2 (synthesis, W = 16 columns per tile = 8 orbits):  tab[jt][chunk][ks][type][par][16][kk]
//     par 1: col < 8 -> angle theta_L(jt*8+col), col >= 8 -> theta_R(jt*8+col-8);  k' = 8 chunk + 4 ks + kk
//     par 0: col < 8 -> theta_L, k' = chunk_pos_inv(8 chunk + 4 ks + kk) (ks 0: k' even, ks 1: k' odd); col >= 8 -> 0
// mode 3 (analysis, 64 output columns per tile): tab[par][kt][chunk][ks][type][64][kk], contraction position
//     p = 8 chunk + 4 ks + kk
//     par 1: p < Mq -> theta_L(p), p >= Mq -> theta_R(p - Mq);  k' = kt*64 + col
//     par 0: only p < Mq (theta_L(p)); output columns are class-ordered: class = kt / (nkt/2),
//            k' = 2*((kt % (nkt/2))*64 + col) + class
__global__ void fill_table_quarter_kernel(double* tab, int mode, int M, int Kh, int Mh, int W, int nA, int nchunk) {
    const long long total = 8LL * nA * nchunk * W * 4;
    const int Mq = Mh / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int kk = (int)(r & 3); r >>= 2;
        const int col = (int)(r % W); r /= W;
        int type, par, ks, chunk, tile;
        if (mode == 2) {
            par = (int)(r & 1); r >>= 1; type = (int)(r & 1); r >>= 1; ks = (int)(r & 1); r >>= 1;
            chunk = (int)(r % nchunk); tile = (int)(r / nchunk);
        } else {
            type = (int)(r & 1); r >>= 1; ks = (int)(r & 1); r >>= 1;
            chunk = (int)(r % nchunk); r /= nchunk; tile = (int)(r % nA); par = (int)(r / nA);
        }
        const int inner = 8 * chunk + 4 * ks + kk;
        int kp = -1, jp = -1;   // wavenumber index k' and mirror-pair index j' (angle theta_{j'}); -1 = zero entry
        if (mode == 2) {
            const int jq = tile * 8 + (col & 7);
            if (jq < Mq) {
                if (par == 1) { kp = inner; jp = (col < 8) ? jq : Mh - 1 - jq; }
                else if (col < 8) { kp = chunk_pos_inv(inner, 0); jp = jq; }
            }
        } else {
            if (par == 1) {
                kp = tile * W + col;
                if (inner < Mq) jp = inner; else if (inner < 2 * Mq) jp = Mh - 1 - (inner - Mq);
            } else {
                const int half = nA / 2, cls = tile / half;
                kp = 2 * ((tile % half) * W + col) + cls;
                if (inner < Mq) jp = inner;
            }
        }
        double v = 0.0;
        if (kp >= 0 && jp >= 0 && kp < Kh && jp < Mh) {
            const int k = 2 * kp + par;
            double c, s;
            trig_kj(k, jp, M, c, s);
            v = type == 0 ? c : s;
            if (k == 0 && type == 1) v = 0.0;
            if (mode == 3) v *= (type == 0 && k == 0) ? (1.0 / M) : (2.0 / M);
        }
        tab[idx] = v;
    }
}

// w[j'] = trapezoid weight(theta_j') * sin(theta_j') on the M3-point grid for j' < M3/2 (mirror point has the
// same value); np.trapz with x = theta on interior nodes only (Main.py:117,130).
__global__ void fill_ke_weights_kernel(double* w, int M3, int Jp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Jp) return;
    double v = 0.0;
    if (j < M3 / 2) {
        const double th = M_PI * (2.0 * j + 1.0) / (2.0 * M3);
        const double dth = M_PI / M3;
        v = ((j == 0) ? 0.5 * dth : dth) * sin(th);
    }
    w[j] = v;
}

// Coefficient rows for the kinetic-energy synthesis: field 0 = J_theta(psi)/r (cosine), field 1 = Dr psi in
// sinusoid indexing (sine).  Tile-major layout [B][Khp/8][2 ks][2 par][2*n8][4].  (Main.py:104-115)
struct KEPrepParams {
    const double* X; long long x_stride;
    const double* JJ;
    double* coef; long long coef_stride;
    const double* DrT;  // [n][n8]
    const double* ir;   // [n] 1/r
    Geo g;
    int rows;           // 1: row-major output [B][n][2][K] for the FFT formulation (k_nlin_fft.cuh)
};

__global__ void __launch_bounds__(256) ke_prep_kernel(KEPrepParams p) {
    extern __shared__ __align__(128) double smem[];
    const Geo& g = p.g;
    const int n = g.n, n8 = g.n8, K = g.K;
    const int b = blockIdx.y, c0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sP = smem;           // psi blocks c0-1 .. c0+30 (q = lane)
    double* mDr = sP + 32 * n;   // [n][n8]
    const double* Xb = p.X + (long long)b * p.x_stride;
    for (int idx = tid; idx < 32 * n; idx += 256) {
        const int bp = c0 - 1 + idx / n;
        double v = 0.0;
        // Kinetic_Energy differentiates every psi block, symmetric or not (Main.py:109-111); only the
        // J_theta part is masked (through the scan)
        if (bp >= 0 && bp < K) v = Xb[(long long)bp * n + (idx % n)];
        sP[idx] = v;
    }
    for (int idx = tid; idx < n * n8; idx += 256) mDr[idx] = p.DrT[idx];
    __syncthreads();
    const int c = c0 + lane;
    if (c >= K) return;
    const int par = c & 1, kp = chunk_pos(c >> 1, par);
    double* cf = p.coef + (long long)b * p.coef_stride;
    const double* Jb = p.JJ + (long long)b * (K + 1) * n;
    for (int i = warp; i < n; i += 8) {
        double d = 0.0;
        for (int ip = 0; ip < n; ++ip) d = fma(mDr[ip * n8 + i], sP[lane * n + ip], d);
        if (p.rows) {
            double* r2 = p.coef + (((long long)b * n + i) * 2) * K + c;
            r2[0] = p.ir[i] * Jb[(long long)c * n + i];
            r2[K] = (c >= 1) ? d : 0.0;
            continue;
        }
        const long long o = (((long long)((kp >> 2) * 2 + par) * (2 * n8)) + i) * 4 + (kp & 3);
        cf[o] = p.ir[i] * Jb[(long long)c * n + i];
        cf[o + (long long)n8 * 4] = (c >= 1) ? d : 0.0;
    }
}

// One CTA per member: ||X||_2, Nusselt numbers at both walls, and the final kinetic-energy reduction.
// out[b] = { norm, KE, Nu_T, Nu_S, Nu_T(outer), Nu_S(outer) }   (Main.py:41-68, 292-295)
__global__ void __launch_bounds__(256) diag_kernel(const double* __restrict__ X, const double* __restrict__ kepart,
                                                   int nke, const double* __restrict__ nu_in,
                                                   const double* __restrict__ nu_out, double ke_scale, Geo g,
                                                   double* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x, n = g.n, N = g.N;
    const double* Xb = X + (long long)b * 3 * N;
    double s2 = 0.0, nt_i = 0.0, ns_i = 0.0, nt_o = 0.0, ns_o = 0.0;
    for (int idx = tid; idx < 3 * N; idx += 256) {
        const double v = Xb[idx];
        s2 = fma(v, v, s2);
    }
    for (int idx = tid; idx < N; idx += 256) {
        const int k = idx / n, i = idx - k * n;
        if ((k & 1) == 0) {
            const double w = 1.0 / (1.0 - (double)k * (double)k);
            const double t = Xb[N + idx] * w, s = Xb[2 * N + idx] * w;
            nt_i = fma(nu_in[i], t, nt_i);  ns_i = fma(nu_in[i], s, ns_i);
            nt_o = fma(nu_out[i], t, nt_o); ns_o = fma(nu_out[i], s, ns_o);
        }
    }
    double ke = 0.0;
    for (int idx = tid; idx < nke; idx += 256) ke += kepart[(long long)b * nke + idx];
    s2 = block_sum(s2, red);
    nt_i = block_sum(nt_i, red); ns_i = block_sum(ns_i, red);
    nt_o = block_sum(nt_o, red); ns_o = block_sum(ns_o, red);
    ke = block_sum(ke, red);
    if (tid == 0) {
        double* o = out + (long long)b * 6;
        o[0] = sqrt(s2); o[1] = ke_scale * ke; o[2] = nt_i; o[3] = ns_i; o[4] = nt_o; o[5] = ns_o;
    }
}

// The same record from partial sums: the back-substitution chains leave sum f^2 and the Nusselt sums of the state they
// produce in dpart[b][field * 2 + parity][3] (k_solve_hot.cuh, DIAG), the kinetic-energy transform its row sums in kepart.
// One warp per member; fixed summation order.
__global__ void __launch_bounds__(128) diag_finish_kernel(const double* __restrict__ dpart, const double* __restrict__ kepart,
                                                          int nke, double ke_scale, int B, double* __restrict__ out) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    double ke = 0.0;
    for (int idx = lane; idx < nke; idx += 32) ke += kepart[(long long)b * nke + idx];
    ke = warp_sum(ke);
    if (lane == 0) {
        const double* d = dpart + (long long)b * 18;
        double s2 = 0.0;
        for (int c = 0; c < 6; ++c) s2 += d[c * 3];
        double* o = out + (long long)b * 6;
        o[0] = sqrt(s2); o[1] = ke_scale * ke;
        o[2] = d[2 * 3 + 1] + d[3 * 3 + 1]; o[3] = d[4 * 3 + 1] + d[5 * 3 + 1];
        o[4] = d[2 * 3 + 2] + d[3 * 3 + 2]; o[5] = d[4 * 3 + 2] + d[5 * 3 + 2];
    }
}

// Generic drop-in transforms (Transforms.py:73-129) by direct summation; one thread per output element.
__global__ void transform_kernel(int kind, const double* __restrict__ in, double* __restrict__ out, int rows,
                                 int n_in, int n_out) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * n_out) return;
    const int r = (int)(idx / n_out), o = (int)(idx - (long long)r * n_out);
    const double* x = in + (long long)r * n_in;
    double acc = 0.0, c, s;
    if (kind == 0) {          // IDCT: f_j = sum_{k < min(K, M)} a_k cos(k theta_j), M = n_out
        const int ku = min(n_in, n_out);
        for (int k = 0; k < ku; ++k) { trig_kj(k, o, n_out, c, s); acc = fma(x[k], c, acc); }
    } else if (kind == 1) {   // IDST: g_j = sum_{1 <= k < min(K, M+1)} b_k sin(k theta_j)
        // a truncating call (n_out < n_in) keeps mode n_out with HALF weight: the reference's shifted, halved array is
        // cut to n_out entries and the DST-III takes its last input without the factor 2 (Transforms.py:41-54,87-100)
        const int ku = min(n_in, n_out + 1);
        for (int k = 1; k < ku; ++k) {
            trig_kj(k, o, n_out, c, s);
            acc = fma(k == n_out ? 0.5 * x[k] : x[k], s, acc);
        }
    } else if (kind == 2) {   // DCT: a_k = (2/M) sum_j f_j cos(k theta_j), a_0 halved, M = n_in
        for (int j = 0; j < n_in; ++j) { trig_kj(o, j, n_in, c, s); acc = fma(x[j], c, acc); }
        acc *= (o == 0 ? 1.0 : 2.0) / n_in;
    } else {                  // DST: b_k = (2/M) sum_j g_j sin(k theta_j), b_0 = 0
        if (o > 0) for (int j = 0; j < n_in; ++j) { trig_kj(o, j, n_in, c, s); acc = fma(x[j], s, acc); }
        acc *= 2.0 / n_in;
    }
    out[idx] = acc;
}

// Resolution transfer in r (INTERP_RADIAL, Matrix_Operators.py:901-941): the reference fits a polynomial through every
// radial profile and evaluates it on the new grid -- a LINEAR map of the nr_o values, the same for every profile, which
// the host builds once with the reference's own np.polyfit call (interp.py).  out[row][i] = sum_j W[i][j] in[row][j].
__global__ void __launch_bounds__(256) interp_radial_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                           const double* __restrict__ W, long long rows, int nr_o, int nr_n) {
    extern __shared__ double sW[];   // [nr_n][nr_o]
    for (int i = threadIdx.x; i < nr_n * nr_o; i += 256) sW[i] = W[i];
    __syncthreads();
    const long long tot = rows * nr_n;
    for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < tot; idx += (long long)gridDim.x * 256) {
        const long long row = idx / nr_n;
        const int i = (int)(idx - row * nr_n);
        const double* x = in + row * nr_o;
        const double* w = sW + i * nr_o;
        double acc = 0.0;
        for (int j = 0; j < nr_o; ++j) acc = fma(w[j], x[j], acc);
        out[idx] = acc;
    }
}

// Resolution transfer in theta (INTERP_THETAS, Matrix_Operators.py:944-1011).  The reference goes to the grid and back
// (IDST / IDCT on max(K_o, K_n) points, DST / DCT truncated to K_n), which on one and the same midpoint grid is the
// identity on every retained coefficient: the result is spectral zero-padding / truncation -- with one quirk kept, the
// stream-function block 0 comes back as zero because IDST ignores entry 0 of what it is handed (Transforms.py:41-54).
__global__ void interp_thetas_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int K_o, int K_n, int nr) {
    const long long per = 3LL * K_n * nr, tot = per * B;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / per;
        const long long r = idx - b * per;
        const int f = (int)(r / ((long long)K_n * nr));
        const int k = (int)((r - (long long)f * K_n * nr) / nr);
        const int i = (int)(r % nr);
        double v = 0.0;
        if (k < K_o && !(f == 0 && k == 0)) v = in[((b * 3 + f) * K_o + k) * nr + i];
        out[idx] = v;
    }
}

// out = a - b (residual / JVP helper for rows the solve does not touch is not needed; used for host tests)
__global__ void axpby_kernel(double* out, const double* a, const double* b, double alpha, double beta, long long nel) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nel; i += (long long)gridDim.x * blockDim.x)
        out[i] = alpha * a[i] + beta * b[i];
}

}  // namespace sddc
