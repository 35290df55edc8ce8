// Generic-size Green's-function blocks, one CTA per energy (any n <= 128; the two work blocks sit in shared memory up to
// n = 80 and in a global scratch beyond):
//   tx_green_column0  first block column G[:,0]  -> wavefunction psi_j = i G[j,0] (A o v_L)
//                     (transport/wfm/__init__.py:98-99)
//   tx_green_full     all M x M blocks of inv(E - H')           (wfm.Green, :139-166)
//   tx_hmat           dense block-pentadiagonal assembly         (wfm.Hmat,  :168-215)
// The reference obtains G by a dense (M n)^3 inverse and reshapes it with a 4-deep Python loop
// (transport/tdfci/utils.py:290-316); here the blocks come from left/right recursive sweeps.
#include <cuda_runtime.h>

#include <cstdio>

#include "block_ops.cuh"
#include "tx_host.h"

using namespace tx;

namespace {

struct GreenParams {
    const cd* h;      // [M][n][n]
    const cd* tnn;    // [M-1][n][n]
    const cd* E;      // [nE]
    const cd* sigL;   // [nE][n][n]
    const cd* sigR;
    cd* out;          // column0: [nE][M][n][n]; full: [nE][M][M][n][n]
    cd* work;         // full: [nE][2][M][n][n]  (gL, gR)
    cd* ax;           // [nE][2][n][n] work blocks when they do not fit in shared memory, else null
    int* singular;
    int n, M, nE;
};

constexpr int GR_THREADS = 128;

__global__ void __launch_bounds__(GR_THREADS) green_column0_kernel(const GreenParams p) {
    extern __shared__ double2 gr_smem[];
    __shared__ int s_ipiv[TX_MAX_N];
    __shared__ cd s_colk[TX_MAX_N];
    const int n = p.n, M = p.M, nn = n * n;
    const int e = blockIdx.x;
    const cd E = p.E[e];
    cd* A = p.ax ? p.ax + (long long)e * 2 * nn : reinterpret_cast<cd*>(gr_smem);
    cd* X = A + nn;
    cd* Gc = p.out + (long long)e * M * nn;
    const cd* SL = p.sigL + (long long)e * nn;
    const cd* SR = p.sigR + (long long)e * nn;

    // right-connected g^R_j, stored in the output slots
    for (int j = M - 1; j >= 0; --j) {
        cta_z_minus(A, E, p.h + (long long)j * nn, n);
        if (j == M - 1) cta_sub(A, SR, nn);
        if (j == 0) cta_sub(A, SL, nn);
        if (j < M - 1) {
            const cd* t = p.tnn + (long long)j * nn;
            cta_gemm(X, t, OP_N, Gc + (long long)(j + 1) * nn, OP_N, n);     // t_j g_{j+1}
            cta_gemm_acc(A, X, OP_N, t, OP_H, n, -1.0);                      // - (.) t_j^dag
        }
        cta_inverse(A, n, s_colk, s_ipiv, p.singular);
        cta_copy(Gc + (long long)j * nn, A, nn);
    }
    // G_{j,0} = g^R_j t_{j-1}^dag G_{j-1,0}
    for (int j = 1; j < M; ++j) {
        const cd* t = p.tnn + (long long)(j - 1) * nn;
        cta_gemm(X, t, OP_H, Gc + (long long)(j - 1) * nn, OP_N, n);
        cta_gemm(A, Gc + (long long)j * nn, OP_N, X, OP_N, n);
        cta_copy(Gc + (long long)j * nn, A, nn);
    }
}

__global__ void __launch_bounds__(GR_THREADS) green_full_kernel(const GreenParams p) {
    extern __shared__ double2 gr_smem[];
    __shared__ int s_ipiv[TX_MAX_N];
    __shared__ cd s_colk[TX_MAX_N];
    const int n = p.n, M = p.M, nn = n * n;
    const int e = blockIdx.x;
    const cd E = p.E[e];
    cd* A = p.ax ? p.ax + (long long)e * 2 * nn : reinterpret_cast<cd*>(gr_smem);
    cd* X = A + nn;
    cd* G = p.out + (long long)e * M * M * nn;
    cd* gL = p.work + (long long)e * 2 * M * nn;
    cd* gR = gL + (long long)M * nn;
    const cd* SL = p.sigL + (long long)e * nn;
    const cd* SR = p.sigR + (long long)e * nn;
    auto blk = [&](int r, int c) { return G + ((long long)r * M + c) * nn; };

    for (int j = 0; j < M; ++j) {                       // left-connected
        cta_z_minus(A, E, p.h + (long long)j * nn, n);
        if (j == 0) cta_sub(A, SL, nn);
        if (j == M - 1) cta_sub(A, SR, nn);
        if (j > 0) {
            const cd* t = p.tnn + (long long)(j - 1) * nn;
            cta_gemm(X, t, OP_H, gL + (long long)(j - 1) * nn, OP_N, n);
            cta_gemm_acc(A, X, OP_N, t, OP_N, n, -1.0);
        }
        cta_inverse(A, n, s_colk, s_ipiv, p.singular);
        cta_copy(gL + (long long)j * nn, A, nn);
    }
    for (int j = M - 1; j >= 0; --j) {                  // right-connected
        cta_z_minus(A, E, p.h + (long long)j * nn, n);
        if (j == 0) cta_sub(A, SL, nn);
        if (j == M - 1) cta_sub(A, SR, nn);
        if (j < M - 1) {
            const cd* t = p.tnn + (long long)j * nn;
            cta_gemm(X, t, OP_N, gR + (long long)(j + 1) * nn, OP_N, n);
            cta_gemm_acc(A, X, OP_N, t, OP_H, n, -1.0);
        }
        cta_inverse(A, n, s_colk, s_ipiv, p.singular);
        cta_copy(gR + (long long)j * nn, A, nn);
    }
    for (int j = 0; j < M; ++j) {                       // diagonal blocks
        cta_z_minus(A, E, p.h + (long long)j * nn, n);
        if (j == 0) cta_sub(A, SL, nn);
        if (j == M - 1) cta_sub(A, SR, nn);
        if (j > 0) {
            const cd* t = p.tnn + (long long)(j - 1) * nn;
            cta_gemm(X, t, OP_H, gL + (long long)(j - 1) * nn, OP_N, n);
            cta_gemm_acc(A, X, OP_N, t, OP_N, n, -1.0);
        }
        if (j < M - 1) {
            const cd* t = p.tnn + (long long)j * nn;
            cta_gemm(X, t, OP_N, gR + (long long)(j + 1) * nn, OP_N, n);
            cta_gemm_acc(A, X, OP_N, t, OP_H, n, -1.0);
        }
        cta_inverse(A, n, s_colk, s_ipiv, p.singular);
        cta_copy(blk(j, j), A, nn);
    }
    for (int c = 0; c < M; ++c) {                       // off-diagonal blocks, column by column
        for (int r = c + 1; r < M; ++r) {
            cta_gemm(X, p.tnn + (long long)(r - 1) * nn, OP_H, blk(r - 1, c), OP_N, n);
            cta_gemm(blk(r, c), gR + (long long)r * nn, OP_N, X, OP_N, n);
        }
        for (int r = c - 1; r >= 0; --r) {
            cta_gemm(X, p.tnn + (long long)r * nn, OP_N, blk(r + 1, c), OP_N, n);
            cta_gemm(blk(r, c), gL + (long long)r * nn, OP_N, X, OP_N, n);
        }
    }
}

__global__ void hmat_kernel(const cd* __restrict__ h, const cd* __restrict__ tnn, const cd* __restrict__ tnnn,
                            cd* __restrict__ H, int n, int M) {
    const long long dim = (long long)n * M;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dim * dim) return;
    const int row = (int)(idx / dim), col = (int)(idx - (long long)row * dim);
    const int si = row / n, a = row - si * n;
    const int sj = col / n, b = col - sj * n;
    const long long nn = (long long)n * n;
    cd v = mk(0.0, 0.0);
    if (si == sj) v = h[si * nn + a * n + b];                       // wfm/__init__.py:200-201
    else if (si == sj + 1) v = conj(tnn[sj * nn + b * n + a]);      // :203-204
    else if (si + 1 == sj) v = tnn[si * nn + a * n + b];            // :206-207
    else if (si == sj + 2) v = conj(tnnn[sj * nn + b * n + a]);     // :209-210
    else if (si + 2 == sj) v = tnnn[si * nn + a * n + b];           // :212-213
    H[idx] = v;
}

// bardeen.Hsysmat (transport/bardeen/__init__.py:690-750): regions  inf | L | C | R | inf.
struct HsysParams {
    const cd *tinf, *tL, *tR, *VinfL, *VinfR, *VL, *VR, *HC;
    cd* H;
    int Ninf, NL, NR, NC, n;
};

__global__ void hsysmat_kernel(const HsysParams p) {
    const int n = p.n, lc = p.NC / 2;
    const int minusinf = -lc - p.NL - p.Ninf;
    const int S = 2 * p.Ninf + p.NL + p.NR + p.NC;
    const long long total = (long long)S * S * n * n;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int b = (int)(idx % n);
    const int a = (int)((idx / n) % n);
    const int k = (int)((idx / ((long long)n * n)) % S);
    const int i = (int)(idx / ((long long)n * n * S));
    const int ji = i + minusinf, jk = k + minusinf;
    const int ab = a * n + b;
    cd v = mk(0.0, 0.0);
    auto hop_left_type = [&](int j, cd& acc) {      // site j in the far-left / left-well region (:724-729)
        if (j < -p.NL - lc) acc = acc - p.tinf[ab];
        else if (j < -lc) acc = acc - p.tL[ab];
    };
    auto hop_right_type = [&](int j, cd& acc) {     // site j in the right-well / far-right region (:730-735)
        if (j > lc + p.NR) acc = acc - p.tinf[ab];
        else if (j > lc) acc = acc - p.tR[ab];
    };
    if (ji >= -lc && ji <= lc && jk >= -lc && jk <= lc) {
        v = p.HC[(((long long)(ji + lc) * p.NC + (jk + lc)) * n + a) * n + b];      // :748
    } else if (i == k) {
        if (ji < -p.NL - lc) v = p.VinfL[ab];                                       // :713-721
        else if (ji < -lc) v = p.VL[ab];
        else if (ji > lc + p.NR) v = p.VinfR[ab];
        else if (ji > lc) v = p.VR[ab];
    } else if (k == i + 1) {
        hop_left_type(ji, v);
        hop_right_type(jk, v);
    } else if (k == i - 1) {
        hop_left_type(jk, v);
        hop_right_type(ji, v);
    }
    p.H[idx] = v;
}

// Landauer current from an already computed transmission function (the
// reference's fcdmft/__init__.py:landauer:238-270 forms jE = Tr[...] (nL - nR)/(2 pi) per energy and
// integrates it with np.trapz, :426).  One CTA per bias point.
__global__ void __launch_bounds__(256) landauer_kernel(const double* __restrict__ T, const double* __restrict__ E, int nE,
                                                       const double* __restrict__ muL, const double* __restrict__ muR,
                                                       double kBT, double* __restrict__ jE, double* __restrict__ I) {
    __shared__ double s_red[8];
    const int b = blockIdx.x;
    const double mL = muL[b], mR = muR[b];
    auto occ = [&](double e, double mu) -> double {
        if (kBT == 0.0) return e <= mu ? 1.0 : 0.0;           // step function (:241-245)
        return 1.0 / (exp((e - mu) / kBT) + 1.0);              // :247-248
    };
    auto dens = [&](int i) -> double { return T[i] * (occ(E[i], mL) - occ(E[i], mR)) / (2.0 * 3.141592653589793); };
    double acc = 0.0;
    for (int i = threadIdx.x; i < nE; i += blockDim.x) {
        const double ji = dens(i);
        if (jE) jE[(long long)b * nE + i] = ji;
        if (i + 1 < nE) acc += 0.5 * (E[i + 1] - E[i]) * (ji + dens(i + 1));   // trapezoid rule
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += s_red[w];
        I[b] = tot;
    }
}

struct DevMem {
    char* p = nullptr;
    ~DevMem() {
        if (p) cudaFree(p);
    }
};

int green_host(int which, const tx_cplx* h, const tx_cplx* tnn, int n, int M, const tx_cplx* E, int nE,
               const tx_cplx* sigL, const tx_cplx* sigR, tx_cplx* out) {
    if (!h || !tnn || !E || !sigL || !sigR || !out) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || n > TX_MAX_N) return fail(TX_ERR_UNSUPPORTED, "block size n=%d outside 1..%d", n, TX_MAX_N);
    if (M < 2 || nE < 1) return fail(TX_ERR_ARG, "bad sizes M=%d nE=%d", M, nE);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t nn = (size_t)n * n, bc = sizeof(cd);
    const size_t b_h = (size_t)M * nn * bc, b_t = (size_t)(M - 1) * nn * bc, b_E = ((size_t)nE * bc + 255) & ~(size_t)255;
    const size_t b_s = (size_t)nE * nn * bc;
    const size_t b_out = (size_t)nE * (which == 0 ? M : (size_t)M * M) * nn * bc;
    const size_t b_work = which == 0 ? 0 : (size_t)nE * 2 * M * nn * bc;
    size_t smem = 2 * nn * bc;
    const size_t b_ax = smem > 200 * 1024 ? (size_t)nE * smem : 0;
    if (b_ax) smem = 0;
    DevMem d;
    TX_CUDA(cudaMalloc(&d.p, b_h + b_t + b_E + 2 * b_s + b_out + b_work + b_ax + 256));
    char* q = d.p;
    GreenParams p{};
    p.h = (const cd*)q;      TX_CUDA(cudaMemcpy(q, h, b_h, cudaMemcpyHostToDevice));      q += b_h;
    p.tnn = (const cd*)q;    TX_CUDA(cudaMemcpy(q, tnn, b_t, cudaMemcpyHostToDevice));    q += b_t;
    p.E = (const cd*)q;      TX_CUDA(cudaMemcpy(q, E, (size_t)nE * bc, cudaMemcpyHostToDevice)); q += b_E;
    p.sigL = (const cd*)q;   TX_CUDA(cudaMemcpy(q, sigL, b_s, cudaMemcpyHostToDevice));   q += b_s;
    p.sigR = (const cd*)q;   TX_CUDA(cudaMemcpy(q, sigR, b_s, cudaMemcpyHostToDevice));   q += b_s;
    p.out = (cd*)q;          q += b_out;
    p.work = (cd*)q;         q += b_work;
    p.ax = b_ax ? (cd*)q : nullptr;  q += b_ax;
    p.singular = (int*)q;    TX_CUDA(cudaMemset(q, 0, 4));
    p.n = n;
    p.M = M;
    p.nE = nE;
    if (which == 0) {
        if (smem > 48 * 1024)
            TX_CUDA(cudaFuncSetAttribute(green_column0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        green_column0_kernel<<<nE, GR_THREADS, smem>>>(p);
    } else {
        if (smem > 48 * 1024)
            TX_CUDA(cudaFuncSetAttribute(green_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        green_full_kernel<<<nE, GR_THREADS, smem>>>(p);
    }
    TX_CUDA(cudaGetLastError());
    TX_CUDA(cudaDeviceSynchronize());
    int flag = 0;
    TX_CUDA(cudaMemcpy(out, p.out, b_out, cudaMemcpyDeviceToHost));
    TX_CUDA(cudaMemcpy(&flag, p.singular, 4, cudaMemcpyDeviceToHost));
    if (flag) return fail(TX_ERR_SINGULAR, "singular block in the Green's function recursion");
    return TX_OK;
}

}  // namespace

extern "C" {

int tx_green_column0(const tx_cplx* h, const tx_cplx* tnn, int n, int M, const tx_cplx* E, int nE,
                     const tx_cplx* sigL, const tx_cplx* sigR, tx_cplx* G) {
    return green_host(0, h, tnn, n, M, E, nE, sigL, sigR, G);
}

int tx_green_full(const tx_cplx* h, const tx_cplx* tnn, int n, int M, const tx_cplx* E, int nE,
                  const tx_cplx* sigL, const tx_cplx* sigR, tx_cplx* G) {
    return green_host(1, h, tnn, n, M, E, nE, sigL, sigR, G);
}

int tx_hmat(const tx_cplx* h, const tx_cplx* tnn, const tx_cplx* tnnn, int n, int M, tx_cplx* H) {
    if (!h || !tnn || !H || (M > 2 && !tnnn)) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || M < 2) return fail(TX_ERR_ARG, "bad sizes n=%d M=%d", n, M);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t nn = (size_t)n * n, bc = sizeof(cd);
    const size_t b_h = (size_t)M * nn * bc, b_t = (size_t)(M - 1) * nn * bc, b_tt = (size_t)(M > 2 ? M - 2 : 1) * nn * bc;
    const size_t dim = (size_t)n * M, b_H = dim * dim * bc;
    DevMem d;
    TX_CUDA(cudaMalloc(&d.p, b_h + b_t + b_tt + b_H));
    char* q = d.p;
    const cd* dh = (const cd*)q;   TX_CUDA(cudaMemcpy(q, h, b_h, cudaMemcpyHostToDevice));     q += b_h;
    const cd* dt = (const cd*)q;   TX_CUDA(cudaMemcpy(q, tnn, b_t, cudaMemcpyHostToDevice));   q += b_t;
    const cd* dtt = (const cd*)q;
    if (M > 2) TX_CUDA(cudaMemcpy(q, tnnn, (size_t)(M - 2) * nn * bc, cudaMemcpyHostToDevice));
    q += b_tt;
    cd* dH = (cd*)q;
    const long long total = (long long)dim * dim;
    hmat_kernel<<<(unsigned)((total + 255) / 256), 256>>>(dh, dt, dtt, dH, n, M);
    TX_CUDA(cudaGetLastError());
    TX_CUDA(cudaMemcpy(H, dH, b_H, cudaMemcpyDeviceToHost));
    return TX_OK;
}

int tx_landauer_current_dev(const double* T, const double* E, int nE, const double* muL, const double* muR, int n_bias,
                            double kBT, double* jE, double* I, void* stream) {
    if (!T || !E || !muL || !muR || !I) return fail(TX_ERR_ARG, "null pointer argument");
    if (nE < 2 || n_bias < 1 || kBT < 0.0) return fail(TX_ERR_ARG, "bad sizes nE=%d n_bias=%d kBT=%g", nE, n_bias, kBT);
    landauer_kernel<<<n_bias, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(T, E, nE, muL, muR, kBT, jE, I);
    TX_CUDA(cudaGetLastError());
    count_launch();
    return TX_OK;
}

int tx_landauer_current(const double* T, const double* E, int nE, const double* muL, const double* muR, int n_bias,
                        double kBT, double* jE, double* I) {
    if (!T || !E || !muL || !muR || !I) return fail(TX_ERR_ARG, "null pointer argument");
    if (nE < 2 || n_bias < 1 || kBT < 0.0) return fail(TX_ERR_ARG, "bad sizes nE=%d n_bias=%d kBT=%g", nE, n_bias, kBT);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t bE = (size_t)nE * 8, bB = (size_t)n_bias * 8, bJ = jE ? (size_t)n_bias * nE * 8 : 0;
    DevMem d;
    TX_CUDA(cudaMalloc(&d.p, 2 * bE + 3 * bB + bJ + 64));
    char* q = d.p;
    double* dT = (double*)q;  TX_CUDA(cudaMemcpy(q, T, bE, cudaMemcpyHostToDevice));    q += bE;
    double* dE = (double*)q;  TX_CUDA(cudaMemcpy(q, E, bE, cudaMemcpyHostToDevice));    q += bE;
    double* dL = (double*)q;  TX_CUDA(cudaMemcpy(q, muL, bB, cudaMemcpyHostToDevice));  q += bB;
    double* dR = (double*)q;  TX_CUDA(cudaMemcpy(q, muR, bB, cudaMemcpyHostToDevice));  q += bB;
    double* dI = (double*)q;  q += bB;
    double* dJ = jE ? (double*)q : nullptr;
    int rc = tx_landauer_current_dev(dT, dE, nE, dL, dR, n_bias, kBT, dJ, dI, nullptr);
    if (rc) return rc;
    TX_CUDA(cudaMemcpy(I, dI, bB, cudaMemcpyDeviceToHost));
    if (jE) TX_CUDA(cudaMemcpy(jE, dJ, bJ, cudaMemcpyDeviceToHost));
    return TX_OK;
}

int tx_hsysmat(const tx_cplx* tinf, const tx_cplx* tL, const tx_cplx* tR, const tx_cplx* VinfL, const tx_cplx* VinfR,
               const tx_cplx* VL, const tx_cplx* VR, int Ninf, int NL, int NR, const tx_cplx* HC, int NC, int n,
               tx_cplx* H) {
    if (!tinf || !tL || !tR || !VinfL || !VinfR || !VL || !VR || !HC || !H) return fail(TX_ERR_ARG, "null pointer argument");
    if (Ninf <= 0 || NL <= 0 || NR <= 0 || NC < 1 || (NC % 2) != 1 || n < 1) return fail(TX_ERR_ARG, "bad sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t nn = (size_t)n * n, bc = sizeof(cd);
    const size_t S = (size_t)2 * Ninf + NL + NR + NC;
    const size_t b_small = nn * bc, b_HC = (size_t)NC * NC * nn * bc, b_H = S * S * nn * bc;
    DevMem d;
    TX_CUDA(cudaMalloc(&d.p, 7 * b_small + b_HC + b_H));
    char* q = d.p;
    const tx_cplx* srcs[7] = {tinf, tL, tR, VinfL, VinfR, VL, VR};
    const cd* dp[7];
    for (int i = 0; i < 7; ++i) {
        TX_CUDA(cudaMemcpy(q, srcs[i], b_small, cudaMemcpyHostToDevice));
        dp[i] = (const cd*)q;
        q += b_small;
    }
    TX_CUDA(cudaMemcpy(q, HC, b_HC, cudaMemcpyHostToDevice));
    HsysParams p{dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], (const cd*)q, (cd*)(q + b_HC), Ninf, NL, NR, NC, n};
    const long long total = (long long)(S * S * nn);
    hsysmat_kernel<<<(unsigned)((total + 255) / 256), 256>>>(p);
    TX_CUDA(cudaGetLastError());
    TX_CUDA(cudaMemcpy(H, p.H, b_H, cudaMemcpyDeviceToHost));
    return TX_OK;
}

}  // extern "C"
