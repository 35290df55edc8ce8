// K1: batched lead surface Green's functions and self-energies, one CTA per energy.
//
// Reference functions replaced (cpbunker/transport, transport/wfm/__init__.py):
//   g_closed   :306-340   closed form per channel of a monatomic lead
//   g_RiceMele :352-414   closed form per spin channel of a diatomic (Rice-Mele) lead
//   g_ith      :532-551   one step of the fixed point  g <- (z - h00 - t g t^dag)^-1
//   g_iter     :479-530   that iteration with the surface-DOS stopping rule
//   SelfEnergies :270-304 Sigma_L = t0^dag g_L t0,  Sigma_R = tN g_R tN^dag
// plus Sancho-Rubio decimation (quadratically convergent) of the same fixed point for
// multi-channel leads, which the reference cannot do (g_iter refuses n > 2, :502).
#include <cuda_runtime.h>

#include <cstdlib>

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "block_ops.cuh"
#include "tx_host.h"

using namespace tx;

namespace {

struct GfParams {
    const cd* h00;      // [n][n]
    const cd* h01;      // [n][n]  upper hopping block of the lead
    const cd* E;        // [nE]
    cd* g;              // [nE][n][n] or null
    cd* sigma;          // [nE][n][n] or null:  side=-1: h01^dag g h01 ; side=+1: h01 g h01^dag
    cd* g2;             // SANCHO only: the OTHER semi-infinite continuation of the same lead (side -p.side), or null
    cd* sigma2;         //   ... and its self-energy, or null
    const cd* g_prev;   // g_ith only
    int* n_iter;        // [nE] or null
    double* conv;       // [nE] or null
    int* converged;     // [nE] or null
    const double* evals;  // SEPARABLE: eigenvalues of h00 [n]
    const cd* evecs;      // SEPARABLE: eigenvectors of h00 [n][n] (columns)
    cd* scratch;        // global workspace (n > 32), per energy `scratch_per` elements
    long long scratch_per;
    int* singular;
    int n, nE, side, mode;
    double imE, conv_tol;
    int conv_rep, min_iter, max_iter, ith;
    int use_tma;        // n = 64: stage the products' B operand in shared memory by TMA (tx_set_option(5, .))
    int herm_shortcut;  // SANCHO: two products per iteration when h01 is Hermitian (tx_set_option(6, .))
};

constexpr int GF_THREADS = 256;
int g_use_tma = 1;   // tx_set_option(5, value)
int g_herm_shortcut = 1;   // tx_set_option(6, value)
constexpr int GF_MAX_MATS = 9;

// wfm.g_RiceMele for spin channel s (wfm/__init__.py:379-403)
__device__ cd ricemele_channel(const cd* diag, const cd* off, int n, int s, cd E) {
    const int ns = n / 2;
    const cd dA = diag[s * n + s], dB = diag[(ns + s) * n + (ns + s)];
    const cd u = scale(dA - dB, 0.5);
    const cd u0 = scale(dA + dB, 0.5);
    const cd v = diag[s * n + (ns + s)];
    const cd w = off[(ns + s) * n + s];
    const cd x = E - u0;
    const cd Z = (x + u) * (x - u) + (w + v) * (w - v);
    const cd den = scale(w * w, 2.0) * (E - u0 - u);
    const cd pref = cdiv(mk(1.0, 0.0), den);
    const cd inner = scale(w * w, 4.0) * (u + E - u0) * (u - E + u0) + Z * Z;
    cd root = csqrt_principal(inner);
    const cd chk = root * pref;
    if (chk.y > 0.0) root = mk(-root.x, -root.y);          // retarded: Im g < 0  (:401-402)
    return pref * (Z + root);
}

// sigma = h01^dag g h01 (side=-1)  or  h01 g h01^dag (side=+1)
template <int FIXN>
__device__ void emit_sigma(const GfParams& p, int side, const cd* g, cd* tmp, cd* out) {
    const int n = p.n;
    if (side == -1) {
        cta_gemm<FIXN>(tmp, g, OP_N, p.h01, OP_N, n);
        cta_gemm<FIXN>(out, p.h01, OP_H, tmp, OP_N, n);
    } else {
        cta_gemm<FIXN>(tmp, g, OP_N, p.h01, OP_H, n);
        cta_gemm<FIXN>(out, p.h01, OP_N, tmp, OP_N, n);
    }
}

// Hermitian eigendecomposition of one n x n block (n <= 64) by parallel cyclic Jacobi, one CTA:
// round-robin tournament of n/2 disjoint rotations per round, rotations J = [[c, s], [-conj s, c]] chosen to
// annihilate A[p][q], applied as A <- J^dag A J (columns, then rows) and V <- V J.  A and V live in shared memory.
// Serves the separable closed form below (once per lead, not per energy).  status: bit 0 = h01 is not a scalar
// multiple of the identity, bit 1 = h00 is not Hermitian, bit 2 = no convergence.
__global__ void __launch_bounds__(GF_THREADS) hermitian_eig_kernel(const cd* __restrict__ h00, const cd* __restrict__ h01,
                                                                   int n, double* __restrict__ evals,
                                                                   cd* __restrict__ evecs, int* __restrict__ status) {
    extern __shared__ double2 eig_smem[];
    __shared__ double s_c[32];
    __shared__ cd s_s[32];
    __shared__ int s_p[32], s_q[32];
    __shared__ double s_red[GF_THREADS / 32];
    __shared__ int s_bad;
    cd* A = reinterpret_cast<cd*>(eig_smem);
    cd* V = A + n * n;
    const int tid = threadIdx.x, nn = n * n;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    double scale2 = 0.0;
    for (int idx = tid; idx < nn; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        const cd a = h00[idx], at = h00[c * n + r], t = h01[idx], t0 = h01[0];
        A[idx] = a;
        V[idx] = mk(r == c ? 1.0 : 0.0, 0.0);
        scale2 = fmax(scale2, norm2(a));
        const double tol = 1e-24 * (norm2(a) + norm2(t0) + 1e-300);
        if (norm2(mk(a.x - at.x, a.y + at.y)) > 1e-24 * (norm2(a) + 1e-300)) atomicOr(&s_bad, 2);
        const cd want = (r == c) ? t0 : mk(0.0, 0.0);
        if (norm2(t - want) > tol) atomicOr(&s_bad, 1);
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) scale2 = fmax(scale2, __shfl_xor_sync(0xffffffffu, scale2, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = scale2;
    __syncthreads();
    scale2 = 1e-300;
    for (int w = 0; w < GF_THREADS / 32; ++w) scale2 = fmax(scale2, s_red[w]);
    __syncthreads();
    const int m = n + (n & 1);                    // tournament size (a bye for odd n)
    const int npairs = m / 2;
    bool converged = false;
    for (int sweep = 0; sweep < 30 && !converged; ++sweep) {
        for (int round = 0; round < m - 1; ++round) {
            if (tid < npairs) {
                int pp, qq;
                if (tid == 0) {
                    pp = m - 1;
                    qq = round % (m - 1);
                } else {
                    pp = (round + tid) % (m - 1);
                    qq = (round - tid + 2 * (m - 1)) % (m - 1);
                }
                double c = 1.0;
                cd sv = mk(0.0, 0.0);
                if (pp < n && qq < n) {
                    const double al = A[pp * n + pp].x, ga = A[qq * n + qq].x;
                    const cd be = A[pp * n + qq];
                    const double ab = sqrt(norm2(be));
                    if (ab > 1e-300) {
                        const double tau = (al - ga) / (2.0 * ab);
                        const double t = (tau == 0.0) ? -1.0 : -copysign(1.0, tau) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                        c = 1.0 / sqrt(fma(t, t, 1.0));
                        const double f = t * c / ab;
                        sv = mk(f * be.x, f * be.y);
                    }
                } else {
                    pp = qq = -1;
                }
                s_c[tid] = c;
                s_s[tid] = sv;
                s_p[tid] = pp;
                s_q[tid] = qq;
            }
            __syncthreads();
            // columns of A and V:  X[:,p] <- c X[:,p] - conj(s) X[:,q],  X[:,q] <- s X[:,p] + c X[:,q]
            for (int w = tid; w < npairs * n; w += blockDim.x) {
                const int k = w / n, r = w - k * n;
                const int pp = s_p[k], qq = s_q[k];
                if (pp < 0) continue;
                const double c = s_c[k];
                const cd sv = s_s[k], sc = conj(sv);
                cd xp = A[r * n + pp], xq = A[r * n + qq];
                A[r * n + pp] = scale(xp, c) - sc * xq;
                A[r * n + qq] = sv * xp + scale(xq, c);
                xp = V[r * n + pp];
                xq = V[r * n + qq];
                V[r * n + pp] = scale(xp, c) - sc * xq;
                V[r * n + qq] = sv * xp + scale(xq, c);
            }
            __syncthreads();
            // rows of A:  A[p,:] <- c A[p,:] - s A[q,:],  A[q,:] <- conj(s) A[p,:] + c A[q,:]
            for (int w = tid; w < npairs * n; w += blockDim.x) {
                const int k = w / n, cc = w - k * n;
                const int pp = s_p[k], qq = s_q[k];
                if (pp < 0) continue;
                const double c = s_c[k];
                const cd sv = s_s[k], sc = conj(sv);
                const cd xp = A[pp * n + cc], xq = A[qq * n + cc];
                A[pp * n + cc] = scale(xp, c) - sv * xq;
                A[qq * n + cc] = sc * xp + scale(xq, c);
            }
            __syncthreads();
        }
        double off = 0.0;
        for (int idx = tid; idx < nn; idx += blockDim.x) {
            const int r = idx / n, c = idx - r * n;
            if (r != c) off = fmax(off, norm2(A[idx]));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = off;
        __syncthreads();
        off = 0.0;
        for (int w = 0; w < GF_THREADS / 32; ++w) off = fmax(off, s_red[w]);
        __syncthreads();
        converged = off <= 1e-30 * scale2;        // |off-diagonal| <= 1e-15 |A|: CTA-uniform
    }
    for (int c = tid; c < n; c += blockDim.x) evals[c] = A[c * n + c].x;
    for (int idx = tid; idx < nn; idx += blockDim.x) evecs[idx] = V[idx];
    if (tid == 0) *status = s_bad | (converged ? 0 : 4);
}

// TX_GF_SEPARABLE, one CTA per energy: scalar hopping h01 = t I, so the surface GF is a function of h00,
//   g = U diag(g_closed(z, eps_m, t)) U^dag   (the reference's g_closed per transverse mode, wfm/__init__.py:337-340),
// Sigma = |t|^2 g on both sides.  U, eps from hermitian_eig_kernel; one n x n block (U diag(g)) in shared memory.
__global__ void __launch_bounds__(GF_THREADS) separable_gf_kernel(const GfParams p) {
    extern __shared__ double2 gf_smem[];
    const int n = p.n;
    const int nn = n * n;
    const int e = blockIdx.x;
    const cd E = p.E[e];
    const cd z = mk(E.x, E.y + p.imE);
    const cd t0 = p.h01[0];
    cd* B = reinterpret_cast<cd*>(gf_smem);
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
        const int c = idx % n;
        const cd gm = g_closed_channel(z, mk(p.evals[c], 0.0), t0);
        B[idx] = p.evecs[idx] * gm;
    }
    __syncthreads();
    cd* out = p.g ? p.g + (long long)e * nn : p.sigma + (long long)e * nn;
    if (threadIdx.x == 0) {
        if (p.n_iter) p.n_iter[e] = 0;
        if (p.conv) p.conv[e] = 0.0;
        if (p.converged) p.converged[e] = 1;
    }
    cta_gemm(out, B, OP_N, p.evecs, OP_H, n);
    const double kap = norm2(t0);
    if (p.g && p.sigma) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) p.sigma[(long long)e * nn + idx] = scale(out[idx], kap);
    } else if (p.sigma) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) out[idx] = scale(out[idx], kap);
    }
}

// The decimation for a Hermitian hopping block (alpha = beta throughout): out of line, so that its registers do not weigh on
// the general loop of surface_gf_body.  On return `alpha` holds the last hopping block (unused), eps / epss the bulk and
// surface blocks.
template <int NMAX, int FIXN>
__device__ __noinline__ void sancho_hermitian_loop(const GfParams& p, const cd z, cd* stage, TmaStage* ts, cd* s_colk, int* s_ipiv,
                                                   cd* alpha, cd* tmp, cd* eps, cd* epss, cd* G, cd* T2, const double tol2,
                                                   int& it, int& ok, double& resid) {
    const int n = p.n, nn = n * n;
    {
            bool fused_h = false;
            if constexpr (FIXN == 64) fused_h = ts != nullptr;
            for (it = 1; it <= p.max_iter; ++it) {
                if constexpr (FIXN == 64) {
                    if (fused_h) {
                        cta_inverse_blocked(G, stage, n, p.singular, eps, z);               // G = (z - eps)^-1
                        cta_gemm<FIXN>(T2, G, OP_N, alpha, OP_N, n, ts);                    // G alpha
                        double tmax = 0.0;
                        GemmEpi epi;
                        epi.C2 = epss;                                                      // epss += alpha G alpha
                        epi.C3 = eps;                                                       // eps  += 2 alpha G alpha
                        epi.c3s = 2.0;
                        epi.tmax = &tmax;
                        cta_gemm_dmma_nn_tma<64>(tmp, alpha, T2, *ts, 1.0, false, 1.0, epi); // new alpha
                        resid = cta_reduce_max(tmax);
                    }
                }
                if (!fused_h) {
                    cta_z_minus(G, z, eps, n);
                    cta_inverse_staged<NMAX>(G, stage, n, s_colk, s_ipiv, p.singular);
                    cta_gemm<FIXN>(T2, G, OP_N, alpha, OP_N, n, ts);                        // G alpha
                    cta_gemm<FIXN>(tmp, alpha, OP_N, T2, OP_N, n, ts);                      // alpha G alpha
                    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
                        const cd v = tmp[idx];
                        const cd a = epss[idx], b = eps[idx];
                        epss[idx] = mk(a.x + v.x, a.y + v.y);
                        eps[idx] = mk(fma(2.0, v.x, b.x), fma(2.0, v.y, b.y));
                    }
                    __syncthreads();
                    resid = cta_maxnorm2(tmp, nn);
                }
                {
                    cd* sw = alpha;
                    alpha = tmp;
                    tmp = sw;
                }
                if (resid < tol2) {
                    ok = 1;
                    break;
                }
            }
    }
}

template <int NMAX, int FIXN>
__device__ __forceinline__ void surface_gf_body(const GfParams& p, const int e) {
    extern __shared__ double2 gf_smem[];
    __shared__ int s_ipiv[NMAX];
    __shared__ cd s_colk[NMAX];
    __shared__ int s_flag;
    const int n = p.n;
    const int nn = n * n;
    const cd E = p.E[e];
    // the global workspace (blocks too large for shared memory) is indexed by CTA, not by energy: only the
    // resident CTAs need one, and a few hundred of them stay L2-resident
    cd* base = (p.scratch != nullptr) ? p.scratch + (long long)blockIdx.x * p.scratch_per : reinterpret_cast<cd*>(gf_smem);
    // inverse staging when blocks are global: shared memory, or (n > 104) one more block of the workspace
    cd* stage = (p.scratch != nullptr) ? reinterpret_cast<cd*>(gf_smem) : nullptr;
    if constexpr (NMAX > TX_STAGE_SMEM_MAX_N)
        if (p.scratch != nullptr && !inverse_stage_in_smem(n)) stage = base + p.scratch_per - (long long)n * (n + 1);
    // n = 64 with the blocks in the global workspace: the products stage their B operand through the same block by TMA
    __shared__ unsigned long long s_tma_bar;
    TmaStage tma_stage;
    TmaStage* ts = nullptr;
    if (stage != nullptr && p.use_tma && (FIXN == 64 || tma_stage_pays(n))) {
        tma_stage_init(tma_stage, stage, &s_tma_bar);
        ts = &tma_stage;
    }
    cd* m0 = base;               // g / result
    cd* m1 = base + nn;
    cd* m2 = base + 2 * nn;

    if (p.mode == TX_GF_CLOSED || p.mode == TX_GF_RICEMELE) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) m0[idx] = mk(0.0, 0.0);
        __syncthreads();
        if (p.mode == TX_GF_CLOSED) {
            for (int c = threadIdx.x; c < n; c += blockDim.x)
                m0[c * n + c] = g_closed_channel(E, p.h00[c * n + c], p.h01[c * n + c]);
        } else {
            const int ns = n / 2;
            for (int s = threadIdx.x; s < ns; s += blockDim.x) {
                const cd gs = ricemele_channel(p.h00, p.h01, n, s, E);
                if (p.side == -1) m0[(ns + s) * n + (ns + s)] = gs;   // <B|g|B>  (:407-410)
                else m0[s * n + s] = gs;                               // <A|g|A>  (:411-413)
            }
        }
        if (threadIdx.x == 0) {
            if (p.n_iter) p.n_iter[e] = 0;
            if (p.conv) p.conv[e] = 0.0;
            if (p.converged) p.converged[e] = 1;
        }
        __syncthreads();
    } else if (p.mode == TX_GF_FIXEDPOINT) {
        // g_iter (:504-530) / g_ith (:541-551)
        const cd z = mk(E.x, fabs(p.imE));
        const int dpos = (p.side == 1) ? 0 : nn - 1;           // :516-517
        double dos_prev = 0.0, check = 0.0;
        int run = 0, it_done = 0, ok = 0;
        const int first = (p.ith >= 0) ? p.ith : 0;
        const int last = (p.ith >= 0) ? p.ith + 1 : p.max_iter;
        if (p.ith > 0) cta_copy(m0, p.g_prev + (long long)e * nn, nn);
        for (int ith = first; ith < last; ++ith) {
            cta_z_minus(m1, z, p.h00, n);
            if (ith > 0) {
                if (p.side == 1) {                              // t g t^dag   (:548)
                    cta_gemm<FIXN>(m2, m0, OP_N, p.h01, OP_H, n);
                    cta_gemm_acc<FIXN>(m1, p.h01, OP_N, m2, OP_N, n, -1.0);
                } else {                                        // t^dag g t   (:550)
                    cta_gemm<FIXN>(m2, m0, OP_N, p.h01, OP_N, n);
                    cta_gemm_acc<FIXN>(m1, p.h01, OP_H, m2, OP_N, n, -1.0);
                }
            }
            cta_inverse_staged<NMAX>(m1, stage, n, s_colk, s_ipiv, p.singular);
            cta_copy(m0, m1, nn);
            it_done = ith;
            if (p.ith < 0) {
                const double dos = (-1.0 / 3.141592653589793) * m0[dpos].y;
                if (ith > p.min_iter) {                         // :520-523
                    check = fabs((dos - dos_prev) / dos);
                    run = (check < p.conv_tol) ? run + 1 : 0;
                    if (run >= p.conv_rep) {
                        ok = 1;
                        break;
                    }
                }
                dos_prev = dos;
            }
        }
        if (threadIdx.x == 0 && p.ith < 0) {
            if (p.n_iter) p.n_iter[e] = it_done;
            if (p.conv) p.conv[e] = check;
            if (p.converged) p.converged[e] = ok;
        }
    } else {   // TX_GF_SANCHO
        // decimation: eps_s -> surface block, alpha/beta -> renormalised hoppings, z = E + i*imE
        const cd z = mk(E.x, E.y + p.imE);
        cd* alpha = base + nn;
        cd* beta = base + 2 * nn;
        cd* eps = base + 3 * nn;
        cd* epss = base + 4 * nn;
        cd* G = base + 5 * nn;
        cd* T1 = base + 6 * nn;
        cd* T2 = base + 7 * nn;
        cd* tmp = base + 8 * nn;
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const int r = idx / n, c = idx - r * n;
            const cd t = p.h01[idx];
            const cd th = conj(p.h01[c * n + r]);
            alpha[idx] = (p.side == 1) ? t : th;
            beta[idx] = (p.side == 1) ? th : t;
            eps[idx] = p.h00[idx];
            epss[idx] = p.h00[idx];
        }
        // Hermitian hopping block (h01 = h01^dag: square-lattice ribbons, any lead whose cells couple symmetrically): then
        // alpha_0 = beta_0, and by induction alpha_i = beta_i = alpha G alpha, alpha G beta = beta G alpha -- two products per
        // iteration instead of six, and both surfaces of the lead coincide.  Detected per CTA (n^2 comparisons).
        int asym = 0;
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const cd a = alpha[idx], b = beta[idx];
            asym |= (a.x != b.x || a.y != b.y) ? 1 : 0;
        }
        const bool herm = !__syncthreads_or(asym | (p.herm_shortcut ? 0 : 1));
        int it = 0, ok = 0;
        double resid = 0.0;
        const double tol2 = p.conv_tol * p.conv_tol;
        if (herm) {
            sancho_hermitian_loop<NMAX, FIXN>(p, z, stage, ts, s_colk, s_ipiv, alpha, tmp, eps, epss, G, T2, tol2, it, ok, resid);
        } else {
        bool fused = false;
        if constexpr (FIXN == 64) fused = ts != nullptr;
        if constexpr (FIXN == 64) if (fused) {
            // n = 64 with TMA-staged products: z - eps is formed while the inverse stages its block, the product alpha G beta
            // is accumulated into both on-site blocks by its own epilogue, and the epilogues of the two hopping products
            // carry the convergence maximum -- no elementwise pass over the blocks is left in the iteration
            for (it = 1; it <= p.max_iter; ++it) {
                cta_inverse_blocked(G, stage, n, p.singular, eps, z);                   // G = (z - eps)^-1
                cta_gemm<FIXN>(T1, G, OP_N, beta, OP_N, n, ts);                               // G beta
                cta_gemm<FIXN>(T2, G, OP_N, alpha, OP_N, n, ts);                              // G alpha
                // the four remaining products in an order that lets two of them reuse the staged operand
                double tmax = 0.0;
                GemmEpi both, conv, conv_reuse, acc_reuse;
                both.C2 = eps;
                conv.tmax = &tmax;
                conv_reuse.tmax = &tmax;
                conv_reuse.reuse_staged = true;
                acc_reuse.reuse_staged = true;
                cta_gemm_dmma_nn_tma<64>(epss, alpha, T1, *ts, 1.0, true, 1.0, both);        // epss, eps += alpha G beta   (stages T1)
                cta_gemm_dmma_nn_tma<64>(G, beta, T1, *ts, 1.0, false, 1.0, conv_reuse);     // beta G beta -> new beta (G is free now)
                cta_gemm_dmma_nn_tma<64>(tmp, alpha, T2, *ts, 1.0, false, 1.0, conv);        // alpha G alpha -> new alpha (stages T2)
                cta_gemm_dmma_nn_tma<64>(eps, beta, T2, *ts, 1.0, true, 1.0, acc_reuse);     // eps += beta G alpha
                {
                    cd* sw = alpha;
                    alpha = tmp;
                    tmp = sw;
                    sw = beta;
                    beta = G;
                    G = sw;
                }
                resid = cta_reduce_max(tmax);
                if (resid < tol2) {
                    ok = 1;
                    break;
                }
            }
        }
        if (!fused)
        for (it = 1; it <= p.max_iter; ++it) {
            cta_z_minus(G, z, eps, n);
            cta_inverse_staged<NMAX>(G, stage, n, s_colk, s_ipiv, p.singular);
            cta_gemm<FIXN>(T1, G, OP_N, beta, OP_N, n, ts);           // G beta
            cta_gemm<FIXN>(T2, G, OP_N, alpha, OP_N, n, ts);          // G alpha
            cta_gemm<FIXN>(tmp, alpha, OP_N, T1, OP_N, n, ts);        // alpha G beta
            {
                const int bd = blockDim.x;
                int idx = threadIdx.x;
                for (; idx + 3 * bd < nn; idx += 4 * bd) {      // four independent loads per array before the stores
                    cd a[4], b[4], c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        a[u] = epss[idx + u * bd];
                        b[u] = eps[idx + u * bd];
                        c[u] = tmp[idx + u * bd];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        epss[idx + u * bd] = a[u] + c[u];
                        eps[idx + u * bd] = b[u] + c[u];
                    }
                }
                for (; idx < nn; idx += bd) {
                    epss[idx] = epss[idx] + tmp[idx];
                    eps[idx] = eps[idx] + tmp[idx];
                }
            }
            __syncthreads();
            cta_gemm_acc<FIXN>(eps, beta, OP_N, T2, OP_N, n, 1.0, 1.0, ts);    // + beta G alpha
            cta_gemm<FIXN>(tmp, alpha, OP_N, T2, OP_N, n, ts);        // alpha G alpha
            cta_gemm<FIXN>(G, beta, OP_N, T1, OP_N, n, ts);           // beta G beta (G is free now)
            {                                                    // new alpha = tmp, new beta = G: swap roles, no copies
                cd* sw = alpha;
                alpha = tmp;
                tmp = sw;
                sw = beta;
                beta = G;
                G = sw;
            }
            resid = cta_maxnorm2(alpha, nn, beta);
            if (resid < tol2) {
                ok = 1;
                break;
            }
        }
        }   // general (non-Hermitian) hopping
        if (p.g2 || p.sigma2) {
            // The decimation carries both surfaces of the lead: eps = h00 + sum (alpha G beta + beta G alpha) is the bulk
            // block, epss = h00 + sum alpha G beta this side's surface block, so the surface block of the OTHER
            // semi-infinite continuation (roles of alpha and beta exchanged) is h00 + sum beta G alpha = eps + h00 - epss:
            // when both leads of a junction are the same material, one decimation serves both.
            for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
                const cd a = eps[idx], b = p.h00[idx], c = epss[idx];
                T1[idx] = mk(a.x + b.x - c.x, a.y + b.y - c.y);
            }
            __syncthreads();
            cta_z_minus(m0, z, T1, n);
            cta_inverse_staged<NMAX>(m0, stage, n, s_colk, s_ipiv, p.singular);
            if (p.g2) cta_copy(p.g2 + (long long)e * nn, m0, nn);
            if (p.sigma2) emit_sigma<FIXN>(p, -p.side, m0, T2, p.sigma2 + (long long)e * nn);
        }
        cta_z_minus(m0, z, epss, n);
        cta_inverse_staged<NMAX>(m0, stage, n, s_colk, s_ipiv, p.singular);
        if (threadIdx.x == 0) {
            if (p.n_iter) p.n_iter[e] = it;
            if (p.conv) p.conv[e] = sqrt(resid);
            if (p.converged) p.converged[e] = ok;
        }
    }
    (void)s_flag;
    if (p.g) cta_copy(p.g + (long long)e * nn, m0, nn);
    if (p.sigma) emit_sigma<FIXN>(p, p.side, m0, m1, p.sigma + (long long)e * nn);
}

// Unit-test hook for the CTA-cooperative inverse of block_ops.cuh (staged / blocked path): one CTA per block.
// `gstage`: per-CTA staging blocks in global memory for n beyond the shared-memory staging (n > 104), else null.
template <int THREADS, int NMAX>
__global__ void __launch_bounds__(THREADS) invert_staged_kernel(const cd* __restrict__ in, cd* __restrict__ out, int n,
                                                                int* singular, int unblocked, cd* gstage) {
    extern __shared__ double2 gf_smem[];
    __shared__ int s_ipiv[NMAX];
    __shared__ cd s_colk[NMAX];
    const long long off = (long long)blockIdx.x * n * n;
    cta_copy(out + off, in + off, n * n);
    cd* S = gstage ? gstage + (long long)blockIdx.x * n * (n + 1) : reinterpret_cast<cd*>(gf_smem);
    if (unblocked) {                                    // 1: the column-by-column Gauss-Jordan, for A/B measurements;
        cta_copy(S, out + off, n * n);                  // 2: the block resident in shared memory (in-place paths, n <= 48)
        if (unblocked == 1) cta_inverse(S, n, s_colk, s_ipiv, singular);
        else cta_inverse_staged<NMAX>(S, nullptr, n, s_colk, s_ipiv, singular);
        cta_copy(out + off, S, n * n);
        return;
    }
    cta_inverse_staged<NMAX>(out + off, S, n, s_colk, s_ipiv, singular);
}

// Measurement hook: `reps` products C = A B per CTA on blocks in global memory (what the large-block paths do).
__global__ void __launch_bounds__(GF_THREADS, 2) gemm_bench_kernel(const cd* __restrict__ A, const cd* __restrict__ B, cd* C, int n,
                                                                    int reps, int use_tma) {
    extern __shared__ double2 gf_smem[];
    __shared__ unsigned long long s_bar;
    const long long off = (long long)blockIdx.x * n * n;
    TmaStage tma_stage;
    TmaStage* ts = nullptr;
    if (use_tma) {
        tma_stage_init(tma_stage, reinterpret_cast<cd*>(gf_smem), &s_bar);
        ts = &tma_stage;
    }
    for (int r = 0; r < reps; ++r) {
        if (n == 64) cta_gemm<64>(C + off, A + off, OP_N, B + off, OP_N, n, ts);
        else cta_gemm<0>(C + off, A + off, OP_N, B + off, OP_N, n, ts);
    }
}

// One CTA per energy, persistent over the energy axis (grid = resident CTAs).  256 threads, two CTAs per SM for n <= 64;
// 512 threads, one CTA per SM for 64 < n <= 128 (the blocked inverse wants 4 n threads).
template <int THREADS, int NMAX, int MINB, int FIXN>
__global__ void __launch_bounds__(THREADS, MINB) surface_gf_kernel(const GfParams p) {
    for (int e = blockIdx.x; e < p.nE; e += gridDim.x) {
        surface_gf_body<NMAX, FIXN>(p, e);
        __syncthreads();
    }
}
constexpr int GF_THREADS_BIG = 512;
// instantiation by block edge: 0: n < 64, 1: n = 64 (compile-time products, TMA staging, fused decimation), 2: n > 64
int gf_kind(int n) { return n > 64 ? 2 : (n == 64 ? 1 : 0); }
const void* gf_kernel(int kind) {
    return kind == 2 ? (const void*)surface_gf_kernel<GF_THREADS_BIG, 128, 1, 0>
                     : (kind == 1 ? (const void*)surface_gf_kernel<GF_THREADS, 64, 2, 64> : (const void*)surface_gf_kernel<GF_THREADS, 64, 2, 0>);
}

// grow-only workspace of the decimation for blocks beyond shared memory (per device)
struct GfScratch {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;        // recorded behind the last kernel that used the workspace
    cudaStream_t last = nullptr;       // ... and the stream it ran on
    bool used = false;
};
GfScratch g_gf_scratch[16];
std::mutex g_gf_scratch_mu;            // the workspace is shared by every caller of a device: uses are serialised

int launch_gf(GfParams p, cudaStream_t st, cd** scratch_owner) {
    const int n = p.n;
    if (p.mode == TX_GF_SEPARABLE) {
        // eigendecomposition of h00 once, then one CTA per energy with a single n x n block in shared memory
        if (!p.g && !p.sigma) return fail(TX_ERR_ARG, "TX_GF_SEPARABLE needs g or sigma");
        const size_t nn = (size_t)n * n;
        char* ws = nullptr;
        TX_CUDA(cudaMalloc(&ws, nn * sizeof(cd) + 64 * sizeof(double) + 256));
        *scratch_owner = reinterpret_cast<cd*>(ws);
        cd* evecs = reinterpret_cast<cd*>(ws);
        double* evals = reinterpret_cast<double*>(ws + nn * sizeof(cd));
        int* status = reinterpret_cast<int*>(ws + nn * sizeof(cd) + 64 * sizeof(double));
        const size_t smem_eig = 2 * nn * sizeof(cd);
        if (smem_eig > 48 * 1024)
            TX_CUDA(cudaFuncSetAttribute(hermitian_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_eig));
        hermitian_eig_kernel<<<1, GF_THREADS, smem_eig, st>>>(p.h00, p.h01, n, evals, evecs, status);
        count_launch();
        TX_CUDA(cudaGetLastError());
        int host_status = 0;
        TX_CUDA(cudaMemcpyAsync(&host_status, status, sizeof(int), cudaMemcpyDeviceToHost, st));
        TX_CUDA(cudaStreamSynchronize(st));
        if (host_status & 1) return fail(TX_ERR_ARG, "TX_GF_SEPARABLE needs a hopping block that is a scalar multiple of the identity");
        if (host_status & 2) return fail(TX_ERR_ARG, "TX_GF_SEPARABLE needs a Hermitian on-site block");
        if (host_status & 4) return fail(TX_ERR_NOCONV, "Jacobi eigensolver did not converge");
        p.evals = evals;
        p.evecs = evecs;
        p.scratch = nullptr;
        p.scratch_per = 0;
        const size_t smem = nn * sizeof(cd);
        if (smem > 48 * 1024)
            TX_CUDA(cudaFuncSetAttribute(separable_gf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        separable_gf_kernel<<<p.nE, GF_THREADS, smem, st>>>(p);
        count_launch();
        TX_CUDA(cudaGetLastError());
        return TX_OK;
    }
    const size_t mats = (p.mode == TX_GF_SANCHO) ? GF_MAX_MATS : 3;
    const size_t bytes = mats * (size_t)n * n * sizeof(cd);
    const int kind = gf_kind(n);
    size_t smem = bytes;
    p.scratch = nullptr;
    p.scratch_per = 0;
    size_t scratch_bytes = 0;
    if (bytes > 200 * 1024) {
        smem = inverse_stage_in_smem(n) ? inverse_stage_bytes(n) : 0;          // staging block for the inverse
        if ((n == 64 || tma_stage_pays(n)) && tma_stage_bytes(n) > smem) smem = tma_stage_bytes(n);   // ... and the TMA-staged product operand
        p.scratch_per = (long long)mats * n * n + (inverse_stage_in_smem(n) ? 0 : (long long)n * (n + 1));
        scratch_bytes = (size_t)p.scratch_per * sizeof(cd);
    }
    if (smem > 16 * 1024)   // the kernel also has 11 KB (n <= 64) / 26 KB of static shared memory
        TX_CUDA(cudaFuncSetAttribute(gf_kernel(kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 1;
    TX_CUDA(cudaGetDevice(&dev));
    TX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gf_kernel(kind), kind == 2 ? GF_THREADS_BIG : GF_THREADS, smem));
    const int grid = (int)((long long)p.nE < (long long)sms * per_sm ? p.nE : sms * (per_sm > 0 ? per_sm : 1));
    std::unique_lock<std::mutex> scratch_lock(g_gf_scratch_mu, std::defer_lock);
    GfScratch* used_scratch = nullptr;
    if (p.scratch_per) {
        // one grow-only workspace per device, indexed by CTA: a call on another stream (or from another host thread) waits
        // on the device for the previous user -- no host synchronisation
        scratch_lock.lock();
        GfScratch& sc = g_gf_scratch[dev & 15];
        used_scratch = &sc;
        if (sc.used && sc.last != st && sc.done) TX_CUDA(cudaStreamWaitEvent(st, sc.done, 0));
        const size_t need = (size_t)grid * scratch_bytes;
        if (need > sc.cap) {
            TX_CUDA(cudaDeviceSynchronize());
            if (sc.ptr) cudaFree(sc.ptr);
            sc.ptr = nullptr;
            sc.cap = 0;
            TX_CUDA(cudaMalloc(&sc.ptr, need));
            sc.cap = need;
        }
        p.scratch = reinterpret_cast<cd*>(sc.ptr);
    }
    if (kind == 2) surface_gf_kernel<GF_THREADS_BIG, 128, 1, 0><<<grid, GF_THREADS_BIG, smem, st>>>(p);
    else if (kind == 1) surface_gf_kernel<GF_THREADS, 64, 2, 64><<<grid, GF_THREADS, smem, st>>>(p);
    else surface_gf_kernel<GF_THREADS, 64, 2, 0><<<grid, GF_THREADS, smem, st>>>(p);
    count_launch();
    TX_CUDA(cudaGetLastError());
    if (used_scratch) {
        if (!used_scratch->done) TX_CUDA(cudaEventCreateWithFlags(&used_scratch->done, cudaEventDisableTiming));
        TX_CUDA(cudaEventRecord(used_scratch->done, st));
        used_scratch->last = st;
        used_scratch->used = true;
    }
    return TX_OK;
}

int check_gf_args(const void* h00, const void* h01, int n, const void* E, int nE, int side, int mode) {
    if (!h00 || !h01 || !E) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || n > TX_MAX_N) return fail(TX_ERR_UNSUPPORTED, "surface GF block size n=%d outside 1..%d", n, TX_MAX_N);
    if (mode == TX_GF_SEPARABLE && n > 64) return fail(TX_ERR_UNSUPPORTED, "separable leads: block size n=%d outside 1..64", n);
    if (nE < 1) return fail(TX_ERR_ARG, "nE must be positive");
    if (side != 1 && side != -1) return fail(TX_ERR_ARG, "side must be +1 or -1");
    if (mode < TX_GF_CLOSED || mode > TX_GF_SEPARABLE) return fail(TX_ERR_ARG, "unknown surface GF mode %d", mode);
    if (mode == TX_GF_RICEMELE && (n % 2)) return fail(TX_ERR_ARG, "Rice-Mele leads need an even block size");
    return TX_OK;
}

}  // namespace

namespace tx {
int set_tma_mode(int on) {
    g_use_tma = on != 0;
    return TX_OK;
}
int tma_mode() { return g_use_tma; }
int set_hermitian_shortcut(int on) {
    g_herm_shortcut = on != 0;
    return TX_OK;
}
}  // namespace tx


extern "C" {

int tx_surface_gf_dev(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, int mode,
                      double imE, double conv_tol, int conv_rep, int min_iter, int max_iter,
                      tx_cplx* g, tx_cplx* sigma, int* n_iter, double* conv, int* converged,
                      int* singular_flag, void* stream) {
    int rc = check_gf_args(h00, h01, n, E, nE, side, mode);
    if (rc) return rc;
    if (!singular_flag) return fail(TX_ERR_ARG, "singular_flag is required");
    GfParams p{};
    p.h00 = reinterpret_cast<const cd*>(h00);
    p.h01 = reinterpret_cast<const cd*>(h01);
    p.E = reinterpret_cast<const cd*>(E);
    p.g = reinterpret_cast<cd*>(g);
    p.sigma = reinterpret_cast<cd*>(sigma);
    p.g_prev = nullptr;
    p.n_iter = n_iter;
    p.conv = conv;
    p.converged = converged;
    p.singular = singular_flag;
    p.n = n;
    p.nE = nE;
    p.side = side;
    p.mode = mode;
    p.imE = imE;
    p.conv_tol = conv_tol;
    p.conv_rep = conv_rep;
    p.min_iter = min_iter;
    p.max_iter = max_iter;
    p.ith = -1;
    p.use_tma = g_use_tma;
    p.herm_shortcut = g_herm_shortcut;
    cd* owned = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    rc = launch_gf(p, st, &owned);
    if (owned) {
        cudaStreamSynchronize(st);
        cudaFree(owned);
    }
    return rc;
}

// Both semi-infinite continuations of ONE periodic lead (on-site block h00, hopping h01 from a cell to the next) from a
// single Sancho-Rubio decimation: the right lead (side +1, cells 0, 1, 2, ...: g_R = (z - h00 - h01 g_R h01^dag)^-1) and the
// left lead (side -1: g_L = (z - h00 - h01^dag g_L h01)^-1).  A junction whose two leads are the same material needs one
// decimation instead of two (wfm.SelfEnergies, wfm/__init__.py:283-287, calls its converger once per lead).
int tx_surface_gf_pair_dev(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, double imE, double conv_tol,
                           int max_iter, tx_cplx* gL, tx_cplx* sigL, tx_cplx* gR, tx_cplx* sigR, int* n_iter, double* conv,
                           int* converged, int* singular_flag, void* stream) {
    int rc = check_gf_args(h00, h01, n, E, nE, 1, TX_GF_SANCHO);
    if (rc) return rc;
    if (!singular_flag) return fail(TX_ERR_ARG, "singular_flag is required");
    if (!gL && !sigL && !gR && !sigR) return fail(TX_ERR_ARG, "no output requested");
    GfParams p{};
    p.h00 = reinterpret_cast<const cd*>(h00);
    p.h01 = reinterpret_cast<const cd*>(h01);
    p.E = reinterpret_cast<const cd*>(E);
    p.g = reinterpret_cast<cd*>(gR);
    p.sigma = reinterpret_cast<cd*>(sigR);
    p.g2 = reinterpret_cast<cd*>(gL);
    p.sigma2 = reinterpret_cast<cd*>(sigL);
    p.n_iter = n_iter;
    p.conv = conv;
    p.converged = converged;
    p.singular = singular_flag;
    p.n = n;
    p.nE = nE;
    p.side = 1;
    p.mode = TX_GF_SANCHO;
    p.imE = imE;
    p.conv_tol = conv_tol;
    p.max_iter = max_iter;
    p.ith = -1;
    p.use_tma = g_use_tma;
    p.herm_shortcut = g_herm_shortcut;
    cd* owned = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    rc = launch_gf(p, st, &owned);
    if (owned) {
        cudaStreamSynchronize(st);
        cudaFree(owned);
    }
    return rc;
}

int tx_surface_gf_pair(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, double imE, double conv_tol,
                       int max_iter, tx_cplx* gL, tx_cplx* sigL, tx_cplx* gR, tx_cplx* sigR, int* n_iter, double* conv,
                       int* converged) {
    int rc = check_gf_args(h00, h01, n, E, nE, 1, TX_GF_SANCHO);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t nn = (size_t)n * n, bm = nn * sizeof(cd), bE = ((size_t)nE * sizeof(cd) + 255) & ~(size_t)255;
    const size_t bg = (size_t)nE * bm, bi = ((size_t)nE * 4 + 255) & ~(size_t)255, bc = ((size_t)nE * 8 + 255) & ~(size_t)255;
    char* d = nullptr;
    TX_CUDA(cudaMalloc(&d, 2 * bm + bE + 4 * bg + 2 * bi + bc + 256));
    auto cleanup = [&](int code) {
        cudaFree(d);
        return code;
    };
    char* q = d;
    auto take = [&](size_t b) { char* r = q; q += b; return r; };
    char *dh = take(bm), *dt = take(bm), *dE = take(bE);
    char* out[4];
    for (int i = 0; i < 4; ++i) out[i] = take(bg);
    char *dni = take(bi), *dok = take(bi), *dcv = take(bc), *dflag = take(256);
    if (cudaMemcpy(dh, h00, bm, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dt, h01, bm, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dE, E, (size_t)nE * sizeof(cd), cudaMemcpyHostToDevice) != cudaSuccess || cudaMemset(dflag, 0, 4) != cudaSuccess)
        return cleanup(fail(TX_ERR_CUDA, "host to device copy failed"));
    tx_cplx* host[4] = {gL, sigL, gR, sigR};
    rc = tx_surface_gf_pair_dev((const tx_cplx*)dh, (const tx_cplx*)dt, n, (const tx_cplx*)dE, nE, imE, conv_tol, max_iter,
                                gL ? (tx_cplx*)out[0] : nullptr, sigL ? (tx_cplx*)out[1] : nullptr, gR ? (tx_cplx*)out[2] : nullptr,
                                sigR ? (tx_cplx*)out[3] : nullptr, (int*)dni, (double*)dcv, (int*)dok, (int*)dflag, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc) return cleanup(rc);
    if (e != cudaSuccess) return cleanup(fail(TX_ERR_CUDA, "surface GF kernel failed: %s", cudaGetErrorString(e)));
    bool bad = false;
    for (int i = 0; i < 4; ++i)
        if (host[i]) bad |= cudaMemcpy(host[i], out[i], bg, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (n_iter) bad |= cudaMemcpy(n_iter, dni, (size_t)nE * 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (conv) bad |= cudaMemcpy(conv, dcv, (size_t)nE * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (converged) bad |= cudaMemcpy(converged, dok, (size_t)nE * 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    int flag = 0;
    bad |= cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (bad) return cleanup(fail(TX_ERR_CUDA, "device to host copy failed"));
    if (flag) return cleanup(fail(TX_ERR_SINGULAR, "singular block in the surface GF iteration"));
    return cleanup(TX_OK);
}

// Host-pointer convenience wrappers ----------------------------------------------------------
static int gf_host(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, int mode,
                   double imE, double conv_tol, int conv_rep, int min_iter, int max_iter, int ith,
                   const tx_cplx* g_prev, tx_cplx* g, tx_cplx* sigma, int* n_iter, double* conv, int* converged) {
    int rc = check_gf_args(h00, h01, n, E, nE, side, mode);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t nn = (size_t)n * n;
    const size_t bm = nn * sizeof(cd), bE = (size_t)nE * sizeof(cd), bg = (size_t)nE * bm;
    char* d = nullptr;
    // layout: h00 | h01 | E | g | sigma | g_prev | n_iter | conv | converged | flag
    const size_t off_h00 = 0, off_h01 = bm, off_E = 2 * bm, off_g = off_E + ((bE + 255) & ~(size_t)255);
    const size_t off_sig = off_g + bg, off_prev = off_sig + bg, off_ni = off_prev + bg;
    const size_t off_cv = off_ni + (((size_t)nE * 4 + 255) & ~(size_t)255);
    const size_t off_ok = off_cv + (size_t)nE * 8, off_flag = off_ok + (((size_t)nE * 4 + 255) & ~(size_t)255);
    const size_t total = off_flag + 256;
    TX_CUDA(cudaMalloc(&d, total));
    auto cleanup = [&](int code) {
        cudaFree(d);
        return code;
    };
    if (cudaMemcpy(d + off_h00, h00, bm, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d + off_h01, h01, bm, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d + off_E, E, bE, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(d + off_flag, 0, 4) != cudaSuccess)
        return cleanup(fail(TX_ERR_CUDA, "host to device copy failed"));
    if (g_prev && cudaMemcpy(d + off_prev, g_prev, bg, cudaMemcpyHostToDevice) != cudaSuccess)
        return cleanup(fail(TX_ERR_CUDA, "host to device copy failed"));
    GfParams p{};
    p.h00 = reinterpret_cast<const cd*>(d + off_h00);
    p.h01 = reinterpret_cast<const cd*>(d + off_h01);
    p.E = reinterpret_cast<const cd*>(d + off_E);
    p.g = g ? reinterpret_cast<cd*>(d + off_g) : nullptr;
    p.sigma = sigma ? reinterpret_cast<cd*>(d + off_sig) : nullptr;
    p.g_prev = g_prev ? reinterpret_cast<const cd*>(d + off_prev) : nullptr;
    p.n_iter = reinterpret_cast<int*>(d + off_ni);
    p.conv = reinterpret_cast<double*>(d + off_cv);
    p.converged = reinterpret_cast<int*>(d + off_ok);
    p.singular = reinterpret_cast<int*>(d + off_flag);
    p.n = n;
    p.nE = nE;
    p.side = side;
    p.mode = mode;
    p.imE = imE;
    p.conv_tol = conv_tol;
    p.conv_rep = conv_rep;
    p.min_iter = min_iter;
    p.max_iter = max_iter;
    p.ith = ith;
    p.use_tma = g_use_tma;
    p.herm_shortcut = g_herm_shortcut;
    cd* owned = nullptr;
    rc = launch_gf(p, nullptr, &owned);
    cudaError_t e = cudaDeviceSynchronize();
    if (owned) cudaFree(owned);
    if (rc) return cleanup(rc);
    if (e != cudaSuccess) return cleanup(fail(TX_ERR_CUDA, "surface GF kernel failed: %s", cudaGetErrorString(e)));
    int flag = 0;
    bool bad = false;
    if (g) bad |= cudaMemcpy(g, d + off_g, bg, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (sigma) bad |= cudaMemcpy(sigma, d + off_sig, bg, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (n_iter) bad |= cudaMemcpy(n_iter, d + off_ni, (size_t)nE * 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (conv) bad |= cudaMemcpy(conv, d + off_cv, (size_t)nE * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (converged) bad |= cudaMemcpy(converged, d + off_ok, (size_t)nE * 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    bad |= cudaMemcpy(&flag, d + off_flag, 4, cudaMemcpyDeviceToHost) != cudaSuccess;
    if (bad) return cleanup(fail(TX_ERR_CUDA, "device to host copy failed"));
    if (flag) return cleanup(fail(TX_ERR_SINGULAR, "singular block in the surface GF iteration"));
    return cleanup(TX_OK);
}

int tx_surface_gf(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, int mode,
                  double imE, double conv_tol, int conv_rep, int min_iter, int max_iter,
                  tx_cplx* g, tx_cplx* sigma, int* n_iter, double* conv, int* converged) {
    return gf_host(h00, h01, n, E, nE, side, mode, imE, conv_tol, conv_rep, min_iter, max_iter, -1, nullptr,
                   g, sigma, n_iter, conv, converged);
}

int tx_g_ith(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, double imE,
             int ith, const tx_cplx* g_prev, tx_cplx* g_out) {
    if (ith < 0) return fail(TX_ERR_ARG, "ith must be >= 0");
    if (ith > 0 && !g_prev) return fail(TX_ERR_ARG, "g_prev required for ith > 0");
    return gf_host(h00, h01, n, E, nE, side, TX_GF_FIXEDPOINT, imE, 0.0, 0, 0, 0, ith, ith > 0 ? g_prev : nullptr,
                   g_out, nullptr, nullptr, nullptr, nullptr);
}

static float g_last_invert_ms = 0.f;
double tx_debug_last_invert_ms(void) { return (double)g_last_invert_ms; }

int tx_debug_invert_staged(const tx_cplx* in, tx_cplx* out, int count, int n) {
    if (!in || !out || count < 1 || n < 1 || n > TX_MAX_N) return fail(TX_ERR_ARG, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    const size_t bytes = (size_t)count * n * n * sizeof(cd);
    const bool big = n > 64, gs = !inverse_stage_in_smem(n);
    cd *din = nullptr, *dout = nullptr, *dstage = nullptr;
    int* dflag = nullptr;
    TX_CUDA(cudaMalloc(&din, bytes));
    TX_CUDA(cudaMalloc(&dout, bytes));
    TX_CUDA(cudaMalloc(&dflag, sizeof(int)));
    if (gs) TX_CUDA(cudaMalloc(&dstage, (size_t)count * inverse_stage_bytes(n)));
    TX_CUDA(cudaMemcpy(din, in, bytes, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemset(dflag, 0, sizeof(int)));
    const size_t smem = gs ? 0 : inverse_stage_bytes(n);
    const int unblocked = getenv("TX_INVERT_UNBLOCKED") ? 1 : (getenv("TX_INVERT_INPLACE") ? 2 : 0);
    const void* kern = big ? (const void*)invert_staged_kernel<GF_THREADS_BIG, 128> : (const void*)invert_staged_kernel<GF_THREADS, 64>;
    TX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    for (int rep = 0; rep < 2; ++rep) {                 // the first launch is a warm-up (same result)
        if (rep == 1) cudaEventRecord(ev0);
        if (big) invert_staged_kernel<GF_THREADS_BIG, 128><<<count, GF_THREADS_BIG, smem>>>((const cd*)din, dout, n, dflag, unblocked, dstage);
        else invert_staged_kernel<GF_THREADS, 64><<<count, GF_THREADS, smem>>>((const cd*)din, dout, n, dflag, unblocked, dstage);
    }
    cudaEventRecord(ev1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) cudaEventElapsedTime(&g_last_invert_ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    int flag = 0;
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(din);
    cudaFree(dout);
    cudaFree(dflag);
    if (dstage) cudaFree(dstage);
    if (e != cudaSuccess) return fail(TX_ERR_CUDA, "invert_staged_kernel failed: %s", cudaGetErrorString(e));
    if (flag) return fail(TX_ERR_SINGULAR, "singular block");
    return TX_OK;
}

static int gemm_bench(int n, int count, int reps, int use_tma, double* ms_out);
int tx_debug_gemm_ms(int n, int count, int reps, double* ms_out) { return gemm_bench(n, count, reps, 0, ms_out); }
int tx_debug_gemm_tma_ms(int n, int count, int reps, double* ms_out) { return gemm_bench(n, count, reps, 1, ms_out); }
static int gemm_bench(int n, int count, int reps, int use_tma, double* ms_out) {
    if (n < 8 || n > 64 || count < 1 || reps < 1 || !ms_out) return fail(TX_ERR_ARG, "bad arguments");
    if (use_tma && ((n & 7) || !inverse_stage_in_smem(n))) return fail(TX_ERR_ARG, "the TMA-staged product needs n % 8 == 0 and n <= 104");
    const size_t smem = use_tma ? tma_stage_bytes(n) : 0;
    if (smem) TX_CUDA(cudaFuncSetAttribute(gemm_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t bytes = (size_t)count * n * n * sizeof(cd);
    cd *dA = nullptr, *dB = nullptr, *dC = nullptr;
    TX_CUDA(cudaMalloc(&dA, bytes));
    TX_CUDA(cudaMalloc(&dB, bytes));
    TX_CUDA(cudaMalloc(&dC, bytes));
    TX_CUDA(cudaMemset(dA, 0, bytes));
    TX_CUDA(cudaMemset(dB, 0, bytes));
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    gemm_bench_kernel<<<count, GF_THREADS, smem>>>(dA, dB, dC, n, 1, use_tma);
    cudaEventRecord(ev0);
    gemm_bench_kernel<<<count, GF_THREADS, smem>>>(dA, dB, dC, n, reps, use_tma);
    cudaEventRecord(ev1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    if (e != cudaSuccess) return fail(TX_ERR_CUDA, "gemm_bench_kernel failed: %s", cudaGetErrorString(e));
    *ms_out = ms;
    return TX_OK;
}

}  // extern "C"
