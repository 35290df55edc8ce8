// K2/K3 for block sizes beyond the register-resident kernel (16 < n <= 128): one CTA per
// (Hamiltonian, energy) point, blocks in shared memory (n <= 40) or an L2-resident global
// workspace, CTA-cooperative GEMM / pivoted Gauss-Jordan from block_ops.cuh.  Two instantiations:
// 256 threads, two CTAs per SM for n <= 64; 512 threads, one CTA per SM for 64 < n <= 128 (the blocked
// inverse wants 4 n threads; run_wfm_gate.get_hblocks gives n = 72 at TwoS = 5, 128 at TwoS = 7).
//
// Same recursion as rgf_small.cuh (replaces wfm.kernel, transport/wfm/__init__.py:29-137, for
// multi-channel leads).  Lead self-energies always come from a table (K1).  Besides the reference's
// per-channel R/T (which use only diag Sigma, :91) it emits the Caroli trace with the full matrix
// Gamma = i(Sigma - Sigma^dag):  Tr[Gamma_R G Gamma_L G^dag], the meaningful total transmission when
// Sigma is not diagonal (SURVEY.md §7 hard part 7, §8d cfg4).
#include <cuda_runtime.h>

#include "block_ops.cuh"
#include "rgf_generic.h"
#include "tx_host.h"

using namespace tx;

namespace tx {

constexpr int GEN_MATS = 5;

template <int GEN_THREADS, int NMAX, int MINB, int FIXN>
__global__ void __launch_bounds__(GEN_THREADS, MINB) rgf_sweep_generic(const GenericParams p) {
    extern __shared__ double2 gen_smem[];
    __shared__ int s_ipiv[NMAX];
    __shared__ cd s_colk[NMAX];
    __shared__ double s_vL[NMAX], s_vR[NMAX];
    __shared__ cd s_zd[NMAX];            // diagonal of z_j at a diagonal site
    __shared__ double s_red[GEN_THREADS / 32];
    __shared__ double s_red4[4][GEN_THREADS / 32];
    __shared__ unsigned long long s_tma_bar;
    // nr: block edge of the inputs and outputs; n: edge of the work blocks, = nr, or nr rounded up to the 8 x 8 tensor-core
    // tile in the telescoped sweep (two spin-1, spin-2, spin-3 impurities give nr = 18, 50, 98): the padding channels are
    // decoupled and carry z = 1 + |t_j|^2, which keeps D_j = 1 and g = 1 there (as in rgf_n8t.cuh), so neither the
    // recurrence nor the inverse meets anything singular, and the products and the inverse run on the tensor cores
    const int nr = p.n, n = p.np, M = p.M, nn = n * n, nnr = nr * nr;
    // persistent grid: the global workspace (blocks too large for shared memory) is indexed by CTA, not by point, so a
    // few hundred of them stay L2-resident however long the sweep is
    cd* base = p.work ? p.work + (long long)blockIdx.x * p.work_per_cta : reinterpret_cast<cd*>(gen_smem);
    // inverse staging when the blocks are global: shared memory, or (n > 104) one more block of the workspace
    cd* stage = p.work ? reinterpret_cast<cd*>(gen_smem) : nullptr;
    if constexpr (NMAX > TX_STAGE_SMEM_MAX_N)
        if (p.work && !inverse_stage_in_smem(n)) stage = base + GEN_MATS * nn;
    TmaStage tma_stage;
    TmaStage* ts = nullptr;   // the products stage their B operand through the same block by TMA (block_ops.cuh)
    // (the staging block must be the shared-memory one; other edges: multiples of 8 run the run-time-n staged product)
    if (stage != nullptr && p.use_tma && (FIXN == 64 || tma_stage_pays(n))) {
        tma_stage_init(tma_stage, stage, &s_tma_bar);
        ts = &tma_stage;
    }
    for (long long pt = blockIdx.x; pt < p.n_pts; pt += gridDim.x) {
    const int ham = (int)(pt / p.nE);
    const int e = (int)(pt - (long long)ham * p.nE);
    const cd E = p.E[e];
    cd* G = base;            // current g_j, finally G_00
    cd* W = base + nn;       // running product, finally G_{M-1,0}
    cd* A = base + 2 * nn;
    cd* X = base + 3 * nn;
    cd* Y = base + 4 * nn;
    const cd* hb = p.h + ham * p.h_stride;
    const cd* tb = p.tnn + ham * p.t_stride;
    const cd* SL = p.sigL + ham * p.sig_stride + (long long)e * nnr;
    const cd* SR = p.sigR + ham * p.sig_stride + (long long)e * nnr;
    const double par = p.h1 ? p.params[ham] : 0.0;
    const int* cls = p.site_class ? p.site_class + ham * p.class_stride : nullptr;

    for (int c = threadIdx.x; c < nr; c += blockDim.x) {     // v = -2 Im diag Sigma  (:91)
        s_vL[c] = -2.0 * SL[c * nr + c].y;
        s_vR[c] = -2.0 * SR[c * nr + c].y;
    }

    // Elementwise passes over blocks that sit in L2 / HBM: four independent loads per thread and array before the first
    // store (otherwise a pass is a chain of dependent memory latencies).  f(idx) -> value to store at dst[idx].
    auto for_each4 = [&](cd* dst, auto load, auto store) {
        const int bd = blockDim.x;
        int idx = threadIdx.x;
        for (; idx + 3 * bd < nn; idx += 4 * bd) {
            auto v0 = load(idx), v1 = load(idx + bd), v2 = load(idx + 2 * bd), v3 = load(idx + 3 * bd);
            dst[idx] = store(idx, v0);
            dst[idx + bd] = store(idx + bd, v1);
            dst[idx + 2 * bd] = store(idx + 2 * bd, v2);
            dst[idx + 3 * bd] = store(idx + 3 * bd, v3);
        }
        for (; idx < nn; idx += bd) dst[idx] = store(idx, load(idx));
        __syncthreads();
    };
    struct ZIn {
        cd h, w, s;
    };
    // z_j = E - h_j (- Sigma on the end blocks) into dst; padding channels: diagonal 1 + padk
    auto build_z = [&](cd* dst, int j, double padk) {
        const cd* hj = hb + (long long)j * nnr;
        const cd* wj = p.h1 ? p.h1 + (long long)j * nnr : nullptr;
        const cd* sg = (j == M - 1) ? SR : (j == 0 ? SL : nullptr);
        for_each4(dst,
                  [&](int idx) {
                      ZIn in;
                      int src = idx;
                      if (n != nr) {
                          const int r = idx / n, c = idx - r * n;
                          src = (r < nr && c < nr) ? r * nr + c : -1;
                      }
                      in.h = src >= 0 ? hj[src] : mk(0.0, 0.0);
                      in.w = (wj && src >= 0) ? wj[src] : mk(0.0, 0.0);
                      in.s = (sg && src >= 0) ? sg[src] : mk(0.0, 0.0);
                      return in;
                  },
                  [&](int idx, ZIn in) {
                      const int r = idx / n, c = idx - r * n;
                      if (r >= nr || c >= nr) return mk(r == c ? 1.0 + padk : 0.0, 0.0);
                      cd hv = in.h;
                      if (wj) {
                          hv.x = fma(par, in.w.x, hv.x);
                          hv.y = fma(par, in.w.y, hv.y);
                      }
                      cd a = (r == c) ? mk(E.x - hv.x, E.y - hv.y) : mk(-hv.x, -hv.y);
                      if (sg) a = a - in.s;
                      return a;
                  });
    };

    if (p.hop_scalar && p.kmax > 1) {
        // ---- telescoped sweep (see rgf_n8t.cuh): inside a chunk [a..b]
        //   D_{b+1} = I, D_{b+2} := g_{b+1},  D_j = z_j D_{j+1} - |t_j|^2 D_{j+2},  X = D_a^-1,  g_a = D_{a+1} X,
        //   W_a = (prod conj t_j) W_{b+1} X  -- one product per site, one inverse per chunk.
        cd* cur = A;      // D_{j+1}
        cd* prev = X;     // D_{j+2}
        cd* tmp = Y;      // z_j / scratch
        const double gmax2 = ldexp(1.0, 2 * p.gexp + 2);
        int j = M - 1;
        while (j >= 0) {
            const int b = j;
            cd tau = mk(1.0, 0.0);
            double S2 = norm2(tb[(long long)((b < M - 1) ? b : M - 2) * nnr]);
            build_z(cur, b, b < M - 1 ? norm2(tb[(long long)b * nnr]) : 0.0);
            if (b < M - 1) {
                const cd t0 = tb[(long long)b * nnr];
                const double kap = norm2(t0);
                tau = conj(t0);
                struct Two {
                    cd a, g;
                };
                for_each4(cur, [&](int idx) { Two t; t.a = cur[idx]; t.g = G[idx]; return t; },
                          [&](int, Two t) { return mk(fma(-kap, t.g.x, t.a.x), fma(-kap, t.g.y, t.a.y)); });
            }
            int k = 1;
            bool prev_is_identity = true;
            while (j > 0 && k < p.kmax) {
                const cd t0 = tb[(long long)(j - 1) * nnr];
                const double kap = norm2(t0);
                if (!(kap > 0.0)) break;                                   // decoupled site: its own chunk
                if (!(cta_maxnorm2(cur, nn) < gmax2 * S2)) break;          // growth bound (also stops on NaN)
                --j;
                ++k;
                if (cls != nullptr && j > 0 && cls[j] >= 1) {
                    // DIAGONAL on-site block (clean spacer sites between the impurities, barrier regions: most sites of the
                    // reference's own chains): z_j D_{j+1} is a row scaling, O(n^2) instead of a product
                    const cd* hj = hb + (long long)j * nnr;
                    const cd* wj = p.h1 ? p.h1 + (long long)j * nnr : nullptr;
                    for (int r = threadIdx.x; r < n; r += blockDim.x) {
                        if (r < nr) {
                            cd hv = hj[r * nr + r];
                            if (wj) {
                                const cd w = wj[r * nr + r];
                                hv.x = fma(par, w.x, hv.x);
                                hv.y = fma(par, w.y, hv.y);
                            }
                            s_zd[r] = mk(E.x - hv.x, E.y - hv.y);
                        } else {
                            s_zd[r] = mk(1.0 + kap, 0.0);                      // padding channels keep D = 1
                        }
                    }
                    __syncthreads();
                    struct Two {
                        cd c, q;
                    };
                    const bool ident = prev_is_identity;
                    for_each4(prev, [&](int idx) { Two t; t.c = cur[idx]; t.q = ident ? mk(0.0, 0.0) : prev[idx]; return t; },
                              [&](int idx, Two t) {
                                  const int r = idx / n, c = idx - r * n;
                                  const cd q = ident ? mk(r == c ? 1.0 : 0.0, 0.0) : t.q;
                                  const cd v = s_zd[r] * t.c;
                                  return mk(fma(-kap, q.x, v.x), fma(-kap, q.y, v.y));
                              });
                } else {
                build_z(tmp, j, kap);
                // D_j = z_j D_{j+1} - kap D_{j+2}: the scaling of D_{j+2} rides on the product's epilogue
                if (prev_is_identity) {
                    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
                        const int r = idx / n, c = idx - r * n;
                        prev[idx] = mk(r == c ? -kap : 0.0, 0.0);
                    }
                    __syncthreads();
                    cta_gemm_acc<FIXN>(prev, tmp, OP_N, cur, OP_N, n, 1.0, 1.0, ts);
                } else {
                    cta_gemm_acc<FIXN>(prev, tmp, OP_N, cur, OP_N, n, 1.0, -kap, ts);
                }
                }
                cd* sw = prev;
                prev = cur;
                cur = sw;
                prev_is_identity = false;
                S2 *= kap;
                tau = tau * conj(t0);
            }
            // close at site a = j:  cur = D_a, prev = D_{a+1} (identity if k == 1)
            cta_inverse_staged<NMAX>(cur, stage, n, s_colk, s_ipiv, p.singular);   // X
            if (prev_is_identity) cta_copy(G, cur, nn);
            else cta_gemm<FIXN>(G, prev, OP_N, cur, OP_N, n, ts);                  // g_a = D_{a+1} X
            if (b == M - 1) {
                for_each4(W, [&](int idx) { return cur[idx]; }, [&](int, cd v) { return v * tau; });
            } else {
                cta_gemm<FIXN>(tmp, W, OP_N, cur, OP_N, n, ts);                    // W X
                for_each4(W, [&](int idx) { return tmp[idx]; }, [&](int, cd v) { return v * tau; });
            }
            --j;
        }
    } else
    for (int j = M - 1; j >= 0; --j) {
        build_z(A, j, 0.0);
        if (j < M - 1) {
            const cd* t = tb + (long long)j * nn;
            if (p.hop_scalar) {
                const cd t0 = t[0];
                const double kap = norm2(t0);
                for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
                    cd a = A[idx];
                    const cd g = G[idx];
                    a.x = fma(-kap, g.x, a.x);
                    a.y = fma(-kap, g.y, a.y);
                    A[idx] = a;
                    W[idx] = W[idx] * conj(t0);                 // W t^dag
                }
                __syncthreads();
            } else {
                cta_gemm<FIXN>(X, t, OP_N, G, OP_N, n);               // t g
                cta_gemm_acc<FIXN>(A, X, OP_N, t, OP_H, n, -1.0);     // A -= (t g) t^dag
                cta_gemm<FIXN>(X, W, OP_N, t, OP_H, n);               // W t^dag
                cta_copy(W, X, nn);
            }
        }
        cta_inverse_staged<NMAX>(A, stage, n, s_colk, s_ipiv, p.singular);
        if (j == M - 1) {
            cta_copy(G, A, nn);
            cta_copy(W, A, nn);
        } else {
            cta_gemm<FIXN>(X, W, OP_N, A, OP_N, n);                   // (W t^dag) g_j
            cta_copy(W, X, nn);
            cta_copy(G, A, nn);
        }
    }

    // ---- epilogue: per-channel R/T with the reference's formulas (:108-130)
    double tsum = 0.0;
    if (p.R || p.ttot || p.gather.n || p.resid || p.spin) {
        for (int i = 0; i < p.n_src; ++i) {
            double f2 = 0.0, wu = 0.0, wd = 0.0;   // wu, wd: incident weight per electron-spin sector
            if (p.sources) {
                for (int c = 0; c < nr; ++c) {
                    const double a = p.sources[(long long)i * nr + c];
                    f2 = fma(a, a * s_vL[c], f2);
                    if (c < p.n_up) wu = fma(a, a, wu); else wd = fma(a, a, wd);
                }
            } else {                               // one-hot source i: O(n) per source instead of O(n^2)
                f2 = s_vL[i];
                wu = i < p.n_up ? 1.0 : 0.0;
                wd = 1.0 - wu;
            }
            const bool src_up = wu >= wd;
            double sp[4] = {0.0, 0.0, 0.0, 0.0};               // T same / other sector, R same / other (emit_spin)
            const double rflux = 1.0 / sqrt(f2);
            double part = 0.0;
            for (int s = threadIdx.x; s < nr; s += blockDim.x) {
                double ur = 0, ui = 0, wr = 0, wi = 0, mine = 0;
                if (p.sources) {
                    for (int c = 0; c < nr; ++c) {
                        const double a = p.sources[(long long)i * nr + c];
                        const double av = a * s_vL[c];
                        const cd g = G[s * n + c], w = W[s * n + c];
                        ur = fma(g.x, av, ur);
                        ui = fma(g.y, av, ui);
                        wr = fma(w.x, av, wr);
                        wi = fma(w.y, av, wi);
                        if (c == s) mine = a;
                    }
                } else {
                    const double av = s_vL[i];
                    const cd g = G[s * n + i], w = W[s * n + i];
                    ur = fma(g.x, av, 0.0);
                    ui = fma(g.y, av, 0.0);
                    wr = fma(w.x, av, 0.0);
                    wi = fma(w.y, av, 0.0);
                    mine = (s == i) ? 1.0 : 0.0;
                }
                const double fr = sqrt(s_vL[s]) * rflux, ft = sqrt(s_vR[s]) * rflux;
                const double xr = (-ui - mine) * fr, xi = ur * fr;
                const double yr = -wi * ft, yi = wr * ft;
                const double rr = fma(xr, xr, xi * xi), tt = fma(yr, yr, yi * yi);
                if (p.R) {
                    p.R[(pt * p.n_src + i) * nr + s] = rr;
                    p.T[(pt * p.n_src + i) * nr + s] = tt;
                }
                part += rr + tt;
                tsum += tt;
                const bool same = (s < p.n_up) == src_up;
                sp[0] += same ? tt : 0.0;
                sp[1] += same ? 0.0 : tt;
                sp[2] += same ? rr : 0.0;
                sp[3] += same ? 0.0 : rr;
            }
            if (p.spin) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) sp[q] += __shfl_xor_sync(0xffffffffu, sp[q], off);
                    if ((threadIdx.x & 31) == 0) s_red4[q][threadIdx.x >> 5] = sp[q];
                }
                __syncthreads();
                if (threadIdx.x < p.spin_n) {
                    double tot = 0.0;
                    for (int w = 0; w < GEN_THREADS / 32; ++w) tot += s_red4[threadIdx.x][w];
                    p.spin[(pt * p.n_src + i) * p.spin_n + threadIdx.x] = tot;
                }
                __syncthreads();
            }
            if (p.resid) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double tot = 0.0;
                    for (int w = 0; w < GEN_THREADS / 32; ++w) tot += s_red[w];
                    p.resid[pt * p.n_src + i] = tot - 1.0;
                }
                __syncthreads();
            }
        }
    }
    if (p.ttot || p.gather.n) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < GEN_THREADS / 32; ++w) tot += s_red[w];
            store_total(p.ttot, p.gather, pt, tot);
        }
        __syncthreads();
    }
    // ---- Caroli trace Tr[Gamma_R W Gamma_L W^dag], Gamma = i (Sigma - Sigma^dag)
    if (p.caroli) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const int r = idx / n, c = idx - r * n;
            if (r >= nr || c >= nr) {                           // padding channels carry no Gamma
                A[idx] = mk(0.0, 0.0);
                X[idx] = mk(0.0, 0.0);
                continue;
            }
            const cd a = SR[r * nr + c], b = conj(SR[c * nr + r]);
            A[idx] = mk(-(a.y - b.y), a.x - b.x);               // i (a - b)
            const cd a2 = SL[r * nr + c], b2 = conj(SL[c * nr + r]);
            X[idx] = mk(-(a2.y - b2.y), a2.x - b2.x);
        }
        __syncthreads();
        cta_gemm<FIXN>(Y, A, OP_N, W, OP_N, n);                       // Gamma_R W
        cta_gemm<FIXN>(A, Y, OP_N, X, OP_N, n);                       // Gamma_R W Gamma_L
        double tr = 0.0;
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const cd y = A[idx], w = W[idx];
            tr += y.x * w.x + y.y * w.y;                        // Re sum_ij Y_ij conj(W_ij)
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, off);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tr;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < GEN_THREADS / 32; ++w) tot += s_red[w];
            p.caroli[pt] = tot;
        }
    }
    __syncthreads();   // the static shared arrays and the workspace are reused by the next point
    }   // persistent loop over points
}

// Number of CTAs the generic sweep keeps resident (its grid; the workspace is per CTA).
constexpr int GEN_THREADS_SMALL = 256, GEN_THREADS_BIG = 512;
static int generic_padded_edge(int n) { return n > 16 ? ((n + 7) & ~7) : n; }

// instantiation by block edge: 0: n < 64 (256 threads), 1: n = 64 (256 threads, compile-time products, TMA staging), 2: n > 64
static int generic_kind(int n) { return n > 64 ? 2 : (n == 64 ? 1 : 0); }
static const void* generic_kernel(int kind) {
    return kind == 2 ? (const void*)rgf_sweep_generic<GEN_THREADS_BIG, 128, 1, 0>
                     : (kind == 1 ? (const void*)rgf_sweep_generic<GEN_THREADS_SMALL, 64, 2, 64> : (const void*)rgf_sweep_generic<GEN_THREADS_SMALL, 64, 2, 0>);
}

static int generic_resident_ctas(size_t smem, int kind) {
    static int cached[16][3] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kind == 2 ? 148 : 296;
    int& slot = cached[dev & 15][kind];
    if (!slot) {
        int per_sm = 0, sms = 0;
        if (smem > 16 * 1024) cudaFuncSetAttribute(generic_kernel(kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic_kernel(kind), kind == 2 ? GEN_THREADS_BIG : GEN_THREADS_SMALL, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slot = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    return slot;
}

// dynamic shared memory when the blocks live in the global workspace: inverse staging and the TMA-staged product operand
static size_t generic_stage_bytes(int n) {
    size_t b = inverse_stage_in_smem(n) ? inverse_stage_bytes(n) : 0;
    if ((n == 64 || tma_stage_pays(n)) && tma_stage_bytes(n) > b) b = tma_stage_bytes(n);
    return b;
}

// elements of the global workspace one resident CTA owns: GEN_MATS blocks (+ the inverse staging block beyond n = 104)
static size_t generic_per_cta_elems(int n) {
    return (size_t)GEN_MATS * n * n + (inverse_stage_in_smem(n) ? 0 : (size_t)n * (n + 1));
}

// Host-side launcher used by capi.cu.  `work` must hold rgf_generic_workspace_elems(n, n_pts) elements when the blocks do
// not fit in shared memory: GEN_MATS blocks per RESIDENT CTA (bounded: a persistent grid strides over the points), not per point.
size_t rgf_generic_workspace_elems(int n_in, long long n_pts) {
    const int n = generic_padded_edge(n_in);     // sized for the padded work blocks whichever sweep runs
    const size_t bytes = (size_t)GEN_MATS * n * n * sizeof(cd);
    if (bytes <= 200 * 1024) return 0;
    long long ctas = generic_resident_ctas(generic_stage_bytes(n), generic_kind(n));
    if (n_pts < ctas) ctas = n_pts;
    return (size_t)ctas * generic_per_cta_elems(n);
}

int launch_rgf_generic(GenericParams p, cudaStream_t st) {
    if (p.n < 1 || p.n > TX_MAX_N) return fail(TX_ERR_UNSUPPORTED, "block size n=%d outside 1..%d", p.n, TX_MAX_N);
    // work-block edge: the telescoped sweep pads block edges beyond 16 to a multiple of the tensor-core tile
    const int np = (p.hop_scalar && p.kmax > 1) ? generic_padded_edge(p.n) : p.n;
    p.np = np;
    const int kind = generic_kind(np);
    const size_t bytes = (size_t)GEN_MATS * np * np * sizeof(cd);
    size_t smem = bytes;
    if (bytes > 200 * 1024) {
        if (!p.work) return fail(TX_ERR_ARG, "generic sweep needs a workspace for n=%d", p.n);
        smem = generic_stage_bytes(np);          // staging block for the inverse and the TMA-staged product operand
        p.work_per_cta = (long long)generic_per_cta_elems(np);
    } else {
        p.work = nullptr;
    }
    p.use_tma = tma_mode();
    if (smem > 16 * 1024)   // the kernel also has 11 KB (n <= 64) / 26 KB of static shared memory
        TX_CUDA(cudaFuncSetAttribute(generic_kernel(kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = p.n_pts;
    if (p.work) {
        const long long ctas = generic_resident_ctas(smem, kind);
        if (grid > ctas) grid = ctas;
    } else if (grid > 0x7fffffffLL) {
        return fail(TX_ERR_ARG, "too many points for one launch");
    }
    if (kind == 2) rgf_sweep_generic<GEN_THREADS_BIG, 128, 1, 0><<<(unsigned)grid, GEN_THREADS_BIG, smem, st>>>(p);
    else if (kind == 1) rgf_sweep_generic<GEN_THREADS_SMALL, 64, 2, 64><<<(unsigned)grid, GEN_THREADS_SMALL, smem, st>>>(p);
    else rgf_sweep_generic<GEN_THREADS_SMALL, 64, 2, 0><<<(unsigned)grid, GEN_THREADS_SMALL, smem, st>>>(p);
    count_launch();
    TX_CUDA(cudaGetLastError());
    return TX_OK;
}

}  // namespace tx
