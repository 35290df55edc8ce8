// K2 (+K1a prologue, +K3 epilogue): batched block-tridiagonal Green's-function sweep for small
// blocks (n_loc_dof <= 16), one (Hamiltonian, energy) point per group of LP lanes.
//
// What it replaces in the reference (cpbunker/transport), per point:
//   SelfEnergies      transport/wfm/__init__.py:270-304   (closed-form leads fused in the prologue)
//   Hmat / Hprime     :168-268   (never assembled: blocks are consumed in place)
//   Green             :139-166   (dense (M n)^3 inverse -> O(M n^3) sweep, only G_00 and G_{M-1,0})
//   mat_2d_to_4d      transport/tdfci/utils.py:290-316 (not needed)
//   R/T channel loop  transport/wfm/__init__.py:91, :108-130 (fused epilogue)
//
// Algorithm (right-to-left):
//   A_j = E - h_j - t_j g_{j+1} t_j^dag   (A_{M-1} = E - h_{M-1} - Sigma_R ; A_0 also - Sigma_L)
//   g_j = A_j^{-1}                        (in-place Gauss-Jordan, implicit partial pivoting)
//   W_j = W_{j+1} t_j^dag g_j             (W_{M-1} = g_{M-1})
//   G_00 = g_0,  G_{M-1,0} = W_0.
//
// Data layout: a point's N x N complex block is spread over LP lanes, each lane holding RP = N/LP
// full rows in registers (complex128 as separate re/im doubles).  The only cross-lane traffic is
// (a) the pivot search (a 64-bit max butterfly over LP lanes), (b) the pivot row and (c) the
// finished inverse, both exchanged through shared memory with 16-byte broadcast loads
// (LDS.128), laid out [element][group] so the GPW groups of a warp read one contiguous 128-byte
// line per load.  No __syncthreads: groups never leave their warp, __syncwarp orders the exchange.
#pragma once
#include "tx_common.cuh"

namespace tx {

struct SweepParams {
    const cd* h;             // [n_h][M][n][n]
    long long h_stride;      // elements between Hamiltonians (0 = shared)
    const cd* h1;            // [M][n][n] or null (affine term)
    const double* params;    // [n_ham] or null
    const cd* tnn;           // [n_t][M-1][n][n]
    long long t_stride;
    const cd* E;             // [nE]
    const cd* sigL;          // TABLE mode: [n_sig][nE][n][n]
    const cd* sigR;
    long long sig_stride;    // elements between Hamiltonians (0 = shared)
    const double* sources;   // [n_src][n] or null (= all n one-hot sources)
    double* R;               // [n_pts][n_src][n]
    double* T;
    double* resid;           // [n_pts][n_src] or null
    double* ttot;            // [n_pts] or null: sum over sources and channels of T (= Tr[Gamma_R G Gamma_L G^dag] for all sources)
    double* spin;            // [n_pts][n_src][spin_n] or null: spin-resolved channel sums (emit_spin below)
    int n_up, spin_n;        // channels [0, n_up) carry electron spin up; spin_n = 2 (T only) or 4 (T and R)
    PeerGather gather;       // n = 0: none; else the per-point total also goes to the peers' buffers
    int* singular;           // device flag, set to 1 if a zero pivot was met
    const int* nondiag_max;  // [n_ham or 1] or null: largest site index whose on-site block is not diagonal (-1: none)
    long long nondiag_stride;
    const int* site_class;   // [n_ham or 1][M] or null: 0 general, 1 diagonal, 2 scalar multiple of I
    long long class_stride;
    const int* class_count;  // null, or device counter: interior sites of the batch with a diagonal on-site block
    long long class_sites;   // interior sites of the batch (rgf_sweep_n8t picks its instantiation from the ratio)
    long long n_pts;
    int n, M, nE, n_src;
};

enum { HOP_SCALAR = 0, HOP_GENERAL = 1 };
enum { LEAD_CLOSED = 0, LEAD_TABLE = 1 };

template <int N, int LP, int HOP>
struct SmallCfg {
    static constexpr int RP = N / LP;
    static constexpr int GPW = 32 / LP;
    // per group, in 16-byte units: piv[2][N], gm[N][N], (tm[N][N]), vv[N], gl[2][N]
    static constexpr int PER_GROUP = 2 * N + N * N + (HOP == HOP_GENERAL ? N * N : 0) + N + 2 * N;
    static constexpr int PER_WARP_BYTES = PER_GROUP * GPW * 16;
};

// acc += a*b on interleaved (re, im) pairs (4 DFMA)
__device__ __forceinline__ void cfma2(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a*b
__device__ __forceinline__ void cfms2(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 cmul2(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

// Spin-resolved channel sums of one source, fused into the K3 epilogue (the reduction the reference's drivers do
// on the host: run_wfm_gate.py:404-405,473-477 split the n_loc_dof channels at n_mol_dof into electron spin up /
// down).  Channels [0, n_up) are the "up" sector.  Output per (point, source):
//     spin[0] = sum of T over the channels of the SOURCE's sector   (no spin flip)
//     spin[1] = sum of T over the other sector                      (spin flip)
//     spin[2], spin[3] = the same for R                             (only when spin_n == 4)
// The LP lanes of a point each hold RP channels (row0 ...); lane li == 0 stores.
template <int RP, int LP>
__device__ __forceinline__ void emit_spin(const SweepParams& p, long long pt, int i, bool src_up, bool active, int li,
                                          int row0, int n, const double (&rr)[RP], const double (&tt)[RP]) {
    double ts = 0.0, to = 0.0, rs = 0.0, ro = 0.0;
#pragma unroll
    for (int s = 0; s < RP; ++s) {
        const int r = row0 + s;
        if (r < n) {
            const bool same = (r < p.n_up) == src_up;
            ts += same ? tt[s] : 0.0;
            to += same ? 0.0 : tt[s];
            rs += same ? rr[s] : 0.0;
            ro += same ? 0.0 : rr[s];
        }
    }
#pragma unroll
    for (int off = LP / 2; off >= 1; off >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, off);
        to += __shfl_xor_sync(0xffffffffu, to, off);
    }
    if (p.spin_n == 4) {
#pragma unroll
        for (int off = LP / 2; off >= 1; off >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, off);
            ro += __shfl_xor_sync(0xffffffffu, ro, off);
        }
    }
    if (active && li == 0) {
        double* o = p.spin + (pt * p.n_src + i) * p.spin_n;
        *reinterpret_cast<double2*>(o) = make_double2(ts, to);
        if (p.spin_n == 4) *reinterpret_cast<double2*>(o + 2) = make_double2(rs, ro);
    }
}

// O = L * B,  L: own rows in registers, B: N x N in shared memory ([k*N+c][group]).
// Blocks are kept as interleaved (re, im) double2 so that a 16-byte shared-memory access maps
// onto an aligned register quad without repacking moves.
template <int N, int RP, int GPW>
__device__ __forceinline__ void rows_times_smem(const double2 (&L)[RP][N], const double2* B, int grp,
                                                double2 (&O)[RP][N]) {
#pragma unroll
    for (int s = 0; s < RP; ++s)
#pragma unroll
        for (int c = 0; c < N; ++c) O[s][c] = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const double2 b = B[(k * N + c) * GPW + grp];
#pragma unroll
            for (int s = 0; s < RP; ++s) cfma2(O[s][c], L[s][k], b);
        }
}

// In-place inverse of the N x N block whose rows are spread over the LP lanes of a group.
// Gauss-Jordan without row exchanges: at step k the not-yet-used row with the largest |A[r][k]|
// is the pivot row p_k; every other row is reduced with it, the pivot row itself is left unscaled
// (one scaling per row at the end).  The result is scattered into shared memory in canonical
// order:  inv[k][p_j] = A[p_k][j] / d_k.
template <int N, int LP, int PIVX = 0>
__device__ __forceinline__ void invert_to_smem(double2 (&A)[N / LP][N], double2* piv,
                                               double2* gm, int grp, int row0, int* singular) {
    constexpr int RP = N / LP;
    constexpr int GPW = 32 / LP;
    unsigned long long perm = 0ull;   // nibble k = pivot row of step k
    unsigned used = 0u;               // bit s: local row s already used as a pivot row
    int krow[RP];                     // step at which local row s was the pivot row
    double2 dinv[RP];                 // reciprocal of its pivot
#pragma unroll
    for (int s = 0; s < RP; ++s) {
        krow[s] = 0;
        dinv[s] = make_double2(0.0, 0.0);
    }

#pragma unroll
    for (int k = 0; k < N; ++k) {
        // ---- pivot search: 64-bit key = |a|^2 bits (low 5 bits replaced by 0x10 | (15-row))
        unsigned long long key = 0ull;
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const double mag = fma(A[s][k].x, A[s][k].x, A[s][k].y * A[s][k].y);
            unsigned long long kk = ((unsigned long long)__double_as_longlong(mag) & ~0x1Full) |
                                    (unsigned long long)(0x10 | (15 - (row0 + s)));
            if ((used >> s) & 1u) kk = 0ull;
            key = kk > key ? kk : key;
        }
#pragma unroll
        for (int off = LP / 2; off >= 1; off >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, off);
            key = other > key ? other : key;
        }
        const int prow = 15 - (int)(key & 0xFull);
        if ((key >> 5) == 0ull) *singular = 1;
        perm |= (unsigned long long)prow << (4 * k);

        double2 pr[N];
        if (PIVX == 1) {
            // ---- pivot row by register shuffle from the owner lane (no shared-memory round trip)
            const int src = ((threadIdx.x & 31) & ~(LP - 1)) | (prow / RP);
            const int sloc = prow % RP;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double2 v = A[0][c];
#pragma unroll
                for (int s = 1; s < RP; ++s)
                    if (sloc == s) v = A[s][c];
                pr[c].x = __shfl_sync(0xffffffffu, v.x, src);
                pr[c].y = __shfl_sync(0xffffffffu, v.y, src);
            }
        } else {
            // ---- the owner publishes the pivot row (double-buffered: one __syncwarp per step)
            double2* pb = piv + (k & 1) * N * GPW;
#pragma unroll
            for (int s = 0; s < RP; ++s)
                if (row0 + s == prow) {
#pragma unroll
                    for (int c = 0; c < N; ++c) pb[c * GPW + grp] = A[s][c];
                }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < N; ++c) pr[c] = pb[c * GPW + grp];
        }
        const cd invc = crecip(mk(pr[k].x, pr[k].y));
        const double2 inv = make_double2(invc.x, invc.y);

        // ---- eliminate column k from every other row; column k then holds the inverse's column
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const bool isp = (row0 + s == prow);
            double2 m = cmul2(A[s][k], inv);
            if (isp) {
                m = make_double2(0.0, 0.0);
                used |= 1u << s;
                krow[s] = k;
                dinv[s] = inv;
            }
#pragma unroll
            for (int c = 0; c < N; ++c)
                if (c != k) cfms2(A[s][c], m, pr[c]);
            A[s][k] = isp ? make_double2(1.0, 0.0) : make_double2(-m.x, -m.y);
        }
    }

    // ---- scale rows and scatter into canonical positions
#pragma unroll
    for (int s = 0; s < RP; ++s) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int col = (int)((perm >> (4 * j)) & 0xFull);
            gm[(krow[s] * N + col) * GPW + grp] = cmul2(A[s][j], dinv[s]);
        }
    }
    __syncwarp();
}

template <int N, int LP, int HOP, int LEAD, int MINB, int PIVX = 0>
__global__ void __launch_bounds__(128, MINB) rgf_sweep_small(const SweepParams p) {
    using Cfg = SmallCfg<N, LP, HOP>;
    constexpr int RP = Cfg::RP;
    constexpr int GPW = Cfg::GPW;
    extern __shared__ double2 smem_all[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / LP;
    const int li = lane % LP;
    const int row0 = li * RP;
    const int n = p.n;
    const int M = p.M;
    const bool full = (n == N);

    double2* wsm = smem_all + (size_t)warp * (Cfg::PER_GROUP * GPW);
    double2* piv = wsm;                        // [2][N][GPW]
    double2* gm = piv + 2 * N * GPW;           // [N*N][GPW]   current inverse g_j (canonical)
    double2* tm = gm + N * N * GPW;            // [N*N][GPW]   t_j^dagger (general hopping only)
    double2* vv = tm + (HOP == HOP_GENERAL ? N * N * GPW : 0);   // [N][GPW]  (v_L, v_R)
    double2* gl = vv + N * GPW;                // [2][N][GPW]  closed-form g_L, g_R per channel

    long long pt = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * GPW + grp;
    const bool active = pt < p.n_pts;
    if (!active) pt = p.n_pts - 1;
    const int ham = (int)(pt / p.nE);
    const int e = (int)(pt - (long long)ham * p.nE);
    const cd E = p.E[e];
    const long long nn = (long long)n * n;
    const double2* hbase = reinterpret_cast<const double2*>(p.h) + ham * p.h_stride;
    const double2* h1base = reinterpret_cast<const double2*>(p.h1);
    const double2* tbase = reinterpret_cast<const double2*>(p.tnn) + ham * p.t_stride;
    const double par = p.h1 ? p.params[ham] : 0.0;

    auto h_at = [&](int j, int r, int c) -> double2 {
        double2 v = hbase[(long long)j * nn + r * n + c];
        if (h1base) {
            const double2 w = h1base[(long long)j * nn + r * n + c];
            v.x = fma(par, w.x, v.x);
            v.y = fma(par, w.y, v.y);
        }
        return v;
    };
    auto t_at = [&](int j, int r, int c) -> double2 { return tbase[(long long)j * nn + r * n + c]; };
    auto tocd = [](double2 v) -> cd { return mk(v.x, v.y); };

    // ------------------------------------------------------------------ prologue: leads
    double2 sigLd[RP], sigRd[RP];   // diagonal Sigma of own rows (scalar-hopping closed leads)
#pragma unroll
    for (int s = 0; s < RP; ++s) sigLd[s] = sigRd[s] = make_double2(0.0, 0.0);

    if (LEAD == LEAD_CLOSED) {
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const int r = row0 + s;
            cd gL = mk(0.0, 0.0), gR = mk(0.0, 0.0);
            if (full || r < n) {
                const cd tL = tocd(t_at(0, r, r)), tR = tocd(t_at(M - 2, r, r));
                gL = g_closed_channel(E, tocd(h_at(0, r, r)), tL);        // wfm/__init__.py:283,286
                gR = g_closed_channel(E, tocd(h_at(M - 1, r, r)), tR);    // wfm/__init__.py:284,287
                if (HOP == HOP_SCALAR) {
                    const cd sl = conj(tL) * (gL * tL);                   // :301 for diagonal blocks
                    const cd sr = tR * (gR * conj(tR));                   // :302
                    sigLd[s] = make_double2(sl.x, sl.y);
                    sigRd[s] = make_double2(sr.x, sr.y);
                }
            }
            if (HOP == HOP_SCALAR) {
                vv[r * GPW + grp] = make_double2(-2.0 * sigLd[s].y, -2.0 * sigRd[s].y);   // :91
            } else {
                gl[r * GPW + grp] = make_double2(gL.x, gL.y);
                gl[(N + r) * GPW + grp] = make_double2(gR.x, gR.y);
            }
        }
        __syncwarp();
    }

    double2 A[RP][N];   // A_j rows, then scratch
    double2 W[RP][N];   // W_j rows

    // own rows of a self-energy (general-hopping closed leads or table mode) subtracted from A;
    // also publishes the velocities v = -2 Im diag Sigma (wfm/__init__.py:91)
    auto subtract_sigma_rows = [&](bool left) {
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const int r = row0 + s;
            double dimag = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double2 sg = make_double2(0.0, 0.0);
                if (full || (r < n && c < n)) {
                    if (LEAD == LEAD_TABLE) {
                        const double2* tab = reinterpret_cast<const double2*>(left ? p.sigL : p.sigR) +
                                             ham * p.sig_stride + (long long)e * nn;
                        sg = tab[r * n + c];
                    } else {
                        const int jt = left ? 0 : M - 2;
                        for (int k = 0; k < n; ++k) {
                            const double2 gk = gl[((left ? 0 : N) + k) * GPW + grp];
                            double2 a, b;
                            if (left) {   // Sigma_L[r][c] = sum_k conj(t0[k][r]) g_L[k] t0[k][c]
                                const double2 t1 = t_at(jt, k, r);
                                a = make_double2(t1.x, -t1.y);
                                b = cmul2(gk, t_at(jt, k, c));
                            } else {      // Sigma_R[r][c] = sum_k tN[r][k] g_R[k] conj(tN[c][k])
                                const double2 t2 = t_at(jt, c, k);
                                a = t_at(jt, r, k);
                                b = cmul2(gk, make_double2(t2.x, -t2.y));
                            }
                            cfma2(sg, a, b);
                        }
                    }
                }
                A[s][c].x -= sg.x;
                A[s][c].y -= sg.y;
                if (c == r) dimag = sg.y;
            }
            if (left)
                vv[r * GPW + grp].x = -2.0 * dimag;
            else
                vv[r * GPW + grp].y = -2.0 * dimag;
        }
    };

    // ------------------------------------------------------------------ sweep
    for (int j = M - 1; j >= 0; --j) {
        double2 tj = make_double2(0.0, 0.0);
        if (j < M - 1 && HOP == HOP_SCALAR) tj = t_at(j, 0, 0);

        // A = E - h_j
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const int r = row0 + s;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double2 v = make_double2(0.0, 0.0);
                if (full || (r < n && c < n)) v = h_at(j, r, c);
                A[s][c] = make_double2(-v.x, -v.y);
                if (c == r) {
                    if (full || r < n) {
                        A[s][c].x += E.x;
                        A[s][c].y += E.y;
                    } else {
                        A[s][c] = make_double2(1.0, 0.0);   // padding channel: identity block
                    }
                }
            }
        }

        if (j < M - 1) {
            if (HOP == HOP_SCALAR) {
                // A -= |t_j|^2 g_{j+1}   (own rows of g from shared memory)
                const double kap = fma(tj.x, tj.x, tj.y * tj.y);
#pragma unroll
                for (int s = 0; s < RP; ++s) {
                    const int r = row0 + s;
                    if (full || r < n) {
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const double2 g = gm[(r * N + c) * GPW + grp];
                            A[s][c].x = fma(-kap, g.x, A[s][c].x);
                            A[s][c].y = fma(-kap, g.y, A[s][c].y);
                        }
                    }
                }
            } else {
                // general hopping: X = t_j g_{j+1},  A -= X t_j^dag,  W <- W t_j^dag
                double2 Tm[RP][N], X[RP][N];
#pragma unroll
                for (int s = 0; s < RP; ++s) {
                    const int r = row0 + s;
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        double2 v = make_double2(0.0, 0.0);
                        if (full || (r < n && c < n)) v = t_at(j, r, c);
                        Tm[s][c] = v;
                        tm[(c * N + r) * GPW + grp] = make_double2(v.x, -v.y);   // (t^dag)[c][r]
                    }
                }
                __syncwarp();
                rows_times_smem<N, RP, GPW>(Tm, gm, grp, X);      // X = t g
                rows_times_smem<N, RP, GPW>(X, tm, grp, Tm);      // (reuse) = X t^dag
#pragma unroll
                for (int s = 0; s < RP; ++s)
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        A[s][c].x -= Tm[s][c].x;
                        A[s][c].y -= Tm[s][c].y;
                    }
                rows_times_smem<N, RP, GPW>(W, tm, grp, X);       // W t^dag
#pragma unroll
                for (int s = 0; s < RP; ++s)
#pragma unroll
                    for (int c = 0; c < N; ++c) W[s][c] = X[s][c];
            }
        }

        // lead self-energies on the end blocks (wfm/__init__.py:246-247)
        if (j == M - 1 || j == 0) {
            if (LEAD == LEAD_CLOSED && HOP == HOP_SCALAR) {
#pragma unroll
                for (int s = 0; s < RP; ++s) {
                    const int r = row0 + s;
                    const double2 sg = (j == 0) ? sigLd[s] : sigRd[s];   // M >= 2: distinct iterations
#pragma unroll
                    for (int c = 0; c < N; ++c)
                        if (c == r) {
                            A[s][c].x -= sg.x;
                            A[s][c].y -= sg.y;
                        }
                }
            } else {
                subtract_sigma_rows(j == 0);
            }
        }

        // every lane has finished reading g_{j+1} from gm before it is overwritten
        __syncwarp();
        invert_to_smem<N, LP, PIVX>(A, piv, gm, grp, row0, p.singular);

        if (j == M - 1) {
            // W_{M-1} = g_{M-1}; padding rows start (and stay) at zero
#pragma unroll
            for (int s = 0; s < RP; ++s) {
                const int r = row0 + s;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    double2 g = gm[(r * N + c) * GPW + grp];
                    if (!(full || r < n)) g = make_double2(0.0, 0.0);
                    W[s][c] = g;
                }
            }
        } else {
            rows_times_smem<N, RP, GPW>(W, gm, grp, A);   // (W t^dag) g_j
            if (HOP == HOP_SCALAR) {
                const double2 tc = make_double2(tj.x, -tj.y);
#pragma unroll
                for (int s = 0; s < RP; ++s)
#pragma unroll
                    for (int c = 0; c < N; ++c) W[s][c] = cmul2(A[s][c], tc);
            } else {
#pragma unroll
                for (int s = 0; s < RP; ++s)
#pragma unroll
                    for (int c = 0; c < N; ++c) W[s][c] = A[s][c];
            }
        }
    }

    // ------------------------------------------------------------------ epilogue (K3)
    // G_00 = g_0 is in gm, G_{M-1,0} = W.  Formulas of wfm/__init__.py:108-130.
    __syncwarp();
#pragma unroll
    for (int s = 0; s < RP; ++s) {
        const int r = row0 + s;
#pragma unroll
        for (int c = 0; c < N; ++c) A[s][c] = gm[(r * N + c) * GPW + grp];
    }
    double vL[N], vR[N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double2 v = vv[c * GPW + grp];
        vL[c] = v.x;
        vR[c] = v.y;
    }
    double sqL[RP], sqR[RP];   // sqrt of the velocities of own output channels
#pragma unroll
    for (int s = 0; s < RP; ++s) {
        const int r = row0 + s;
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (c == r) {
                a = vL[c];
                b = vR[c];
            }
        sqL[s] = sqrt(a);
        sqR[s] = sqrt(b);
    }

    const int n_src = p.n_src;
    const long long obase = pt * (long long)n_src * n;

    double tsum = 0.0;   // sum of T over own channels and all sources
    auto emit = [&](int i, bool src_up, const double (&rr)[RP], const double (&tt)[RP]) {
        double sum = 0.0;
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            const int r = row0 + s;
            if (full || r < n) {
                sum += rr[s] + tt[s];
                tsum += tt[s];
                if (active && p.R) {
                    p.R[obase + (long long)i * n + r] = rr[s];
                    p.T[obase + (long long)i * n + r] = tt[s];
                }
            }
        }
        if (p.resid) {
#pragma unroll
            for (int off = LP / 2; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (active && li == 0) p.resid[pt * n_src + i] = sum - 1.0;
        }
        if (p.spin) emit_spin<RP, LP>(p, pt, i, src_up, active, li, row0, n, rr, tt);
    };

    if (p.sources == nullptr) {
        // all n one-hot sources: alpha = i (static column index).  With A = e_alpha:
        //   i_flux = sqrt(vL[alpha])                                   (:108)
        //   r = (i G00[s][alpha] vL[alpha] - delta) sqrt(vL[s]) / i_flux  (:116-117)
        //   t =  i G_{M-1,0}[s][alpha] vL[alpha]  sqrt(vR[s]) / i_flux    (:124-125)
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (full || a < n) {
                const double rflux = 1.0 / sqrt(vL[a]);
                double rr[RP], tt[RP];
#pragma unroll
                for (int s = 0; s < RP; ++s) {
                    const int r = row0 + s;
                    const double fr = sqL[s] * rflux, ft = sqR[s] * rflux;
                    const double xr = (-A[s][a].y * vL[a] - (r == a ? 1.0 : 0.0)) * fr;
                    const double xi = (A[s][a].x * vL[a]) * fr;
                    rr[s] = fma(xr, xr, xi * xi);
                    const double yr = (-W[s][a].y * vL[a]) * ft;
                    const double yi = (W[s][a].x * vL[a]) * ft;
                    tt[s] = fma(yr, yr, yi * yi);
                }
                emit(a, a < p.n_up, rr, tt);
            }
        }
    } else {
        for (int i = 0; i < n_src; ++i) {
            const double* Asrc = p.sources + (long long)i * n;
            double f2 = 0.0, wu = 0.0, wd = 0.0;   // wu, wd: incident weight per electron-spin sector (emit_spin)
            double ur[RP], ui[RP], wr[RP], wi[RP], mine[RP];
#pragma unroll
            for (int s = 0; s < RP; ++s) ur[s] = ui[s] = wr[s] = wi[s] = mine[s] = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                if (full || c < n) {
                    const double a = Asrc[c];
                    const double av = a * vL[c];                  // Ajsigma * v_L
                    f2 = fma(a, av, f2);                          // :108
                    if (c < p.n_up) wu = fma(a, a, wu); else wd = fma(a, a, wd);
#pragma unroll
                    for (int s = 0; s < RP; ++s) {
                        ur[s] = fma(A[s][c].x, av, ur[s]);
                        ui[s] = fma(A[s][c].y, av, ui[s]);
                        wr[s] = fma(W[s][c].x, av, wr[s]);
                        wi[s] = fma(W[s][c].y, av, wi[s]);
                        if (c == row0 + s) mine[s] = a;
                    }
                }
            }
            const double rflux = 1.0 / sqrt(f2);
            double rr[RP], tt[RP];
#pragma unroll
            for (int s = 0; s < RP; ++s) {
                const double fr = sqL[s] * rflux, ft = sqR[s] * rflux;
                const double xr = (-ui[s] - mine[s]) * fr;
                const double xi = ur[s] * fr;
                rr[s] = fma(xr, xr, xi * xi);
                const double yr = -wi[s] * ft;
                const double yi = wr[s] * ft;
                tt[s] = fma(yr, yr, yi * yi);
            }
            emit(i, wu >= wd, rr, tt);
        }
    }
    if (p.ttot || p.gather.n) {
#pragma unroll
        for (int off = LP / 2; off >= 1; off >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
        if (active && li == 0) store_total(p.ttot, p.gather, pt, tsum);
    }
}

// Debug / unit-test kernel: inverts a batch of n x n blocks with the same device routine.
template <int N, int LP>
__global__ void __launch_bounds__(128) invert_batch_small(const cd* in, cd* out, long long count, int n, int* singular) {
    constexpr int RP = N / LP;
    constexpr int GPW = 32 / LP;
    extern __shared__ double2 smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / LP, li = lane % LP, row0 = li * RP;
    double2* piv = smem_all + (size_t)warp * ((2 * N + N * N) * GPW);
    double2* gm = piv + 2 * N * GPW;
    long long b = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * GPW + grp;
    const bool active = b < count;
    if (!active) b = count - 1;
    const double2* src = reinterpret_cast<const double2*>(in) + b * n * n;
    double2 A[RP][N];
#pragma unroll
    for (int s = 0; s < RP; ++s)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const int r = row0 + s;
            A[s][c] = (r < n && c < n) ? src[r * n + c] : make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
    invert_to_smem<N, LP>(A, piv, gm, grp, row0, singular);
    double2* dst = reinterpret_cast<double2*>(out) + b * n * n;
#pragma unroll
    for (int s = 0; s < RP; ++s)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const int r = row0 + s;
            if (active && r < n && c < n) dst[r * n + c] = gm[(r * N + c) * GPW + grp];
        }
}

}  // namespace tx
