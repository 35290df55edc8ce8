// Telescoped sweep for 5 <= n <= 8 with scalar hopping (tnn[j] = t_j I): the hot kernel of BASELINE configs 2 and 3.
//
// The plain sweep (rgf_n8.cuh) inverts one 8x8 block per site; ncu shows that inverse (pivot rows travelling between
// the four lanes of a point) is what binds it, not the FP64 pipe.  Here the left-connected g_j are not formed inside a
// chunk of sites [a..b]:
//
//     D_{b+1} = I,  D_{b+2} := g_{b+1},    D_j = z_j D_{j+1} - |t_j|^2 D_{j+2},    z_j = E - h_j (- Sigma_{L,R})
//     g_j = D_{j+1} D_j^-1          =>     g_b g_{b-1} ... g_a = D_a^-1      (the product telescopes)
//     X = D_a^-1,      g_a = D_{a+1} X,      W_a = (prod_{j=a..b} conj t_j) W_{b+1} X
//
// so a chunk of k sites costs k-1 matrix products + 1 inverse + 2 products instead of k inverses + k products, and the
// k-1 recurrence products never leave the register file:  with Y_j = D_j^T the recurrence reads
//     Y_j = Y_{j+1} Z_j - |t_j|^2 Y_{j+2},   Z_j = z_j^T,
// the accumulator fragment of mma.sync.m8n8k4.f64 (lane (g,t) holds C[g][2t], C[g][2t+1]) IS a valid A fragment when
// the contraction index of the two k-steps is taken as k = 2t and k = 2t+1, and the matching B fragment
// B[k][g] = Z_j[2t+s][g] = z_j[g][2t+s] is 32 contiguous bytes of row g of h_j (read straight from L1/L2, shared by the
// eight points of the warp when they belong to one Hamiltonian).  Per point and site: 8 DMMA + 8 FP64 lane operations, no
// shared-memory traffic, no shuffles.  (tcgen05 has no f64 kind; DMMA is the FP64 tensor path of sm_100a.)
//
// Error growth: the recurrence is a transfer-matrix product, so rounding errors are amplified by the growth of D inside
// the chunk.  A chunk is closed after `kmax` sites or as soon as max|D_j| / prod|t| over the warp exceeds 2^gexp
// (default 2^6): evanescent regions fall back to short chunks (k = 1 is exactly the plain sweep), propagating regions
// run at full length.  oracle/rgf_oracle.py:rgf_blocks_telescoped is the numpy restatement used to choose the bound
// (same error level as the plain sweep against the dense reference inverse).
//
// Layouts.  "F": lane (g,t) holds elements [g][2t], [g][2t+1] of the matrices of ALL eight points of the warp
// (fragment layout; recurrence and products).  "Rw": lane (q,li) holds rows 2li, 2li+1 of the matrix of point q
// (Gauss-Jordan inverse, prologue, epilogue).  Shared memory per warp, 16-byte complex units, slot (q; r; c) at
// q*65 + r*8 + (c ^ r) -- conflict-free for F reads/writes, for Rw row reads and for the transposing F write:
//   XB  D_a (for the inverse), then X stored transposed, finally G_00        YB  Y_{a+1} stored transposed
//   WB  W                                                                     PV  pivot rows / Sigma_R / velocities
//   SL  E - Sigma_L per (point, channel)     EB  energies of the eight points
#pragma once
#include "rgf_n8.cuh"

namespace tx {

struct N8TCfg {
    static constexpr int PSTRIDE = 65;
    static constexpr int PER_WARP = 3 * 8 * PSTRIDE + 128 + 64 + 8;   // in double2
    static constexpr int PER_WARP_BYTES = PER_WARP * 16;
};

__device__ __forceinline__ int n8t_idx(int r, int c) { return r * 8 + (c ^ r); }

__device__ __forceinline__ unsigned hi_abs(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }

// acc(re, im) += A * B for one k-step; a = A[g][k(t)], b = B[k(t)][g]
__device__ __forceinline__ void cdmma(double (&re)[2], double (&im)[2], const double2 a, const double2 b) {
    dmma_8x8x4(re, a.x, b.x);
    dmma_8x8x4(im, a.x, b.y);
    dmma_8x8x4(re, -a.y, b.y);
    dmma_8x8x4(im, a.y, b.x);
}

template <int LEAD, int MINB, bool FULL, bool DIAG, bool GUARD>
__global__ void __launch_bounds__(128, MINB) rgf_sweep_n8t(const SweepParams p, const int kmax, const int gexp2,
                                                           const int warps_per_ham, const long long n_warps,
                                                           double2* __restrict__ restart_scratch) {
    constexpr int N = 8, RP = 2, GPW = 8;
    constexpr unsigned FULLM = 0xffffffffu;
    extern __shared__ double2 smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gid = lane >> 2;     // F: fragment row        Rw: own point
    const int tid = lane & 3;      // F: column pair         Rw: row pair
    const int row0 = tid * RP;
    const int n = p.n;
    const int M = p.M;
    constexpr bool full = FULL;    // n == 8 (no padding channels)

    double2* wsm = smem_all + (size_t)warp * N8TCfg::PER_WARP;
    double2* XB = wsm;
    double2* YB = XB + 8 * N8TCfg::PSTRIDE;
    double2* WB = YB + 8 * N8TCfg::PSTRIDE;
    double2* PV = WB + 8 * N8TCfg::PSTRIDE;   // [128]  pivot rows; before the first inverse: E - Sigma_R per (point, channel)
    double2* SL = PV + 128;                   // [8 points][8 channels]  E - Sigma_L
    double2* EB = SL + 64;                    // [8]  energies of the eight points

    // Two instantiations are launched back to back when the structure scan ran: the one with the diagonal-site
    // path serves batches where at least 1/8 of the interior sites are diagonal, the lean one the rest (the
    // decision is taken on the device, no host synchronisation; the grid is persistent, so the idle one is a
    // few hundred CTAs that return at once).
    if (p.class_count != nullptr && ((long long)(*p.class_count) * 8 >= p.class_sites) != DIAG) return;
    const long long nn = (long long)n * n;
    const double2 zero2 = make_double2(0.0, 0.0);
    auto tocd = [](double2 v) -> cd { return mk(v.x, v.y); };
    // F-role constants
    const double m0 = (gid == 2 * tid) ? 1.0 : 0.0;        // diagonal masks of the two fragment elements
    const double m1 = (gid == 2 * tid + 1) ? 1.0 : 0.0;
    const bool in0 = full || (gid < n && 2 * tid < n);      // element inside the real n x n block?
    const bool in1 = full || (gid < n && 2 * tid + 1 < n);
    const int f0i = n8t_idx(gid, 2 * tid), f1i = n8t_idx(gid, 2 * tid + 1);          // own F slots
    const int x0i = n8t_idx(2 * tid, gid), x1i = n8t_idx(2 * tid + 1, gid);          // transposing F write
    const unsigned zmask = (unsigned)(p.n_pts >> 62);   // run-time zero (see n8_invert)

    // Persistent grid: a warp owns eight consecutive energies of ONE Hamiltonian at a time (the energy axis is
    // padded to a multiple of eight per Hamiltonian), so the block fragments of a site are loaded once and shared by
    // the eight points; finished warps take the next item on their own (no CTA-wide tail).
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    // g_{b+1}^T of the chunk in flight, kept in an L2-resident scratch (8 KB per resident warp) for the rare restart
    // of a chunk by the conditioning guard below
    auto gsave = [&]() -> double2* {   // recomputed where it is used: not worth a live register pair
        return restart_scratch + ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * (GPW * 64) + gid * 8 + 2 * tid;
    };
    for (long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + warp; wg < n_warps; wg += wstride) {
    const int ham = (int)(wg / warps_per_ham);
    const int e_raw = (int)(wg - (long long)ham * warps_per_ham) * GPW + gid;
    const bool active = e_raw < p.nE;
    const int e = min(e_raw, p.nE - 1);
    const long long pt = (long long)ham * p.nE + e;
    const cd E = p.E[e];
    const double2* hbase = reinterpret_cast<const double2*>(p.h) + ham * p.h_stride;
    const double2* tbase = reinterpret_cast<const double2*>(p.tnn) + ham * p.t_stride;

    // ------------------------------------------------------------------ prologue (Rw): leads of own channels
    double vLo[RP], vRo[RP];
    {
        auto same2 = [](cd a, cd b) -> bool { return a.x == b.x && a.y == b.y; };
        cd eLp = mk(0.0, 0.0), tLp = mk(0.0, 0.0), gLp = mk(0.0, 0.0);
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            double2 sigLd = zero2, sigRd = zero2;
            vLo[s] = vRo[s] = 0.0;
            const int r = row0 + s;
            if (full || r < n) {
                if (LEAD == LEAD_CLOSED) {
                    const cd tL = tocd(__ldg(tbase + (long long)r * n + r));
                    const cd tR = tocd(__ldg(tbase + (long long)(M - 2) * nn + r * n + r));
                    const cd eL = tocd(__ldg(hbase + (long long)r * n + r));
                    const cd eR = tocd(__ldg(hbase + (long long)(M - 1) * nn + r * n + r));
                    // identical channels / identical leads (the usual case) are evaluated once
                    const bool sameLprev = (s > 0) && same2(eL, eLp) && same2(tL, tLp);
                    const cd gL = sameLprev ? gLp : g_closed_channel(E, eL, tL);                                            // wfm/__init__.py:283,286
                    const bool sameRL = same2(eR, eL) && same2(tR, tL);
                    const cd gR = sameRL ? gL : g_closed_channel(E, eR, tR);                                                // :284,287
                    eLp = eL;  tLp = tL;  gLp = gL;
                    const cd sl = conj(tL) * (gL * tL);                                                                      // :301
                    const cd sr = tR * (gR * conj(tR));                                                                      // :302
                    sigLd = make_double2(sl.x, sl.y);
                    sigRd = make_double2(sr.x, sr.y);
                    vLo[s] = -2.0 * sl.y;                                                                                    // :91
                    vRo[s] = -2.0 * sr.y;
                } else {
                    const long long off = ham * p.sig_stride + (long long)e * nn + (long long)r * n + r;
                    vLo[s] = -2.0 * reinterpret_cast<const double2*>(p.sigL)[off].y;
                    vRo[s] = -2.0 * reinterpret_cast<const double2*>(p.sigR)[off].y;
                }
            }
            // effective energies of the two end sites (closed leads: Sigma is diagonal; table mode: Sigma is
            // subtracted as a full block where the site is built, the tables here just hold E)
            PV[gid * 8 + r] = make_double2(E.x - sigRd.x, E.y - sigRd.y);
            SL[gid * 8 + r] = make_double2(E.x - sigLd.x, E.y - sigLd.y);
        }
        if (tid == 0) EB[gid] = make_double2(E.x, E.y);
    }
    // all energies real?  (every reference driver: then Im of the B fragment does not depend on the point)
    const bool realE = __all_sync(FULLM, E.y == 0.0);

    __syncwarp();

    // block fragment [g][2t+s] (contiguous) / transposed fragment [2t+s][g] of an n x n block
    auto load_frag = [&](const double2* blk, double2& a0, double2& a1) {
        const double2* b = blk + gid * n + 2 * tid;
        a0 = in0 ? __ldg(b) : zero2;
        a1 = in1 ? __ldg(b + 1) : zero2;
    };
    auto load_fragT = [&](const double2* blk, double2& a0, double2& a1) {
        a0 = in0 ? __ldg(blk + (2 * tid) * n + gid) : zero2;
        a1 = in1 ? __ldg(blk + (2 * tid + 1) * n + gid) : zero2;
    };
    auto sigptr = [&](const cd* tab, int q) -> const double2* {
        const int eq = min((e_raw - gid) + q, p.nE - 1);
        return reinterpret_cast<const double2*>(tab) + ham * p.sig_stride + (long long)eq * nn;
    };
    auto expo = [](double v) -> int { return (int)((hi_abs(v) >> 20) & 0x7ffu) - 1023; };

    // growth of the chunk's recurrence (warp-uniform): 2 log2 max|U| + log2(scale2) > gexp2 ?
    // U = Y / d,  scale2 = d^2 / prod |t|^2
    auto grown = [&](const double2 (&U)[GPW][2], double scale2) -> bool {
        unsigned mx = 0u;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            mx = max(mx, max(hi_abs(U[q][0].x), hi_abs(U[q][0].y)));
            mx = max(mx, max(hi_abs(U[q][1].x), hi_abs(U[q][1].y)));
        }
        mx = __reduce_max_sync(FULLM, mx);
        return 2 * ((int)(mx >> 20) - 1023) + expo(scale2) > gexp2;
    };

    // Site classes (structure scan of the launcher: 0 general, >= 1 diagonal on-site block) through a 32-site
    // window held in registers: lane l = site win_top - l.  Table-mode Sigma_L is a full block: site 0 is general there.
    const int* cls = p.site_class ? p.site_class + (long long)ham * p.class_stride : nullptr;
    int win_top = -1, win_val = 0;
    auto site_is_diag = [&](int i) -> bool {
        if (!DIAG || cls == nullptr) return false;
        if (LEAD == LEAD_TABLE && i == 0) return false;
        if (i > win_top || i < win_top - 31) {
            win_top = i;
            const int site = i - lane;
            win_val = (site >= 0) ? __ldg(cls + site) : 0;
        }
        return __shfl_sync(FULLM, win_val, win_top - i) >= 1;
    };

    // One site inside a chunk, in the scaled variables U_j = Y_j / d_j with d_j = -|t_j|^2 d_{j+2}:
    //     Q <- P (f_j Z_j) + Q,   f_j = d_{j+1} / d_j      (P = U_{j+1}, Q = U_{j+2} -> U_j),  1 <= j+1 <= M-1
    // so the accumulators start from Q itself (no multiplication); for |t| = 1 the factors are +-1 and the
    // arithmetic is bit-identical to the unscaled recurrence.  (H0, H1) = f_j * h_j fragment, fj = f_j.
    auto site_step = [&](const double2 (&P)[GPW][2], double2 (&Q)[GPW][2], int j, double fj, double2 H0, double2 H1) {
        const double fm0 = fj * m0, fm1 = fj * m1;
        double2 pad0 = zero2, pad1 = zero2;
        if (!full) {   // padding channels: z = 1 + |t|^2 keeps D = 1 there; U_pad = 1/d
            // f_j (1 + kap) with kap = -d_j / d_{j+2}: supplied by the caller through H on the diagonal (see below)
            pad0 = H0;
            pad1 = H1;
        }
        const bool endsite = (j == 0);
        const double2* ep = endsite ? SL + gid : EB;
        const int es = endsite ? 8 : 1;
        if (realE && !endsite) {
            // B = (b_s.x, -H_s.y): the imaginary part is common to the eight points.  The DMMAs are issued in four
            // rounds over four points (asm volatile keeps this order): dependent instructions sit 8 apart.
            const double hy0 = H0.y, hy1 = H1.y, ny0 = -H0.y, ny1 = -H1.y;
#pragma unroll
            for (int q4 = 0; q4 < GPW; q4 += 4) {
                double b0x[4], b1x[4], re[4][2], im[4][2];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q4 + u;
                    const double ex = ep[q * es].x;
                    b0x[u] = fma(fm0, ex, -H0.x);
                    b1x[u] = fma(fm1, ex, -H1.x);
                    if (!full) {
                        if (!in0) b0x[u] = pad0.x;
                        if (!in1) b1x[u] = pad1.x;
                    }
                    re[u][0] = Q[q][0].x;  re[u][1] = Q[q][1].x;
                    im[u][0] = Q[q][0].y;  im[u][1] = Q[q][1].y;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { dmma_8x8x4(re[u], P[q4 + u][0].x, b0x[u]);  dmma_8x8x4(im[u], P[q4 + u][0].y, b0x[u]); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { dmma_8x8x4(re[u], P[q4 + u][0].y, hy0);     dmma_8x8x4(im[u], P[q4 + u][0].x, ny0); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { dmma_8x8x4(re[u], P[q4 + u][1].x, b1x[u]);  dmma_8x8x4(im[u], P[q4 + u][1].y, b1x[u]); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { dmma_8x8x4(re[u], P[q4 + u][1].y, hy1);     dmma_8x8x4(im[u], P[q4 + u][1].x, ny1); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    Q[q4 + u][0] = make_double2(re[u][0], im[u][0]);
                    Q[q4 + u][1] = make_double2(re[u][1], im[u][1]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                const double2 Eq = ep[q * es];
                double2 b0 = make_double2(fma(fm0, Eq.x, -H0.x), fma(fm0, Eq.y, -H0.y));
                double2 b1 = make_double2(fma(fm1, Eq.x, -H1.x), fma(fm1, Eq.y, -H1.y));
                if (LEAD == LEAD_TABLE && endsite) {
                    double2 s0, s1;
                    load_frag(sigptr(p.sigL, q), s0, s1);
                    b0.x = fma(-fj, s0.x, b0.x);  b0.y = fma(-fj, s0.y, b0.y);
                    b1.x = fma(-fj, s1.x, b1.x);  b1.y = fma(-fj, s1.y, b1.y);
                }
                if (!full) {
                    if (!in0) b0 = make_double2(pad0.x, 0.0);
                    if (!in1) b1 = make_double2(pad1.x, 0.0);
                }
                double re[2] = {Q[q][0].x, Q[q][1].x};
                double im[2] = {Q[q][0].y, Q[q][1].y};
                cdmma(re, im, P[q][0], b0);
                cdmma(re, im, P[q][1], b1);
                Q[q][0] = make_double2(re[0], im[0]);
                Q[q][1] = make_double2(re[1], im[1]);
            }
        }
    };

    // The same site when h_j is DIAGONAL (clean spacer sites, barrier regions, lead cells -- most sites of the
    // reference's own chains): P (f_j Z_j) is a column scaling, so the step is two complex FMAs per point on
    // the FP64 pipe instead of eight DMMAs.   (D0, D1) = f_j * h_j[c][c] for the lane's columns c = 2t, 2t+1.
    auto diag_step = [&](const double2 (&P)[GPW][2], double2 (&Q)[GPW][2], int j, double fj, double2 D0, double2 D1,
                         double padz) {
        const bool endsite = (j == 0);
        const double2* ep = endsite ? SL + 2 * tid : EB;
        const int es = endsite ? 8 : 1;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const double2 E0 = ep[q * es];
            const double2 E1 = endsite ? ep[q * es + 1] : E0;
            double2 z0 = make_double2(fma(fj, E0.x, -D0.x), fma(fj, E0.y, -D0.y));
            double2 z1 = make_double2(fma(fj, E1.x, -D1.x), fma(fj, E1.y, -D1.y));
            if (!full) {
                if (2 * tid >= n) z0 = make_double2(padz, 0.0);
                if (2 * tid + 1 >= n) z1 = make_double2(padz, 0.0);
            }
            cfma2(Q[q][0], P[q][0], z0);
            cfma2(Q[q][1], P[q][1], z1);
        }
    };

    // Conditioning guard (GUARD: chains longer than one chunk).  In long, nearly closed chains (localised regime) the
    // right-connected g_a can have a pole next to the real axis: |g_a| ~ 1e6, and the next chunk, which starts from
    // z - |t|^2 g_a, undoes it with a cancellation that costs |t g_a| eps of relative accuracy (the plain recursion meets
    // this at every site; measured on BASELINE config 3: unitarity residual 4.8e-8, relative errors 1e-8 where
    // kappa(E) eps = 1e-11).  g_a is therefore formed BEFORE the (irreversible) update of W; if |t_{a-1} g_a| exceeds
    // 2^GMAXX for any of the warp's points the chunk is not closed at a: it is restarted from its first site (g_{b+1}
    // reloaded from an L2 scratch) and made one site longer, so the pole is never inverted; at most three times per
    // chunk.  oracle/rgf_oracle.py:rgf_blocks_telescoped(cond_max=...) restates the idea in numpy.
    constexpr int GMAXX = 8;
    int xlen = 0;   // 0: normal chunk; else (restarts << 16) | forced length of the chunk being redone
    double2 Ya[GPW][2], Yb[GPW][2];
    int j = M - 1;
    while (j >= 0) {
        // -------------------------------------------------------------- chunk start at site b:
        //   Ya <- Z_b - |t_b|^2 g_{b+1}^T   (Ya holds g_{b+1}^T; first chunk: Z_{M-1} with Sigma_R),   Yb <- I
        const int b = j;
        double2 tau = make_double2(1.0, 0.0);
        double S2;                                    // prod |t_i|^2 over the chunk's sites (growth normaliser)
        {
            const int jt = (b < M - 1) ? b : M - 2;
            double2 a0, a1;
            load_fragT(hbase + (long long)b * nn, a0, a1);
            const double2 tj = __ldg(tbase + (long long)jt * nn);
            S2 = fma(tj.x, tj.x, tj.y * tj.y);
            const double kap = (b < M - 1) ? S2 : 0.0;
            if (b < M - 1) tau = make_double2(tj.x, -tj.y);
            const bool endsite = (b == 0 || b == M - 1);
            const double2* ep = (b == 0) ? SL + gid : ((b == M - 1) ? PV + gid : EB);
            const int es = endsite ? 8 : 1;
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                const double2 Eq = ep[q * es];
                double2 z0 = make_double2(fma(m0, Eq.x, -a0.x), fma(m0, Eq.y, -a0.y));
                double2 z1 = make_double2(fma(m1, Eq.x, -a1.x), fma(m1, Eq.y, -a1.y));
                if (LEAD == LEAD_TABLE && endsite) {
                    double2 s0, s1;
                    load_fragT(sigptr(b == 0 ? p.sigL : p.sigR, q), s0, s1);
                    z0.x -= s0.x;  z0.y -= s0.y;
                    z1.x -= s1.x;  z1.y -= s1.y;
                }
                if (!full) {
                    if (!in0) z0 = make_double2(m0 * (1.0 + kap), 0.0);
                    if (!in1) z1 = make_double2(m1 * (1.0 + kap), 0.0);
                }
                if (b == M - 1) {
                    Ya[q][0] = z0;
                    Ya[q][1] = z1;
                } else {
                    Ya[q][0] = make_double2(fma(-kap, Ya[q][0].x, z0.x), fma(-kap, Ya[q][0].y, z0.y));
                    Ya[q][1] = make_double2(fma(-kap, Ya[q][1].x, z1.x), fma(-kap, Ya[q][1].y, z1.y));
                }
                Yb[q][0] = make_double2(m0, 0.0);
                Yb[q][1] = make_double2(m1, 0.0);
            }
        }
        int k = 1, phase = 0;
        // scaled recurrence: d_j = -|t_j|^2 d_{j+2}; both d and 1/d are tracked by multiplication (one MUFU
        // reciprocal of |t_j|^2 per site, no division)
        double dcur = 1.0, dprev = 1.0, rdcur = 1.0, rdprev = 1.0;   // d_j, d_{j+1} and reciprocals
        double rS2 = fast_rcp(S2);                                   // 1 / prod |t_i|^2
        // The growth test looks at the newest matrix before every further site.  (Letting it lag one site behind,
        // to overlap its integer reduction with the DMMAs, was measured: no gain, and in evanescent regions the
        // extra site of growth costs a factor 20 in the accuracy of 1e-24 tunnelling entries.)
        double nrm = rS2;                             // d_{j+1}^2 / prod |t|^2 for the matrix under test
        // next site's fragment and hopping, one site ahead of the tensor-core work
        double2 Hn0 = zero2, Hn1 = zero2, tn = zero2;
        bool diag_next = false;                       // is site j-1 diagonal?  (then Hn holds its diagonal entries)
        auto prefetch = [&](int i) {                  // site i >= 0
            diag_next = site_is_diag(i);
            if (diag_next) {
                const double2* blk = hbase + (long long)i * nn;
                Hn0 = (full || 2 * tid < n) ? __ldg(blk + (2 * tid) * (n + 1)) : zero2;
                Hn1 = (full || 2 * tid + 1 < n) ? __ldg(blk + (2 * tid + 1) * (n + 1)) : zero2;
            } else {
                load_frag(hbase + (long long)i * nn, Hn0, Hn1);
            }
            tn = __ldg(tbase + (long long)i * nn);
        };
        if (j > 0) prefetch(j - 1);
        // one more site: Q <- U_j from P = U_{j+1}; false if the chunk has to be closed at the current site
        auto advance = [&](const double2 (&P)[GPW][2], double2 (&Q)[GPW][2]) -> bool {
            if (!(j > 0 && k < (xlen ? (xlen & 0xffff) : kmax))) return false;
            const double2 tj = tn;
            const double kap = fma(tj.x, tj.x, tj.y * tj.y);
            if (!(kap > 0.0)) return false;           // decoupled site (or NaN): it starts its own chunk
            if (xlen == 0 && grown(P, nrm)) return false;   // P = U_{j+1} is the newest matrix of the chunk
            --j;
            ++k;
            const double rk = fast_rcp(kap);
            const double dj = -kap * dprev;           // d_j = -|t_j|^2 d_{j+2}
            const double rdj = -rk * rdprev;
            const double fj = dcur * rdj;             // f_j = d_{j+1} / d_j
            double2 H0 = make_double2(fj * Hn0.x, fj * Hn0.y);
            double2 H1 = make_double2(fj * Hn1.x, fj * Hn1.y);
            const bool diag_now = diag_next;
            if (!full && !diag_now) {                 // padding: B diagonal = f_j (1 + kap), handed over in H
                if (!in0) H0 = make_double2(m0 * fj * (1.0 + kap), 0.0);
                if (!in1) H1 = make_double2(m1 * fj * (1.0 + kap), 0.0);
            }
            if (j > 0) prefetch(j - 1);
            if (diag_now) diag_step(P, Q, j, fj, H0, H1, fj * (1.0 + kap));
            else site_step(P, Q, j, fj, H0, H1);
            phase ^= 1;
            dprev = dcur;
            dcur = dj;
            rdprev = rdcur;
            rdcur = rdj;
            rS2 *= rk;
            nrm = dj * dj * rS2;
            tau = cmul2(tau, make_double2(tj.x, -tj.y));
            return true;
        };
        for (;;) {                                    // two sites per round: the arrays swap roles without moves
            if (!advance(Ya, Yb)) break;
            if (!advance(Yb, Ya)) break;
        }
        const int a = j;

        // -------------------------------------------------------------- close the chunk at site a
        // D_a / d_a -> XB (slot (q; r; c) = U_a[c][r]),  U_{a+1} -> YB stored transposed
        if (phase == 0) {
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                XB[q * N8TCfg::PSTRIDE + x0i] = Ya[q][0];
                XB[q * N8TCfg::PSTRIDE + x1i] = Ya[q][1];
                YB[q * N8TCfg::PSTRIDE + x0i] = Yb[q][0];
                YB[q * N8TCfg::PSTRIDE + x1i] = Yb[q][1];
            }
        } else {
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                XB[q * N8TCfg::PSTRIDE + x0i] = Yb[q][0];
                XB[q * N8TCfg::PSTRIDE + x1i] = Yb[q][1];
                YB[q * N8TCfg::PSTRIDE + x0i] = Ya[q][0];
                YB[q * N8TCfg::PSTRIDE + x1i] = Ya[q][1];
            }
        }
        __syncwarp();
        {
            double2 A[RP][N];
            double2* xb = XB + gid * N8TCfg::PSTRIDE;
#pragma unroll
            for (int s = 0; s < RP; ++s)
#pragma unroll
                for (int c = 0; c < N; ++c) A[s][c] = xb[n8t_idx(row0 + s, c)];
            __syncwarp();
            n8_invert<0, 1>(A, PV, xb, gid, row0, p.singular, zmask);   // X[k][c] -> slot (c; k): stored transposed
        }
        // W <- tau W X   and   g_a^T = X^T Y_{a+1}   (the X fragment serves as B of the first and A of the second);
        // undoing the scaling:  X = Xs / d_a,  Y_{a+1} = d_{a+1} U_{a+1}
        {
            const double rda = rdcur;
            const double gsc = dprev * rda;
            const double2 tq = make_double2(tau.x * rda, tau.y * rda);
            if (GUARD) {
                // g_a^T = X^T Y_{a+1} first, into Ya (the register copies of U_a, U_{a+1} are dead after the close)
                unsigned gmx = 0u;
#pragma unroll
                for (int q = 0; q < GPW; ++q) {
                    const double2* xq = XB + q * N8TCfg::PSTRIDE;
                    const double2* yq = YB + q * N8TCfg::PSTRIDE;
                    const double2 x0 = xq[f0i], x1 = xq[f1i];     // Xs[2t+s][g]
                    const double2 y0 = yq[f0i], y1 = yq[f1i];     // U_{a+1}[2t+s][g]
                    double gr[2] = {0.0, 0.0}, gi[2] = {0.0, 0.0};
                    cdmma(gr, gi, x0, y0);
                    cdmma(gr, gi, x1, y1);
                    Ya[q][0] = make_double2(gr[0] * gsc, gi[0] * gsc);
                    Ya[q][1] = make_double2(gr[1] * gsc, gi[1] * gsc);
                    gmx = max(gmx, max(max(hi_abs(Ya[q][0].x), hi_abs(Ya[q][0].y)), max(hi_abs(Ya[q][1].x), hi_abs(Ya[q][1].y))));
                }
                if (a > 0 && (xlen >> 16) < 3) {
                    gmx = __reduce_max_sync(FULLM, gmx);
                    const double kapn = fma(tn.x, tn.x, tn.y * tn.y);       // |t_{a-1}|^2 (prefetched with site a-1)
                    if (kapn > 0.0 && 2 * ((int)(gmx >> 20) - 1023) + expo(kapn) > 2 * GMAXX) {
                        xlen = ((xlen >> 16) + 1) << 16 | (k + 1);
                        if (b < M - 1) {
                            const double2* gs = gsave();
#pragma unroll
                            for (int q = 0; q < GPW; ++q) {
                                Ya[q][0] = gs[q * 64];
                                Ya[q][1] = gs[q * 64 + 1];
                            }
                        }
                        __syncwarp();
                        j = b;
                        continue;                     // same chunk again, one site longer
                    }
                }
                xlen = 0;
                // W <- tau W X
#pragma unroll
                for (int q = 0; q < GPW; ++q) {
                    const double2* xq = XB + q * N8TCfg::PSTRIDE;
                    double2* wq = WB + q * N8TCfg::PSTRIDE;
                    if (b == M - 1) {
                        const double2 c0 = xq[x0i], c1 = xq[x1i];     // Xs[g][2t+s]
                        wq[f0i] = cmul2(c0, tq);
                        wq[f1i] = cmul2(c1, tq);
                    } else {
                        const double2 x0 = xq[f0i], x1 = xq[f1i];
                        const double2 w0 = wq[f0i], w1 = wq[f1i];     // W[g][2t+s]
                        double wr[2] = {0.0, 0.0}, wi[2] = {0.0, 0.0};
                        cdmma(wr, wi, w0, x0);
                        cdmma(wr, wi, w1, x1);
                        wq[f0i] = make_double2(fma(wr[0], tq.x, -(wi[0] * tq.y)), fma(wi[0], tq.x, wr[0] * tq.y));
                        wq[f1i] = make_double2(fma(wr[1], tq.x, -(wi[1] * tq.y)), fma(wi[1], tq.x, wr[1] * tq.y));
                    }
                }
            } else
            if (b == M - 1) {
                // first chunk: W_M = I, so W = tau X is a scaled copy (X[g][2t+s] sits at the transposing slots)
#pragma unroll
                for (int q = 0; q < GPW; ++q) {
                    const double2* xq = XB + q * N8TCfg::PSTRIDE;
                    double2* wq = WB + q * N8TCfg::PSTRIDE;
                    const double2* yq = YB + q * N8TCfg::PSTRIDE;
                    const double2 x0 = xq[f0i], x1 = xq[f1i];     // Xs[2t+s][g]
                    const double2 y0 = yq[f0i], y1 = yq[f1i];     // U_{a+1}[2t+s][g]
                    const double2 c0 = xq[x0i], c1 = xq[x1i];     // Xs[g][2t+s]
                    double gr[2] = {0.0, 0.0}, gi[2] = {0.0, 0.0};
                    cdmma(gr, gi, x0, y0);
                    cdmma(gr, gi, x1, y1);
                    wq[f0i] = cmul2(c0, tq);
                    wq[f1i] = cmul2(c1, tq);
                    Ya[q][0] = make_double2(gr[0] * gsc, gi[0] * gsc);
                    Ya[q][1] = make_double2(gr[1] * gsc, gi[1] * gsc);
                }
            } else {
#pragma unroll
                for (int q = 0; q < GPW; ++q) {
                    const double2* xq = XB + q * N8TCfg::PSTRIDE;
                    double2* wq = WB + q * N8TCfg::PSTRIDE;
                    const double2* yq = YB + q * N8TCfg::PSTRIDE;
                    const double2 x0 = xq[f0i], x1 = xq[f1i];     // Xs[2t+s][g]
                    const double2 w0 = wq[f0i], w1 = wq[f1i];     // W[g][2t+s]
                    const double2 y0 = yq[f0i], y1 = yq[f1i];     // U_{a+1}[2t+s][g]
                    double wr[2] = {0.0, 0.0}, wi[2] = {0.0, 0.0};
                    double gr[2] = {0.0, 0.0}, gi[2] = {0.0, 0.0};
                    cdmma(wr, wi, w0, x0);
                    cdmma(gr, gi, x0, y0);
                    cdmma(wr, wi, w1, x1);
                    cdmma(gr, gi, x1, y1);
                    wq[f0i] = make_double2(fma(wr[0], tq.x, -(wi[0] * tq.y)), fma(wi[0], tq.x, wr[0] * tq.y));
                    wq[f1i] = make_double2(fma(wr[1], tq.x, -(wi[1] * tq.y)), fma(wi[1], tq.x, wr[1] * tq.y));
                    Ya[q][0] = make_double2(gr[0] * gsc, gi[0] * gsc);
                    Ya[q][1] = make_double2(gr[1] * gsc, gi[1] * gsc);
                }
            }
        }
        __syncwarp();   // every fragment of X and Y_{a+1} has been read: XB / YB may be overwritten
        if (GUARD && a > 0) {    // g_a^T: the start of the next chunk, should the guard have to restart it
            double2* gs = gsave();
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                gs[q * 64] = Ya[q][0];
                gs[q * 64 + 1] = Ya[q][1];
            }
        }
        j = a - 1;
    }

    // G_00 = g_0: slot (q; r; c) = G_00[r][c] = g_0^T[c][r]
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        XB[q * N8TCfg::PSTRIDE + x0i] = Ya[q][0];
        XB[q * N8TCfg::PSTRIDE + x1i] = Ya[q][1];
    }

    // ------------------------------------------------------------------ epilogue (K3, Rw role)
    double sqL[RP], sqR[RP];
#pragma unroll
    for (int s = 0; s < RP; ++s) {   // equal velocities (the usual case) share one square root / reciprocal
        sqL[s] = (s > 0 && vLo[s] == vLo[0]) ? sqL[0] : sqrt(vLo[s]);
        sqR[s] = (vRo[s] == vLo[s]) ? sqL[s] : ((s > 0 && vRo[s] == vRo[0]) ? sqR[0] : sqrt(vRo[s]));
    }
    {
        const double r0 = 1.0 / sqL[0];
        const double r1 = (sqL[1] == sqL[0]) ? r0 : 1.0 / sqL[1];
        PV[(row0 + 0) * GPW + gid] = make_double2(vLo[0], r0);
        PV[(row0 + 1) * GPW + gid] = make_double2(vLo[1], r1);
    }
    __syncwarp();
    double vL[N], rsL[N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const double2 v = PV[c * GPW + gid];
        vL[c] = v.x;
        rsL[c] = v.y;
    }
    const double2* g00 = XB + gid * N8TCfg::PSTRIDE;
    const double2* gn0 = WB + gid * N8TCfg::PSTRIDE;

    const int n_src = p.n_src;
    const long long obase = pt * (long long)n_src * n;
    double tsum = 0.0;
    auto emit = [&](int i, bool src_up, const double (&rr)[RP], const double (&tt)[RP]) {
        double sum = 0.0;
        if (full) {
            sum = (rr[0] + tt[0]) + (rr[1] + tt[1]);
            tsum += tt[0] + tt[1];
            if (active && p.R) {
                *reinterpret_cast<double2*>(p.R + obase + (long long)i * N + row0) = make_double2(rr[0], rr[1]);
                *reinterpret_cast<double2*>(p.T + obase + (long long)i * N + row0) = make_double2(tt[0], tt[1]);
            }
        } else {
#pragma unroll
            for (int s = 0; s < RP; ++s) {
                const int r = row0 + s;
                if (r < n) {
                    sum += rr[s] + tt[s];
                    tsum += tt[s];
                    if (active && p.R) {
                        p.R[obase + (long long)i * n + r] = rr[s];
                        p.T[obase + (long long)i * n + r] = tt[s];
                    }
                }
            }
        }
        if (p.resid) {
            sum += __shfl_xor_sync(FULLM, sum, 2);
            sum += __shfl_xor_sync(FULLM, sum, 1);
            if (active && tid == 0) p.resid[pt * n_src + i] = sum - 1.0;
        }
        if (p.spin) emit_spin<RP, 4>(p, pt, i, src_up, active, tid, row0, n, rr, tt);
    };

    if (p.sources == nullptr) {
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (full || a < n) {
                const double rflux = rsL[a];                           // 1 / i_flux   (wfm/__init__.py:108)
                double rr[RP], tt[RP];
#pragma unroll
                for (int s = 0; s < RP; ++s) {
                    const int r = row0 + s;
                    const double fr = sqL[s] * rflux, ft = sqR[s] * rflux;
                    const double2 g = g00[n8t_idx(r, a)];     // G_00[r][a]
                    const double2 w = gn0[n8t_idx(r, a)];     // G_{M-1,0}[r][a]
                    const double xr = (-g.y * vL[a] - (r == a ? 1.0 : 0.0)) * fr;   // (:116-117)
                    const double xi = (g.x * vL[a]) * fr;
                    rr[s] = fma(xr, xr, xi * xi);
                    const double yr = (-w.y * vL[a]) * ft;                           // (:124-125)
                    const double yi = (w.x * vL[a]) * ft;
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
                    const double av = a * vL[c];
                    f2 = fma(a, av, f2);
                    if (c < p.n_up) wu = fma(a, a, wu); else wd = fma(a, a, wd);
#pragma unroll
                    for (int s = 0; s < RP; ++s) {
                        const double2 g = g00[n8t_idx(row0 + s, c)];
                        const double2 w = gn0[n8t_idx(row0 + s, c)];
                        ur[s] = fma(g.x, av, ur[s]);
                        ui[s] = fma(g.y, av, ui[s]);
                        wr[s] = fma(w.x, av, wr[s]);
                        wi[s] = fma(w.y, av, wi[s]);
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
        tsum += __shfl_xor_sync(FULLM, tsum, 2);
        tsum += __shfl_xor_sync(FULLM, tsum, 1);
        if (active && tid == 0) store_total(p.ttot, p.gather, pt, tsum);
    }
    __syncwarp();   // the epilogue's shared-memory reads are done before the next item's prologue writes
    }   // persistent loop over the warp's items
}

}  // namespace tx
