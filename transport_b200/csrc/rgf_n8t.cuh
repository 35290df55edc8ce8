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
//   SL  Sigma_L diagonal per point     EB  energies of the eight points       IQ  Hamiltonian / energy index per point
#pragma once
#include "rgf_n8.cuh"

namespace tx {

struct N8TCfg {
    static constexpr int PSTRIDE = 65;
    static constexpr int PER_WARP = 3 * 8 * PSTRIDE + 128 + 64 + 8 + 4;   // in double2
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

template <int LEAD, int MINB>
__global__ void __launch_bounds__(128, MINB) rgf_sweep_n8t(const SweepParams p, const int kmax, const int gexp2) {
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
    const bool full = (n == N);

    double2* wsm = smem_all + (size_t)warp * N8TCfg::PER_WARP;
    double2* XB = wsm;
    double2* YB = XB + 8 * N8TCfg::PSTRIDE;
    double2* WB = YB + 8 * N8TCfg::PSTRIDE;
    double2* PV = WB + 8 * N8TCfg::PSTRIDE;   // [128]
    double2* SL = PV + 128;                   // [8 points][8 channels]
    double2* EB = SL + 64;                    // [8]
    int* IQ = reinterpret_cast<int*>(EB + 8); // [0..7] Hamiltonian index, [8..15] energy index

    const long long pt0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * GPW;
    long long pt = pt0 + gid;
    const bool active = pt < p.n_pts;
    if (!active) pt = p.n_pts - 1;
    const int ham = (int)(pt / p.nE);
    const int e = (int)(pt - (long long)ham * p.nE);
    const cd E = p.E[e];
    const long long nn = (long long)n * n;
    const double2* hall = reinterpret_cast<const double2*>(p.h);
    const double2* tall = reinterpret_cast<const double2*>(p.tnn);
    const double2* hbase = hall + ham * p.h_stride;
    const double2* tbase = tall + ham * p.t_stride;
    // one Hamiltonian for the whole warp?  (then block fragments are loaded once per site, not once per point)
    const bool uniform = __all_sync(FULLM, ham == __shfl_sync(FULLM, ham, 0));
    const double2 zero2 = make_double2(0.0, 0.0);

    auto tocd = [](double2 v) -> cd { return mk(v.x, v.y); };

    // ------------------------------------------------------------------ prologue (Rw): leads of own channels
    double vLo[RP], vRo[RP];
    {
        double2 sigLd[RP], sigRd[RP];
#pragma unroll
        for (int s = 0; s < RP; ++s) {
            sigLd[s] = sigRd[s] = zero2;
            vLo[s] = vRo[s] = 0.0;
            const int r = row0 + s;
            if (full || r < n) {
                if (LEAD == LEAD_CLOSED) {
                    const cd tL = tocd(__ldg(tbase + (long long)r * n + r));
                    const cd tR = tocd(__ldg(tbase + (long long)(M - 2) * nn + r * n + r));
                    const cd gL = g_closed_channel(E, tocd(__ldg(hbase + (long long)r * n + r)), tL);                       // wfm/__init__.py:283,286
                    const cd gR = g_closed_channel(E, tocd(__ldg(hbase + (long long)(M - 1) * nn + r * n + r)), tR);        // :284,287
                    const cd sl = conj(tL) * (gL * tL);                                                                      // :301
                    const cd sr = tR * (gR * conj(tR));                                                                      // :302
                    sigLd[s] = make_double2(sl.x, sl.y);
                    sigRd[s] = make_double2(sr.x, sr.y);
                    vLo[s] = -2.0 * sl.y;                                                                                    // :91
                    vRo[s] = -2.0 * sr.y;
                } else {
                    const long long off = ham * p.sig_stride + (long long)e * nn + (long long)r * n + r;
                    vLo[s] = -2.0 * reinterpret_cast<const double2*>(p.sigL)[off].y;
                    vRo[s] = -2.0 * reinterpret_cast<const double2*>(p.sigR)[off].y;
                }
            }
            PV[gid * 8 + r] = sigRd[s];
            SL[gid * 8 + r] = sigLd[s];
        }
        if (tid == 0) {
            EB[gid] = make_double2(E.x, E.y);
            IQ[gid] = ham;
            IQ[8 + gid] = e;
        }
    }

    // ------------------------------------------------------------------ F-role constants
    const double m0 = (gid == 2 * tid) ? 1.0 : 0.0;        // diagonal masks of the two fragment elements
    const double m1 = (gid == 2 * tid + 1) ? 1.0 : 0.0;
    const bool in0 = full || (gid < n && 2 * tid < n);      // element inside the real n x n block?
    const bool in1 = full || (gid < n && 2 * tid + 1 < n);
    const int f0i = n8t_idx(gid, 2 * tid), f1i = n8t_idx(gid, 2 * tid + 1);          // own F slots
    const int x0i = n8t_idx(2 * tid, gid), x1i = n8t_idx(2 * tid + 1, gid);          // transposing F write
    const unsigned zmask = (unsigned)(p.n_pts >> 62);   // run-time zero (see n8_invert)

    // W_M = I
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        WB[q * N8TCfg::PSTRIDE + f0i] = make_double2(m0, 0.0);
        WB[q * N8TCfg::PSTRIDE + f1i] = make_double2(m1, 0.0);
    }
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
    auto hptr = [&](int q, int j) -> const double2* {
        return (uniform ? hbase : hall + (long long)IQ[q] * p.h_stride) + (long long)j * nn;
    };
    auto tval = [&](int q, int j) -> double2 {
        return __ldg((uniform ? tbase : tall + (long long)IQ[q] * p.t_stride) + (long long)j * nn);
    };
    auto sigptr = [&](const cd* tab, int q) -> const double2* {
        return reinterpret_cast<const double2*>(tab) + (long long)IQ[q] * p.sig_stride + (long long)IQ[8 + q] * nn;
    };
    auto expo = [](double v) -> int { return (int)((hi_abs(v) >> 20) & 0x7ffu) - 1023; };

    // growth of the chunk's recurrence: 2 log2(max|Y|) - log2(prod |t|^2) > gexp2 ?   (warp-uniform)
    auto grown = [&](const double2 (&Y)[GPW][2], int ke) -> bool {
        unsigned mx = 0u;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            mx = max(mx, max(hi_abs(Y[q][0].x), hi_abs(Y[q][0].y)));
            mx = max(mx, max(hi_abs(Y[q][1].x), hi_abs(Y[q][1].y)));
        }
        mx = __reduce_max_sync(FULLM, mx);
        return 2 * ((int)(mx >> 20) - 1023) - ke > gexp2;
    };

    // one site inside a chunk:  Q <- P Z_j - |t_j|^2 Q     (P = Y_{j+1}, Q = Y_{j+2} -> Y_j), 0 <= j <= M-2
    auto site_step = [&](const double2 (&P)[GPW][2], double2 (&Q)[GPW][2], int j, int& ke, double2& tau) -> bool {
        double2 H0 = zero2, H1 = zero2, tj = zero2;
        if (uniform) {
            load_frag(hbase + (long long)j * nn, H0, H1);
            tj = __ldg(tbase + (long long)j * nn);
        }
        double kmin = 1e300;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            double2 a0 = H0, a1 = H1, t = tj;
            if (!uniform) {
                load_frag(hptr(q, j), a0, a1);
                t = tval(q, j);
            }
            const double kap = fma(t.x, t.x, t.y * t.y);
            kmin = fmin(kmin, kap);
            const double2 Eq = EB[q];
            double2 b0 = make_double2(fma(m0, Eq.x, -a0.x), fma(m0, Eq.y, -a0.y));
            double2 b1 = make_double2(fma(m1, Eq.x, -a1.x), fma(m1, Eq.y, -a1.y));
            if (j == 0) {
                if (LEAD == LEAD_CLOSED) {
                    const double2 sg = SL[q * 8 + gid];
                    b0.x = fma(-m0, sg.x, b0.x);  b0.y = fma(-m0, sg.y, b0.y);
                    b1.x = fma(-m1, sg.x, b1.x);  b1.y = fma(-m1, sg.y, b1.y);
                } else {
                    double2 s0, s1;
                    load_frag(sigptr(p.sigL, q), s0, s1);
                    b0.x -= s0.x;  b0.y -= s0.y;
                    b1.x -= s1.x;  b1.y -= s1.y;
                }
            }
            if (!in0) b0 = make_double2(m0 * (1.0 + kap), 0.0);    // padding channels: D stays the identity there
            if (!in1) b1 = make_double2(m1 * (1.0 + kap), 0.0);
            double re[2] = {-kap * Q[q][0].x, -kap * Q[q][1].x};
            double im[2] = {-kap * Q[q][0].y, -kap * Q[q][1].y};
            cdmma(re, im, P[q][0], b0);
            cdmma(re, im, P[q][1], b1);
            Q[q][0] = make_double2(re[0], im[0]);
            Q[q][1] = make_double2(re[1], im[1]);
        }
        ke += expo(kmin);
        if (uniform) tau = cmul2(tau, make_double2(tj.x, -tj.y));
        return grown(Q, ke);
    };

    double2 Ya[GPW][2], Yb[GPW][2];
    bool first = true;
    int j = M - 1;
    while (j >= 0) {
        // -------------------------------------------------------------- chunk start at site b:
        //   Ya <- Z_b - |t_b|^2 g_{b+1}^T   (Ya holds g_{b+1}^T; first chunk: Z_{M-1} with Sigma_R),   Yb <- I
        const int b = j;
        int ke = 0;
        double2 tau = make_double2(1.0, 0.0);
        {
            const int jt = (b < M - 1) ? b : M - 2;
            double2 H0 = zero2, H1 = zero2, tj = zero2;
            if (uniform) {
                load_fragT(hbase + (long long)b * nn, H0, H1);
                tj = __ldg(tbase + (long long)jt * nn);
            }
            double kmin = 1e300;
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                double2 a0 = H0, a1 = H1, t = tj;
                if (!uniform) {
                    load_fragT(hptr(q, b), a0, a1);
                    t = tval(q, jt);
                }
                const double kscale = fma(t.x, t.x, t.y * t.y);
                kmin = fmin(kmin, kscale);
                const double kap = (b < M - 1) ? kscale : 0.0;
                const double2 Eq = EB[q];
                double2 z0 = make_double2(fma(m0, Eq.x, -a0.x), fma(m0, Eq.y, -a0.y));
                double2 z1 = make_double2(fma(m1, Eq.x, -a1.x), fma(m1, Eq.y, -a1.y));
                if (b == M - 1 || b == 0) {
                    if (LEAD == LEAD_CLOSED) {
                        const double2 sg = (b == 0) ? SL[q * 8 + gid] : PV[q * 8 + gid];
                        z0.x = fma(-m0, sg.x, z0.x);  z0.y = fma(-m0, sg.y, z0.y);
                        z1.x = fma(-m1, sg.x, z1.x);  z1.y = fma(-m1, sg.y, z1.y);
                    } else {
                        double2 s0, s1;
                        load_fragT(sigptr(b == 0 ? p.sigL : p.sigR, q), s0, s1);
                        z0.x -= s0.x;  z0.y -= s0.y;
                        z1.x -= s1.x;  z1.y -= s1.y;
                    }
                }
                if (!in0) z0 = make_double2(m0 * (1.0 + kap), 0.0);
                if (!in1) z1 = make_double2(m1 * (1.0 + kap), 0.0);
                if (first) {
                    Ya[q][0] = z0;
                    Ya[q][1] = z1;
                } else {
                    Ya[q][0] = make_double2(fma(-kap, Ya[q][0].x, z0.x), fma(-kap, Ya[q][0].y, z0.y));
                    Ya[q][1] = make_double2(fma(-kap, Ya[q][1].x, z1.x), fma(-kap, Ya[q][1].y, z1.y));
                }
                Yb[q][0] = make_double2(m0, 0.0);
                Yb[q][1] = make_double2(m1, 0.0);
            }
            ke = expo(kmin);
            if (uniform && b < M - 1) tau = make_double2(tj.x, -tj.y);
        }
        int k = 1, phase = 0;
        bool grow = grown(Ya, ke);
        while (j > 0 && k < kmax && !grow) {
            --j;
            ++k;
            if (phase == 0) grow = site_step(Ya, Yb, j, ke, tau);
            else grow = site_step(Yb, Ya, j, ke, tau);
            phase ^= 1;
        }
        const int a = j;

        // -------------------------------------------------------------- close the chunk at site a
        // D_a -> XB (slot (q; r; c) = D_a[r][c] = Y_a[c][r]),  Y_{a+1} -> YB stored transposed
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const double2 c0 = phase ? Yb[q][0] : Ya[q][0], c1 = phase ? Yb[q][1] : Ya[q][1];
            const double2 p0 = phase ? Ya[q][0] : Yb[q][0], p1 = phase ? Ya[q][1] : Yb[q][1];
            XB[q * N8TCfg::PSTRIDE + x0i] = c0;
            XB[q * N8TCfg::PSTRIDE + x1i] = c1;
            YB[q * N8TCfg::PSTRIDE + x0i] = p0;
            YB[q * N8TCfg::PSTRIDE + x1i] = p1;
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
        // W <- tau W X   and   g_a^T = X^T Y_{a+1}   (the X fragment serves as B of the first and A of the second)
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const double2* xq = XB + q * N8TCfg::PSTRIDE;
            double2* wq = WB + q * N8TCfg::PSTRIDE;
            const double2* yq = YB + q * N8TCfg::PSTRIDE;
            const double2 x0 = xq[f0i], x1 = xq[f1i];     // X[2t+s][g]
            const double2 w0 = wq[f0i], w1 = wq[f1i];     // W[g][2t+s]
            const double2 y0 = yq[f0i], y1 = yq[f1i];     // Y_{a+1}[2t+s][g]
            double wr[2] = {0.0, 0.0}, wi[2] = {0.0, 0.0};
            double gr[2] = {0.0, 0.0}, gi[2] = {0.0, 0.0};
            cdmma(wr, wi, w0, x0);
            cdmma(gr, gi, x0, y0);
            cdmma(wr, wi, w1, x1);
            cdmma(gr, gi, x1, y1);
            double2 tq = tau;
            if (!uniform) {
                tq = make_double2(1.0, 0.0);
                for (int i = a; i <= b && i < M - 1; ++i) {
                    const double2 t = tval(q, i);
                    tq = cmul2(tq, make_double2(t.x, -t.y));
                }
            }
            wq[f0i] = make_double2(fma(wr[0], tq.x, -(wi[0] * tq.y)), fma(wi[0], tq.x, wr[0] * tq.y));
            wq[f1i] = make_double2(fma(wr[1], tq.x, -(wi[1] * tq.y)), fma(wi[1], tq.x, wr[1] * tq.y));
            Ya[q][0] = make_double2(gr[0], gi[0]);
            Ya[q][1] = make_double2(gr[1], gi[1]);
        }
        __syncwarp();   // every fragment of X and Y_{a+1} has been read: XB / YB may be overwritten
        first = false;
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
    for (int s = 0; s < RP; ++s) {
        sqL[s] = sqrt(vLo[s]);
        sqR[s] = sqrt(vRo[s]);
        PV[(row0 + s) * GPW + gid] = make_double2(vLo[s], 1.0 / sqL[s]);
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
    auto emit = [&](int i, const double (&rr)[RP], const double (&tt)[RP]) {
        double sum = 0.0;
        if (full) {
            sum = (rr[0] + tt[0]) + (rr[1] + tt[1]);
            tsum += tt[0] + tt[1];
            if (active) {
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
                    if (active) {
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
                emit(a, rr, tt);
            }
        }
    } else {
        for (int i = 0; i < n_src; ++i) {
            const double* Asrc = p.sources + (long long)i * n;
            double f2 = 0.0;
            double ur[RP], ui[RP], wr[RP], wi[RP], mine[RP];
#pragma unroll
            for (int s = 0; s < RP; ++s) ur[s] = ui[s] = wr[s] = wi[s] = mine[s] = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                if (full || c < n) {
                    const double a = Asrc[c];
                    const double av = a * vL[c];
                    f2 = fma(a, av, f2);
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
            emit(i, rr, tt);
        }
    }
    if (p.ttot) {
        tsum += __shfl_xor_sync(FULLM, tsum, 2);
        tsum += __shfl_xor_sync(FULLM, tsum, 1);
        if (active && tid == 0) p.ttot[pt] = tsum;
    }
}

}  // namespace tx
