// Bardeen transfer-Hamiltonian path (BASELINE config 5): bound states of the
// left/right well Hamiltonians, Oppenheimer matrix elements with energy-window averaging, Fermi-weighted
// current and the Bardeen transmission estimate.  Reference: transport/bardeen/__init__.py
//     kernel_well :17-257, kernel_well_prime :400-532, matrix_element :835-857,
//     current :537-575 (nFD :1014-1020), Ts_bardeen :577-632.
//
// The reference diagonalises dense (S n)^2 matrices with np.linalg.eigh and keeps the states below an
// energy cutoff.  For the chains its drivers build (rbbarrier.py, rbstep.py, rbmenez_nosup_el.py) every
// per-channel well Hamiltonian is a real symmetric TRIDIAGONAL matrix, so here:
//   * eigenvalues below the cutoff: Sturm-count multisection, one warp per eigenvalue, the 32 lanes
//     evaluating 32 shifts of the bracket per pass (5 bits per pass instead of 1);
//   * eigenvectors: twisted factorisation (forward LDL^T and backward UDU^T pivots of T - lambda,
//     twist where |gamma| is smallest) — O(S) per state, and the exponentially small tails under the
//     barrier, which are all the matrix elements see, come out with high RELATIVE accuracy;
//   * matrix elements: one CTA per (problem, alpha, m, beta), y = Hdiff[beta,alpha] psi_m in shared
//     memory, dot products only with the right-well states inside the energy window.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

#include <cstring>
#include <mutex>

#include "../../include/transport_b200.h"
#include "tx_host.h"

using namespace tx;

namespace {

constexpr int EIG_WARPS = 4;

// Number of eigenvalues < x = number of sign changes in the sequence of leading principal minors of T - x,
//     p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2},      p_{-1} = 1,
// evaluated without divisions: one dependent DFMA per site (the pivot form q_i = p_i / p_{i-1} costs a dependent
// FP64 division, ~10x the pipe time).  de(i) returns (d_i, e_{i-1}^2) of the matrix scaled by an exact power of two
// that brings its Gershgorin norm below 1 (xs is scaled alike), so |p| grows by less than 3 per site; every 8 sites
// the pair (p_{i-1}, p_i) is renormalised by the exponent of its larger entry (an exact scaling that keeps every
// sign).  Signs are shifted into a bit history (one SHF per site) and the changes counted with one POPC per 8 sites.
// A zero minor inside the sequence contributes exactly one change together with its successor (p_{i+1} = -e^2
// p_{i-1}), as Sturm's rule demands; a zero LAST minor (x is an eigenvalue) counts as a change: "< x" becomes "<= x"
// on that set of measure zero, the same convention as the -pivmin rule of the pivot form.
template <typename FDE>
__device__ __forceinline__ int sturm_count(int S, double xs, FDE de) {
    double p0 = 1.0;
    double p1 = de(0).x - xs;
    if (p1 == 0.0) p1 = -0x1p-100;
    int cnt = p1 < 0.0;
    unsigned sgn = (unsigned)__double2hiint(p1) >> 31;          // bit 0 = sign of the newest minor
    int i = 1;
    for (; i + 8 <= S; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 v = de(i + j);
            const double u = v.y * p0;
            p0 = p1;
            p1 = fma(v.x - xs, p1, -u);
            sgn = __funnelshift_l((unsigned)__double2hiint(p1), sgn, 1);
        }
        cnt += __popc((sgn ^ (sgn >> 1)) & 0xffu);
        const int h = max(__double2hiint(p0) & 0x7ff00000, __double2hiint(p1) & 0x7ff00000);
        const double sc = __hiloint2double(0x7fe00000 - h, 0);  // 2^(1023 - exponent)
        p0 *= sc;
        p1 *= sc;
    }
    for (; i < S; ++i) {
        const double2 v = de(i);
        const double u = v.y * p0;
        p0 = p1;
        p1 = fma(v.x - xs, p1, -u);
        sgn = __funnelshift_l((unsigned)__double2hiint(p1), sgn, 1);
        cnt += (sgn ^ (sgn >> 1)) & 1u;
    }
    if (p1 == 0.0 && S > 1) cnt += !(((sgn >> 1) ^ sgn) & 1u);  // zero last minor: count it as a change (once)
    return cnt;
}

// The same count, also returning the value of the last minor p_{S-1}(x) = det(T - x) as mant * 2^expo (the exponents
// taken out by the renormalisations are summed): the lane-per-eigenvalue kernel interpolates on it.
template <typename FDE>
__device__ __forceinline__ int sturm_count_f(int S, double xs, FDE de, double& mant, int& expo) {
    double p0 = 1.0;
    double p1 = de(0).x - xs;
    if (p1 == 0.0) p1 = -0x1p-100;
    int cnt = p1 < 0.0;
    int E = 0;
    unsigned sgn = (unsigned)__double2hiint(p1) >> 31;
    int i = 1;
    for (; i + 8 <= S; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 v = de(i + j);
            const double u = v.y * p0;
            p0 = p1;
            p1 = fma(v.x - xs, p1, -u);
            sgn = __funnelshift_l((unsigned)__double2hiint(p1), sgn, 1);
        }
        cnt += __popc((sgn ^ (sgn >> 1)) & 0xffu);
        const int h = max(__double2hiint(p0) & 0x7ff00000, __double2hiint(p1) & 0x7ff00000);
        const double sc = __hiloint2double(0x7fe00000 - h, 0);
        p0 *= sc;
        p1 *= sc;
        E += (h >> 20) - 1023;
    }
    for (; i < S; ++i) {
        const double2 v = de(i);
        const double u = v.y * p0;
        p0 = p1;
        p1 = fma(v.x - xs, p1, -u);
        sgn = __funnelshift_l((unsigned)__double2hiint(p1), sgn, 1);
        cnt += (sgn ^ (sgn >> 1)) & 1u;
    }
    if (p1 == 0.0 && S > 1) cnt += !(((sgn >> 1) ^ sgn) & 1u);
    const int hi = __double2hiint(p1);
    const int eb = (hi >> 20) & 0x7ff;
    if (eb == 0 || eb == 0x7ff) {                                // zero / denormal / non-finite: no usable value
        mant = 0.0;
        expo = 0;
    } else {
        mant = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(p1));
        expo = E + eb - 1023;
    }
    return cnt;
}

// exact power of two s with s * tnorm in [0.5, 1)  (1 for a zero matrix)
__device__ __forceinline__ double pow2_scale(double tnorm) {
    if (!(tnorm > 0.0) || !isfinite(tnorm)) return 1.0;
    int ex;
    frexp(tnorm, &ex);
    return ldexp(1.0, -ex);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// n_states[mat] = #eig < cut_states[mat], n_below[mat] = #eig < cut_count[mat].  One warp per matrix, lanes 0/1
// take the two shifts (global-memory reads: this kernel is a few microseconds).
__global__ void tridiag_count_kernel(const double* __restrict__ d, const double* __restrict__ e, int S, int n_mat,
                                     const double* __restrict__ cut_states, const double* __restrict__ cut_count,
                                     int* __restrict__ n_states, int* __restrict__ n_below) {
    const int mat = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (mat >= n_mat) return;
    const double* dm = d + (size_t)mat * S;
    const double* em = e + (size_t)mat * (S - 1);
    double tnorm = 0.0;
    for (int i = lane; i < S; i += 32) {
        const double r = (i ? fabs(em[i - 1]) : 0.0) + (i < S - 1 ? fabs(em[i]) : 0.0);
        tnorm = fmax(tnorm, fabs(dm[i]) + r);
    }
    tnorm = warp_max(tnorm);
    const double sc = pow2_scale(tnorm);
    if (lane < 2) {
        double x = lane == 0 ? cut_states[mat] : cut_count[mat];
        x = fmin(fmax(x, -4.0 * tnorm), 4.0 * tnorm);            // +-inf cutoffs: all / none
        const int cnt = sturm_count(S, x * sc, [&](int i) {
            const double es = i ? em[i - 1] * sc : 0.0;
            return make_double2(dm[i] * sc, es * es);
        });
        if (lane == 0) n_states[mat] = cnt;
        else n_below[mat] = cnt;
    }
}

// One warp per (matrix, k): eigenvalue k (ascending) and its eigenvector.  grid = (ceil(max_states /
// EIG_WARPS), n_mat).  Dynamic shared memory: (d, e^2)[S] interleaved, e[S].
__global__ void __launch_bounds__(EIG_WARPS * 32)
tridiag_eig_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, int S,
                   const double* __restrict__ cut_states, const int* __restrict__ n_states, int max_states,
                   double* __restrict__ evals) {
    extern __shared__ double2 sm2[];
    double2* de = sm2;                                  // (d_i, e_{i-1}^2), scaled
    double* e = reinterpret_cast<double*>(sm2 + S);     // e_i, scaled
    const int mat = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * EIG_WARPS + warp;
    int nb = n_states[mat];
    if (nb > max_states) nb = max_states;
    if (blockIdx.x * EIG_WARPS >= nb) return;                   // whole CTA idle
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        de[i].x = d_g[(size_t)mat * S + i];
        e[i] = i < S - 1 ? e_g[(size_t)mat * (S - 1) + i] : 0.0;
    }
    __syncthreads();
    // Gershgorin bounds (every warp computes the same numbers), then the matrix is rescaled in place by the exact
    // power of two that brings its norm below 1: eigenvectors are unchanged, eigenvalues scale exactly
    double glo = DBL_MAX, ghi = -DBL_MAX;
    for (int i = lane; i < S; i += 32) {
        const double r = (i ? fabs(e[i - 1]) : 0.0) + fabs(e[i]);
        glo = fmin(glo, de[i].x - r);
        ghi = fmax(ghi, de[i].x + r);
    }
    glo = warp_min(glo);
    ghi = warp_max(ghi);
    const double sc = pow2_scale(fmax(fabs(glo), fabs(ghi)));
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double es = e[i] * sc;
        e[i] = es;
        de[i].x *= sc;
        if (i + 1 < S) de[i + 1].y = es * es;
        if (i == 0) de[0].y = 0.0;
    }
    __syncthreads();
    if (k >= nb) return;
    const double pivmin = DBL_MIN * 4.0;
    const double tnorm = fmax(fabs(glo), fabs(ghi)) * sc;        // in [0.5, 1)
    glo = glo * sc - 2.0 * DBL_EPSILON * S - 2.0 * pivmin;

    // ---- multisection (scaled units): count(lo) <= k < count(hi)
    double lo = glo, hi = fmin(fmax(cut_states[mat], -4.0 * tnorm / sc), 4.0 * tnorm / sc) * sc;
    for (int pass = 0; pass < 40; ++pass) {
        const double x = lo + (hi - lo) * ((lane + 1) * (1.0 / 33.0));
        const int c = sturm_count(S, x, [&](int i) { return de[i]; });
        const unsigned above = __ballot_sync(0xffffffffu, c > k);
        double nlo = lo, nhi = hi;
        if (above == 0u) {
            nlo = __shfl_sync(0xffffffffu, x, 31);
        } else {
            const int first = __ffs(above) - 1;
            nhi = __shfl_sync(0xffffffffu, x, first);
            const double below = __shfl_sync(0xffffffffu, x, first > 0 ? first - 1 : 0);
            if (first > 0) nlo = below;
        }
        lo = fmax(lo, nlo);
        hi = fmin(hi, nhi);
        if (hi - lo <= DBL_EPSILON * (fabs(lo) + fabs(hi)) + DBL_EPSILON * 1e-3 * tnorm + 2.0 * pivmin) break;
    }
    const double lam = 0.5 * (lo + hi);
    if (lane == 0) evals[(size_t)mat * max_states + k] = lam / sc;
}

// ---- lane-per-state path (large batches) -------------------------------------------------------------------
// With thousands of states in flight a warp per state wastes the machine: multisection spends 32 Sturm counts per
// 5 bits.  Here every LANE owns one eigenvalue of the CTA's matrix and bisects it alone (1 bit per count, 6x fewer
// FP64 operations per eigenvalue and 32 eigenvalues per instruction), helped by interpolation on the last minor once
// its bracket is isolated; all lanes read the same (d_i, e_i^2) word of shared memory (broadcast).  A CTA of LP_THREADS lanes covers LP_THREADS consecutive eigenvalues of one matrix.
constexpr int LP_THREADS = 128;

__device__ __forceinline__ double stage_scaled(const double* __restrict__ d_g, const double* __restrict__ e_g, int S,
                                               double2* de, double* e, double& glo_out, double& tnorm_out) {
    __shared__ double s_lo[LP_THREADS / 32], s_hi[LP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        de[i].x = d_g[i];
        e[i] = i < S - 1 ? e_g[i] : 0.0;
    }
    __syncthreads();
    double glo = DBL_MAX, ghi = -DBL_MAX;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double r = (i ? fabs(e[i - 1]) : 0.0) + fabs(e[i]);
        glo = fmin(glo, de[i].x - r);
        ghi = fmax(ghi, de[i].x + r);
    }
    glo = warp_min(glo);
    ghi = warp_max(ghi);
    if (lane == 0) {
        s_lo[warp] = glo;
        s_hi[warp] = ghi;
    }
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        glo = fmin(glo, s_lo[w]);
        ghi = fmax(ghi, s_hi[w]);
    }
    const double sc = pow2_scale(fmax(fabs(glo), fabs(ghi)));
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double es = e[i] * sc;
        e[i] = es;
        de[i].x *= sc;
        if (i + 1 < S) de[i + 1].y = es * es;
        if (i == 0) de[0].y = 0.0;
    }
    __syncthreads();
    tnorm_out = fmax(fabs(glo), fabs(ghi)) * sc;
    glo_out = glo * sc;
    return sc;
}

// Same staging for the eigenvector kernel, which wants the off-diagonal itself (sign) next to the diagonal:
// de[i] = (d_i, e_{i-1}) scaled (e_{-1} = 0); 16 bytes per site instead of 24 leaves room for more work rows.
__device__ __forceinline__ double stage_scaled_de(const double* __restrict__ d_g, const double* __restrict__ e_g, int S,
                                                  double2* de) {
    __shared__ double s_hi[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S; i += blockDim.x) de[i] = make_double2(d_g[i], i ? e_g[i - 1] : 0.0);
    __syncthreads();
    double tn = 0.0;
    for (int i = threadIdx.x; i < S; i += blockDim.x)
        tn = fmax(tn, fabs(de[i].x) + fabs(de[i].y) + (i + 1 < S ? fabs(de[i + 1].y) : 0.0));
    tn = warp_max(tn);
    if (lane == 0) s_hi[warp] = tn;
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tn = fmax(tn, s_hi[w]);
    const double sc = pow2_scale(tn);
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        de[i].x *= sc;
        de[i].y *= sc;
    }
    __syncthreads();
    return sc;
}

__global__ void __launch_bounds__(LP_THREADS)
tridiag_eigval_lane_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, int S,
                           const double* __restrict__ cut_states, const int* __restrict__ n_states, int max_states,
                           double* __restrict__ evals) {
    extern __shared__ double2 sm2[];
    double2* de = sm2;
    double* e = reinterpret_cast<double*>(sm2 + S);
    const int mat = blockIdx.y;
    const int k = blockIdx.x * LP_THREADS + threadIdx.x;
    const int nb = min(n_states[mat], max_states);
    if (blockIdx.x * LP_THREADS >= nb) return;
    double glo, tnorm;
    const double sc = stage_scaled(d_g + (size_t)mat * S, e_g + (size_t)mat * (S - 1), S, de, e, glo, tnorm);
    const double pivmin = DBL_MIN * 4.0;
    double lo = glo - 2.0 * DBL_EPSILON * S - 2.0 * pivmin;
    double hi = fmin(fmax(cut_states[mat], -4.0 * tnorm / sc), 4.0 * tnorm / sc) * sc;
    int c_lo = 0, c_hi = n_states[mat];
    double m_lo = 0.0, m_hi = 0.0;
    int E_lo = 0, E_hi = 0, side = 0;
    // First pass, cooperative: the CTA's 128 threads probe 128 equispaced shifts of [lo, hi]; every eigenvalue then
    // starts from the probe interval that contains it (7 bisections' worth for the price of one count, and most
    // intervals already isolate their eigenvalue, so interpolation starts at once).
    {
        __shared__ double s_px[LP_THREADS], s_pm[LP_THREADS];
        __shared__ int s_pc[LP_THREADS], s_pE[LP_THREADS];
        const double x = lo + (hi - lo) * ((threadIdx.x + 1) * (1.0 / (LP_THREADS + 1)));
        double m;
        int E;
        s_pc[threadIdx.x] = sturm_count_f(S, x, [&](int i) { return de[i]; }, m, E);
        s_px[threadIdx.x] = x;
        s_pm[threadIdx.x] = m;
        s_pE[threadIdx.x] = E;
        __syncthreads();
        if (k >= nb) return;
        int t = 0;
        while (t < LP_THREADS && s_pc[t] <= k) ++t;              // first probe above eigenvalue k
        if (t < LP_THREADS) {
            hi = s_px[t]; c_hi = s_pc[t]; m_hi = s_pm[t]; E_hi = s_pE[t];
        }
        if (t > 0) {
            lo = s_px[t - 1]; c_lo = s_pc[t - 1]; m_lo = s_pm[t - 1]; E_lo = s_pE[t - 1];
        }
    }
    // Bracketing by Sturm counts (rigorous), the next shift by interpolation once the bracket isolates the eigenvalue:
    // count(hi) - count(lo) == 1 means det(T - x) changes sign exactly once inside, and the Illinois variant of
    // regula falsi on the last minor converges superlinearly (21-33 counts per eigenvalue instead of 53 bisections).
    // Every fourth step is a bisection, so the bracket halves at least that often whatever the minor's rounding does.
    for (int it = 0; it < 1100; ++it) {                          // c_lo = count(lo) <= k < count(hi) = c_hi
        if (hi - lo <= DBL_EPSILON * (fabs(lo) + fabs(hi)) + DBL_EPSILON * 1e-3 * tnorm + 2.0 * pivmin) break;
        const double mid = lo + 0.5 * (hi - lo);
        if (!(mid > lo && mid < hi)) break;
        double x = mid;
        if (c_hi - c_lo == 1 && m_lo != 0.0 && m_hi != 0.0 && (it & 3) != 3) {
            const int dE = max(-1000, min(1000, E_hi - E_lo));
            const double ratio = ldexp(m_hi / m_lo, dE);         // f(hi) / f(lo) < 0
            if (ratio < 0.0) {
                // The estimate is clamped into the bracket, not rejected, when it rounds onto an end (|f| at that end
                // is ~0: the end IS the root to working precision) — the probe just inside that end then closes the
                // bracket at once, where rejecting it would leave one-sided bisection (43 counts instead of 20).
                // Then step over the estimate, away from the nearer end, by 2 % of the distance to it (at least half
                // the tolerance): regula falsi alone closes in from one side and leaves the other end behind.
                const double xi = fmin(fmax(lo + (hi - lo) / (1.0 - ratio), lo), hi);
                const double tol = DBL_EPSILON * (fabs(lo) + fabs(hi)) + DBL_EPSILON * 1e-3 * tnorm + 2.0 * pivmin;
                const double dl = xi - lo, dh = hi - xi;
                const double xs = dl < dh ? fmin(xi + fmax(0.02 * dl, 0.5 * tol), mid) : fmax(xi - fmax(0.02 * dh, 0.5 * tol), mid);
                if (xs > lo && xs < hi) x = xs;
            }
        }
        double m;
        int E;
        const int c = sturm_count_f(S, x, [&](int i) { return de[i]; }, m, E);
        if (c > k) {
            hi = x; c_hi = c; m_hi = m; E_hi = E;
            if (side > 0) E_lo -= 1;                             // Illinois: the retained end's value is halved
            side = 1;
        } else {
            lo = x; c_lo = c; m_lo = m; E_lo = E;
            if (side < 0) E_hi -= 1;
            side = -1;
        }
    }
    evals[(size_t)mat * max_states + k] = 0.5 * (lo + hi) / sc;
}

// Eigenvectors: one warp per state, twisted factorisation (forward LDL^T pivots D+, backward UDU^T pivots D-, twist
// at the smallest |gamma|) with every site recurrence spread over the 32 lanes.
//   * Pivots.  D_i = p_i / p_{i-1} with p the minor recurrence, i.e. (p_i, p_{i-1}) = M_i (p_{i-1}, p_{i-2}),
//     M_i = [[d_i - lam, -e^2], [1, 0]].  Each lane multiplies the M_i of its contiguous piece of the chain (two
//     sequences, from (1,0) and (0,1)), a warp scan of the 2x2 piece matrices gives every piece its incoming pair,
//     and a second pass over the piece emits the pivots: critical path 2 S/32 sites instead of S.  Only ratios of
//     the pair matter, so pieces and prefixes are renormalised by exact powers of two at will.
//   * Vector.  z_r = 1, z_i = rho_i z_{i+1} below the twist and z_{i+1} = rho_i z_i above it: running products,
//     again a per-lane piece plus a warp scan of the piece totals (these are eigenvector components themselves).
// Everything lives in shared memory: D+ in the warp's work row, D- consumed on the fly (and recomputed once the
// twist is known), the ratios and finally the normalised vector in the same row, one coalesced copy to HBM.

__device__ __forceinline__ void renorm2(double& a, double& b) {
    const int h = max(__double2hiint(a) & 0x7ff00000, __double2hiint(b) & 0x7ff00000);
    const double rs = __hiloint2double(0x7fe00000 - h, 0);
    a *= rs;
    b *= rs;
}
__device__ __forceinline__ void renorm4(double& a, double& b, double& c, double& d) {
    const int h = max(max(__double2hiint(a) & 0x7ff00000, __double2hiint(b) & 0x7ff00000),
                      max(__double2hiint(c) & 0x7ff00000, __double2hiint(d) & 0x7ff00000));
    const double rs = __hiloint2double(0x7fe00000 - h, 0);
    a *= rs;
    b *= rs;
    c *= rs;
    d *= rs;
}

// Incoming pair (p_{j0-1}, p_{j0-2}) of this lane's piece [j0, j1) of one direction.  site(j) maps the j-th site of
// the direction to (d - lam, e^2 coupling it to the site before it in the direction; ignored for j = 0).
template <typename FS>
__device__ __forceinline__ void piece_incoming(int j0, int j1, int lane, FS site, double& p1, double& p0) {
    // piece matrix [[a, b], [c, d]]: (p_{j1-1}, p_{j1-2}) = [[a, b], [c, d]] (p_{j0-1}, p_{j0-2})
    double a = 1.0, b = 0.0, c = 0.0, d = 1.0;
    for (int j = j0; j < j1; ++j) {
        const double2 v = site(j);
        const double e2 = j ? v.y : 0.0;
        const double na = fma(v.x, a, -(e2 * c)), nb = fma(v.x, b, -(e2 * d));
        c = a;
        d = b;
        a = na;
        b = nb;
        if (((j - j0) & 7) == 7) renorm4(a, b, c, d);
    }
    renorm4(a, b, c, d);
    // inclusive scan: prefix_l = piece_l * piece_{l-1} * ... * piece_0
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double oa = __shfl_up_sync(0xffffffffu, a, off), ob = __shfl_up_sync(0xffffffffu, b, off);
        const double oc = __shfl_up_sync(0xffffffffu, c, off), od = __shfl_up_sync(0xffffffffu, d, off);
        if (lane >= off) {
            const double na = fma(a, oa, b * oc), nb = fma(a, ob, b * od);
            const double nc = fma(c, oa, d * oc), nd = fma(c, ob, d * od);
            a = na;
            b = nb;
            c = nc;
            d = nd;
            renorm4(a, b, c, d);
        }
    }
    // (exclusive prefix) applied to (p_{-1}, p_{-2}) = (1, 0): its first column
    p1 = __shfl_up_sync(0xffffffffu, a, 1);
    p0 = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) {
        p1 = 1.0;
        p0 = 0.0;
    }
}

// Pivots q_j = p_j / p_{j-1} of the piece from its incoming pair (the division is off the dependent chain).
template <typename FS, typename FO>
__device__ __forceinline__ void piece_pivots(int j0, int j1, double p1, double p0, double pivmin, FS site, FO out) {
    for (int j = j0; j < j1; ++j) {
        const double2 v = site(j);
        const double e2 = j ? v.y : 0.0;
        double pn = fma(v.x, p1, -(e2 * p0));
        if (pn == 0.0) pn = -p1 * 0x1p-100;
        double q = pn / p1;
        if (!(fabs(q) >= pivmin)) q = -pivmin;
        out(j, q);
        p0 = p1;
        p1 = pn;
        if (((j - j0) & 7) == 7) renorm2(p0, p1);
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
tridiag_eigvec_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, int S,
                      const int* __restrict__ n_states, int max_states, const double* __restrict__ evals,
                      double* __restrict__ evecs) {
    extern __shared__ double2 sm2[];
    double2* de = sm2;                                           // (d_i, e_{i-1}), scaled
    const int mat = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* w = reinterpret_cast<double*>(sm2 + S) + (size_t)warp * S;   // this warp's work row: D+ -> rho -> vector
    const int k = blockIdx.x * WARPS + warp;
    const int nb = min(n_states[mat], max_states);
    if (blockIdx.x * WARPS >= nb) return;
    const double sc = stage_scaled_de(d_g + (size_t)mat * S, e_g + (size_t)mat * (S - 1), S, de);
    if (k >= nb) return;
    auto e = [&](int i) { return de[i + 1].y; };                 // e_i couples sites i and i + 1
    const double pivmin = DBL_MIN * 4.0;
    const double lam = evals[(size_t)mat * max_states + k] * sc;
    // contiguous pieces of odd length: lane-strided shared-memory accesses then hit distinct banks
    const int per = ((S + 31) >> 5) | 1;
    const int j0 = min(S, lane * per), j1 = min(S, j0 + per);
    auto fwd = [&](int j) { const double2 v = de[j]; return make_double2(v.x - lam, v.y * v.y); };
    auto bwd = [&](int j) {
        const int i = S - 1 - j;
        const double ee = i + 1 < S ? de[i + 1].y : 0.0;
        return make_double2(de[i].x - lam, ee * ee);
    };
    double p1, p0;
    piece_incoming(j0, j1, lane, fwd, p1, p0);
    piece_pivots(j0, j1, p1, p0, pivmin, fwd, [&](int j, double q) { w[j] = q; });
    __syncwarp();
    // backward pivots on the fly: gamma_i = D+_i + D-_i - (d_i - lam), twist index r = argmin |gamma_i|
    double b1, b0;
    piece_incoming(j0, j1, lane, bwd, b1, b0);
    double best = DBL_MAX;
    int r = 0;
    piece_pivots(j0, j1, b1, b0, pivmin, bwd, [&](int j, double q) {
        const int i = S - 1 - j;
        const double g = fabs(w[i] + q - (de[i].x - lam));
        if (g < best || (g == best && i < r)) {
            best = g;
            r = i;
        }
    });
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int orr = __shfl_xor_sync(0xffffffffu, r, off);
        if (ob < best || (ob == best && orr < r)) {
            best = ob;
            r = orr;
        }
    }
    __syncwarp();
    // link ratios in place: below the twist rho_i = -e_i / D+_i (z_i = rho_i z_{i+1}); from the twist upwards
    // rho_i = -e_i / D-_{i+1} (z_{i+1} = rho_i z_i), the backward pivots being recomputed from the saved incoming pair
    for (int i = lane; i < r; i += 32) w[i] = -(e(i) / w[i]);
    piece_pivots(j0, j1, b1, b0, pivmin, bwd, [&](int j, double q) {
        const int i = S - 1 - j;
        if (i - 1 >= r) w[i - 1] = -(e(i - 1) / q);
    });
    __syncwarp();
    // running products away from the twist: per-lane pieces, a warp scan of the piece totals, and the piece sums of
    // squares give the norm before anything is written
    double start[2], ssq[2];
    const int Ls[2] = {r, S - 1 - r};
    int c0[2], c1[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const int L = Ls[dir];
        const int cper = ((L + 31) >> 5) | 1;
        c0[dir] = min(L, lane * cper);
        c1[dir] = min(L, c0[dir] + cper);
        double t = 1.0, q = 0.0;
        for (int j = c0[dir]; j < c1[dir]; ++j) {
            t *= dir == 0 ? w[r - 1 - j] : w[r + j];
            q = fma(t, t, q);
        }
        ssq[dir] = q;
        double in = t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, in, off);
            if (lane >= off) in *= o;
        }
        double st = __shfl_up_sync(0xffffffffu, in, 1);
        if (lane == 0) st = 1.0;
        start[dir] = st;                                         // component at the piece start (z_r = 1)
    }
    double nrm = fma(start[0] * start[0], ssq[0], start[1] * start[1] * ssq[1]);
    nrm = warp_sum(nrm) + 1.0;
    const double zs = 1.0 / sqrt(nrm);
    __syncwarp();
    // downwards: each lane reads a ratio and overwrites it with the normalised component (indices < r)
    {
        double v = start[0] * zs;
        for (int j = c0[0]; j < c1[0]; ++j) {
            v *= w[r - 1 - j];
            w[r - 1 - j] = v;
        }
    }
    // upwards: component r+j+1 lands where ratio r+j+1 is still to be read by the next step of the same lane or the
    // first step of the next lane, so every lane reads one ratio ahead and the first one before anybody writes
    {
        double v = start[1] * zs;
        double next = c0[1] < c1[1] ? w[r + c0[1]] : 0.0;
        __syncwarp();
        for (int j = c0[1]; j < c1[1]; ++j) {
            const double rho = next;
            if (j + 1 < c1[1]) next = w[r + j + 1];
            v *= rho;
            w[r + j + 1] = v;
        }
    }
    if (lane == 0) w[r] = zs;
    __syncwarp();
    double* z = evecs + ((size_t)mat * max_states + k) * S;
    for (int i = lane; i < S; i += 32) z[i] = w[i];
}

// Rows of y = Hdiff[beta, alpha] psi that can be nonzero: [lo, hi] per (problem, alpha, beta), computed once per call
// (Hdiff = Hsys - HL lives on the right well for kernel_well, on the few central sites for kernel_well_prime), so that
// the matrix-element CTAs touch only that part of the states.  support[(p*n*n + alpha*n + beta)*2 + {0,1}] = lo, hi
// (lo > hi: Hdiff[beta, alpha] vanishes).
__global__ void __launch_bounds__(128)
hdiff_support_kernel(const double* __restrict__ hd0, const double* __restrict__ hdU, const double* __restrict__ hdL, int n, int S,
                     int* __restrict__ support) {
    __shared__ int s_lohi[2];
    const int alpha = (blockIdx.x / n) % n, beta = blockIdx.x % n, p = blockIdx.x / (n * n);
    const size_t nn = (size_t)n * n, ba = (size_t)beta * n + alpha;
    const double* h0 = hd0 + (size_t)p * S * nn + ba;
    const double* hU = hdU + (size_t)p * (S - 1) * nn + ba;
    const double* hL = hdL + (size_t)p * (S - 1) * nn + ba;
    if (threadIdx.x == 0) {
        s_lohi[0] = S;
        s_lohi[1] = -1;
    }
    __syncthreads();
    int lo = S, hi = -1;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const bool nz = h0[(size_t)j * nn] != 0.0 || (j + 1 < S && hU[(size_t)j * nn] != 0.0) ||
                        (j > 0 && hL[(size_t)(j - 1) * nn] != 0.0);
        if (nz) {
            lo = min(lo, j);
            hi = max(hi, j);
        }
    }
    if (hi >= 0) {
        atomicMin(&s_lohi[0], lo);
        atomicMax(&s_lohi[1], hi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        support[2 * blockIdx.x] = s_lohi[0];
        support[2 * blockIdx.x + 1] = s_lohi[1];
    }
}

// Window-averaged Oppenheimer matrix elements.  grid = (max_states, n*n, P): CTA (m, alpha*n+beta, p).
//   Mavg[p][beta][m][alpha] = mean over k with |EL[alpha][m] - ER[beta][k]| < interval of
//                             <psiR[beta][k] | Hdiff[:, :, beta, alpha] | psiL[alpha][m]>
// (sign fixed to >= 0 before averaging when sign_fix, bardeen:141; nearest k when interval is NaN and the
// window is empty, :146-151; 0 when nothing qualifies).  Hdiff is block tridiagonal in the site index:
// hd0[p][S][n][n], hdU[p][S-1][n][n] = Hdiff[j, j+1], hdL[p][S-1][n][n] = Hdiff[j+1, j].
__global__ void __launch_bounds__(128)
bardeen_melem_kernel(const double* __restrict__ EL, const int* __restrict__ nL, const double* __restrict__ psiL,
                     const double* __restrict__ ER, const int* __restrict__ nR, const double* __restrict__ psiR,
                     const double* __restrict__ hd0, const double* __restrict__ hdU, const double* __restrict__ hdL,
                     const int* __restrict__ support, int n, int S, int max_states, double interval, int sign_fix,
                     double* __restrict__ Mavg) {
    extern __shared__ double y[];
    __shared__ double s_red[4];
    const int m = blockIdx.x, alpha = blockIdx.y / n, beta = blockIdx.y % n, p = blockIdx.z;
    const size_t chL = (size_t)p * n + alpha, chR = (size_t)p * n + beta;
    double* out = Mavg + (((size_t)p * n + beta) * max_states + m) * n + alpha;
    const int nl = min(nL[chL], max_states), nr = min(nR[chR], max_states);
    if (m >= nl) {
        if (threadIdx.x == 0) *out = 0.0;
        return;
    }
    const double Em = EL[chL * max_states + m];
    const double* Er = ER + chR * max_states;
    // kernel_well_prime (no sign fix) un-raggeds the right-well channels by appending copies of a channel's LAST bound
    // state up to the longest channel (:484-490); the copies take part in the window average, i.e. the last state
    // carries the weight 1 + (n_bound_right - nR[beta]).  kernel_well (sign fix) has no such padding.
    int pad = 0;
    if (!sign_fix) {
        int longest = 0;
        for (int b = 0; b < n; ++b) longest = max(longest, min(nR[(size_t)p * n + b], max_states));
        pad = longest - nr;
    }
    // window (uniform over the CTA)
    int cnt = 0, nearest = -1;
    double dmin = DBL_MAX;
    for (int k = 0; k < nr; ++k) {
        const double dE = fabs(Em - Er[k]);
        if (dE < interval) cnt += (k == nr - 1) ? 1 + pad : 1;
        if (dE < dmin) {
            dmin = dE;
            nearest = k;
        }
    }
    const bool use_nearest = (cnt == 0) && isnan(interval) && nearest >= 0;
    if (cnt == 0 && !use_nearest) {
        if (threadIdx.x == 0) *out = 0.0;
        return;
    }
    // y = Hdiff[beta, alpha] psi_m on the rows where it can be nonzero
    const int j0 = support[2 * ((p * n + alpha) * n + beta)], j1 = support[2 * ((p * n + alpha) * n + beta) + 1] + 1;
    if (j0 >= j1) {
        if (threadIdx.x == 0) *out = 0.0;
        return;
    }
    const double* psi = psiL + (chL * max_states + m) * S;
    const size_t nn = (size_t)n * n, ba = (size_t)beta * n + alpha;
    const double* h0 = hd0 + (size_t)p * S * nn + ba;
    const double* hU = hdU + (size_t)p * (S - 1) * nn + ba;
    const double* hL = hdL + (size_t)p * (S - 1) * nn + ba;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        double v = h0[(size_t)j * nn] * psi[j];
        if (j + 1 < S) v = fma(hU[(size_t)j * nn], psi[j + 1], v);
        if (j > 0) v = fma(hL[(size_t)(j - 1) * nn], psi[j - 1], v);
        y[j] = v;
    }
    __syncthreads();
    double sum = 0.0;
    for (int k = 0; k < nr; ++k) {
        const bool take = use_nearest ? (k == nearest) : (fabs(Em - Er[k]) < interval);
        if (!take) continue;
        const double* pr = psiR + (chR * max_states + k) * S;
        double acc = 0.0;
        for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) acc = fma(pr[j], y[j], acc);
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        double el = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        __syncthreads();
        if (sign_fix && el < 0.0) el = -el;
        sum += (!use_nearest && k == nr - 1) ? el * (double)(1 + pad) : el;
    }
    if (threadIdx.x == 0) *out = sum / (double)(use_nearest ? 1 : cnt);
}

// Iab[p][alpha][beta][i] = 2 pi sum_m stat(E[p][alpha][m], Vb[i]) M2[p][alpha][m][beta]   (bardeen:537-575)
// stat = fL (1 - fR) - fR (1 - fL), f = 1/(exp((eps - mu)/kBT) + 1), muL = muR + Vb.  One warp per output.
__global__ void bardeen_current_kernel(const double* __restrict__ E, const double* __restrict__ M2, int P, int n, int nb,
                                       const double* __restrict__ Vb, int nVb, double muR, double kBT,
                                       double* __restrict__ I) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)P * n * n * nVb;
    if (w >= total) return;
    const int i = (int)(w % nVb);
    const int beta = (int)((w / nVb) % n);
    const int alpha = (int)((w / ((long long)nVb * n)) % n);
    const int p = (int)(w / ((long long)nVb * n * n));
    const double muL = muR + Vb[i];
    const double* Ea = E + ((size_t)p * n + alpha) * nb;
    const double* Ma = M2 + ((size_t)p * n + alpha) * nb * n + beta;
    double acc = 0.0;
    for (int m = lane; m < nb; m += 32) {
        const double fL = 1.0 / (exp((Ea[m] - muL) / kBT) + 1.0);
        const double fR = 1.0 / (exp((Ea[m] - muR) / kBT) + 1.0);
        acc = fma(fL * (1.0 - fR) - fR * (1.0 - fL), Ma[(size_t)m * n], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) I[w] = 2.0 * 3.141592653589793 * acc;
}

// Tbmas[alpha][m][beta] = NL/(k tL[alpha]) NR/tR[beta] dos M2[alpha][m][beta]   (bardeen:577-632)
// k = Re arccos((E - VL[alpha]) / (-2 tL[alpha])), dos = Re 1/sqrt(1 - ((E - VR[beta]) / (2 tR[beta]))^2) with
// numpy's complex branches (the reference's Emas is complex): outside the band k = 0 or pi and dos = 0.
// flag[0] counts wavenumbers outside the band (:611 raises on any, for tL > 0), flag[2 + alpha*n + beta] the in-band
// densities of states: outside the band (x < -1 AND x > 1: numpy's 1/sqrt(1 - x^2) of a complex x with zero imaginary
// part has a NEGATIVE imaginary part on both sides) the reference's check (:624, max of the imaginary parts) fires
// only if NO m of that (alpha, beta) is in band; flag[1] is kept in the layout and stays 0.
__global__ void bardeen_ts_kernel(const double* __restrict__ E, const double* __restrict__ M2, int n, int nb,
                                  const double* __restrict__ tL, const double* __restrict__ VL,
                                  const double* __restrict__ tR, const double* __restrict__ VR, double NL, double NR,
                                  double* __restrict__ T, int* __restrict__ flag) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nb * n) return;
    const int beta = idx % n, m = (idx / n) % nb, alpha = idx / (n * nb);
    const double Em = E[alpha * nb + m];
    const double ck = (Em - VL[alpha]) / (-2.0 * tL[alpha]);
    double ka;
    if (fabs(ck) <= 1.0) {
        ka = acos(ck);
    } else {
        ka = ck > 0.0 ? 0.0 : 3.141592653589793;
        atomicAdd(flag, 1);
    }
    const double x = (Em - VR[beta]) / (2.0 * tR[beta]);
    const double arg = 1.0 - x * x;
    double dos = 0.0;
    if (arg >= 0.0) {
        dos = 1.0 / sqrt(arg);
        atomicAdd(flag + 2 + alpha * n + beta, 1);
    }
    T[idx] = NL / (ka * tL[alpha]) * NR / tR[beta] * dos * M2[idx];
}


// Device-side assembly of the three Hamiltonians of kernel_well / kernel_well_prime (bardeen/__init__.py:90-125,
// :455-500) from the REGION PARAMETERS of Hsysmat (:690-750), one thread per (geometry, site): regions
// inf | L | C | R | inf with on-site potentials per channel and scalar hoppings per channel; the centre is the chain
// HC (Hsys) / HCprime (both wells).  Nothing of size S travels over PCIe: a sweep over step heights or barrier
// parameters ships a few doubles per geometry (as affine_expand_kernel does for the T(E) sweeps).
//   well L: (Vinf, VL, HCprime, VRprime, Vinf)     well R: (Vinf, VLprime, HCprime, VR, Vinf)     sys: (Vinf, VL, HC, VR, Vinf)
// Outputs: per-channel tridiagonals dL, eL, dR, eR [P][n][S(-1)] and Hdiff = Hsys - HL as block tridiagonal
// hd0 [P][S][n][n], hdU / hdL [P][S-1][n][n].
struct RegionParams {
    const double *tinf, *tL, *tR, *Vinf;            // [n]
    const double *VL, *VLp, *VR, *VRp;              // [P or 1][n]
    long long v_stride;                             // n, or 0 when the potentials are shared by all geometries
    const double *hc0, *hcU, *hcL;                  // Hsys centre: [NC][n][n], [NC-1][n][n] (upper = [j][j+1], lower = [j+1][j])
    const double *hp0, *hpU, *hpL;                  // wells' centre (HCprime)
    long long hc_stride0, hc_stride1;               // per-geometry strides of the centre arrays (0 = shared)
    int Ninf, NL, NR, NC, n, P, S;
};

__global__ void bardeen_assemble_kernel(const RegionParams q, double* __restrict__ dL, double* __restrict__ eL,
                                        double* __restrict__ dR, double* __restrict__ eR, double* __restrict__ hd0,
                                        double* __restrict__ hdU, double* __restrict__ hdL) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)q.P * q.S) return;
    const int p = (int)(idx / q.S), j = (int)(idx - (long long)p * q.S);
    const int n = q.n, S = q.S;
    const int a = q.Ninf, b = a + q.NL, c = b + q.NC, d = c + q.NR;
    const double* VL = q.VL + p * q.v_stride;
    const double* VLp = q.VLp + p * q.v_stride;
    const double* VR = q.VR + p * q.v_stride;
    const double* VRp = q.VRp + p * q.v_stride;
    const double* hc0 = q.hc0 + p * q.hc_stride0;
    const double* hp0 = q.hp0 + p * q.hc_stride0;
    const double* hcU = q.hcU + p * q.hc_stride1;
    const double* hcL = q.hcL + p * q.hc_stride1;
    const double* hpU = q.hpU + p * q.hc_stride1;
    const double* hpL = q.hpL + p * q.hc_stride1;
    const bool centre = j >= b && j < c;
    const bool cbond = j >= b && j < c - 1;          // bond j -> j+1 inside the centre
    for (int ch = 0; ch < n; ++ch) {
        double vl, vr;                               // on-site energy in well L / well R
        if (j < a || j >= d) vl = vr = q.Vinf[ch];
        else if (j < b) { vl = VL[ch]; vr = VLp[ch]; }
        else if (j < c) vl = vr = hp0[((long long)(j - b) * n + ch) * n + ch];
        else { vl = VRp[ch]; vr = VR[ch]; }
        dL[((long long)p * n + ch) * S + j] = vl;
        dR[((long long)p * n + ch) * S + j] = vr;
        if (j < S - 1) {
            double t;
            if (j < a || j >= d - 1) t = -q.tinf[ch];
            else if (j < b) t = -q.tL[ch];
            else if (cbond) t = hpU[((long long)(j - b) * n + ch) * n + ch];
            else t = -q.tR[ch];
            eL[((long long)p * n + ch) * (S - 1) + j] = t;
            eR[((long long)p * n + ch) * (S - 1) + j] = t;
        }
    }
    // Hdiff = Hsys - HL: the right region's potential (VR - VRprime) and the centre (HC - HCprime)
    for (int r = 0; r < n; ++r)
        for (int cc = 0; cc < n; ++cc) {
            double v = 0.0;
            if (centre) v = hc0[((long long)(j - b) * n + r) * n + cc] - hp0[((long long)(j - b) * n + r) * n + cc];
            else if (j >= c && j < d && r == cc) v = VR[r] - VRp[r];
            hd0[(((long long)p * S + j) * n + r) * n + cc] = v;
            if (j < S - 1) {
                double u = 0.0, l = 0.0;
                if (cbond) {
                    u = hcU[((long long)(j - b) * n + r) * n + cc] - hpU[((long long)(j - b) * n + r) * n + cc];
                    l = hcL[((long long)(j - b) * n + r) * n + cc] - hpL[((long long)(j - b) * n + r) * n + cc];
                }
                hdU[(((long long)p * (S - 1) + j) * n + r) * n + cc] = u;
                hdL[(((long long)p * (S - 1) + j) * n + r) * n + cc] = l;
            }
        }
}

// y = x^2 on Mavg[p][beta][m][alpha]; with row_cut, rows m >= n_below[p][beta] are zeroed (the reference's
// Mnumus[:nu_cut] of kernel_well :181)
__global__ void square_kernel(const double* __restrict__ x, double* __restrict__ y, long long count, int ms, int n,
                              const int* __restrict__ n_below, int row_cut) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long row = i / n;               // (p*n + beta)*ms + m
    const int m = (int)(row % ms);
    const long long mat = row / ms;
    y[i] = (row_cut && m >= n_below[mat]) ? 0.0 : x[i] * x[i];
}

// C[i][j] = sum_k opA(A)[i][k] w[k] opB(B)[k][j] for small real matrices (basis changes of kernel_well :156-166):
// transA: A stored [K][M] (else [M][K]); transB: B stored [N][K] (else [K][N]); w may be null.  One thread per element.
__global__ void real_gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ w,
                                 double* __restrict__ C, int M, int N, int K, int transA, int transB) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= M || j >= N) return;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
        const double a = transA ? A[(size_t)k * M + i] : A[(size_t)i * K + k];
        const double b = transB ? B[(size_t)j * K + k] : B[(size_t)k * N + j];
        acc = fma(w ? a * w[k] : a, b, acc);
    }
    C[(size_t)i * N + j] = acc;
}

__global__ void max_int_kernel(const int* __restrict__ a, const int* __restrict__ b, int count, int* __restrict__ out) {
    int m = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) m = max(m, max(a[i], b[i]));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

struct DevMem {
    char* p = nullptr;
    ~DevMem() {
        if (p) cudaFree(p);
    }
};

int have_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    return TX_OK;
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

int g_eig_mode = 0;   // tx_set_option(3, .): 0 auto, 1 warp per state (multisection), 2 lane per state (bisection)

int launch_eig(const double* d, const double* e, int S, int n_mat, const double* cut_states, const int* n_states,
               int max_states, double* evals, double* evecs, double* scratch, cudaStream_t st) {
    const size_t smem = (size_t)3 * S * sizeof(double);
    if (smem > 200 * 1024) return fail(TX_ERR_UNSUPPORTED, "chain length S=%d exceeds the shared-memory staged limit (8533)", S);
    static bool attr_set = false;
    if (!attr_set) {
        TX_CUDA(cudaFuncSetAttribute(tridiag_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TX_CUDA(cudaFuncSetAttribute(tridiag_eigval_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TX_CUDA(cudaFuncSetAttribute(tridiag_eigvec_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TX_CUDA(cudaFuncSetAttribute(tridiag_eigvec_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TX_CUDA(cudaFuncSetAttribute(tridiag_eigvec_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    // A warp per eigenvalue has the shorter critical path (11 multisection passes against 21-33 interpolated counts)
    // and serves small calls; a lane per eigenvalue does ~10x less arithmetic per eigenvalue and wins as soon as
    // there are a few thousand states (measured on B200, both wells of 256 geometries: S = 451, 10 461 states:
    // 0.58 vs 0.43 ms; S = 1651, 42 225 states: 7.3 vs 2.3 ms).  Four interleaved shifts per lane (5-section) were
    // measured too and are slower (1.69 vs 1.44 ms per well): with ~1.3 warps per scheduler the doubly loaded
    // schedulers are FP64-pipe bound and 5-section does 1.7x the arithmetic.
    const bool lanes = g_eig_mode == 2 || (g_eig_mode == 0 && (long long)n_mat * max_states >= 2048);
    if (!lanes) {
        dim3 grid((max_states + EIG_WARPS - 1) / EIG_WARPS, n_mat);
        tridiag_eig_kernel<<<grid, EIG_WARPS * 32, smem, st>>>(d, e, S, cut_states, n_states, max_states, evals);
    } else {
        dim3 grid((max_states + LP_THREADS - 1) / LP_THREADS, n_mat);
        tridiag_eigval_lane_kernel<<<grid, LP_THREADS, smem, st>>>(d, e, S, cut_states, n_states, max_states, evals);
    }
    TX_CUDA(cudaGetLastError());
    count_launch();
    if (evecs) {
        // (d, e) and one work row per warp in shared memory: 16 S + 8 S per warp bytes.  Six warps per CTA while two
        // CTAs fit an SM (12 warps/SM), else four, else one (S <= 8533)
        (void)scratch;
        if ((size_t)64 * S <= 110 * 1024) {
            dim3 grid((max_states + 5) / 6, n_mat);
            tridiag_eigvec_kernel<6><<<grid, 192, (size_t)64 * S, st>>>(d, e, S, n_states, max_states, evals, evecs);
        } else if ((size_t)48 * S <= 200 * 1024) {
            dim3 grid((max_states + 3) / 4, n_mat);
            tridiag_eigvec_kernel<4><<<grid, 128, (size_t)48 * S, st>>>(d, e, S, n_states, max_states, evals, evecs);
        } else if ((size_t)24 * S <= 200 * 1024) {
            dim3 grid(max_states, n_mat);
            tridiag_eigvec_kernel<1><<<grid, 32, (size_t)24 * S, st>>>(d, e, S, n_states, max_states, evals, evecs);
        } else {
            return fail(TX_ERR_UNSUPPORTED, "chain length S=%d exceeds the eigenvector kernel's shared-memory limit (8533)", S);
        }
        TX_CUDA(cudaGetLastError());
        count_launch();
    }
    return TX_OK;
}

int launch_count(const double* d, const double* e, int S, int n_mat, const double* cut_states, const double* cut_count,
                 int* n_states, int* n_below, cudaStream_t st) {
    tridiag_count_kernel<<<(n_mat + 3) / 4, 128, 0, st>>>(d, e, S, n_mat, cut_states, cut_count, n_states, n_below);
    TX_CUDA(cudaGetLastError());
    count_launch();
    return TX_OK;
}

}  // namespace

namespace tx {
int set_eig_mode(int mode) {
    if (mode < 0 || mode > 2) return fail(TX_ERR_ARG, "eigen-solver mode %d outside 0..2", mode);
    g_eig_mode = mode;
    return TX_OK;
}
}  // namespace tx

extern "C" {

int tx_tridiag_count_below_dev(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                               const double* cut_count, int* n_states, int* n_below, void* stream) {
    if (!d || !e || !cut_states || !cut_count || !n_states || !n_below) return fail(TX_ERR_ARG, "null pointer argument");
    if (S < 2 || n_mat < 1) return fail(TX_ERR_ARG, "bad sizes S=%d n_mat=%d", S, n_mat);
    return launch_count(d, e, S, n_mat, cut_states, cut_count, n_states, n_below, reinterpret_cast<cudaStream_t>(stream));
}

int tx_tridiag_eigh_below_dev(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                              const int* n_states, int max_states, double* evals, double* evecs, double* scratch,
                              void* stream) {
    if (!d || !e || !cut_states || !n_states || !evals) return fail(TX_ERR_ARG, "null pointer argument");
    if (S < 2 || n_mat < 1 || max_states < 1) return fail(TX_ERR_ARG, "bad sizes S=%d n_mat=%d max_states=%d", S, n_mat, max_states);
    return launch_eig(d, e, S, n_mat, cut_states, n_states, max_states, evals, evecs, scratch,
                      reinterpret_cast<cudaStream_t>(stream));
}

int tx_tridiag_eigh_below(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                          const double* cut_count, int max_states, double* evals, double* evecs, int* n_states,
                          int* n_below) {
    if (!d || !e || !cut_states || !n_states) return fail(TX_ERR_ARG, "null pointer argument");
    if (S < 2 || n_mat < 1 || max_states < 0) return fail(TX_ERR_ARG, "bad sizes S=%d n_mat=%d max_states=%d", S, n_mat, max_states);
    if (max_states > 0 && !evals) return fail(TX_ERR_ARG, "null eigenvalue buffer");
    if (int rc = have_device()) return rc;
    if (!cut_count) cut_count = cut_states;
    const size_t bd = align256((size_t)n_mat * S * 8), be = align256((size_t)n_mat * (S - 1) * 8), bc = align256((size_t)n_mat * 8);
    const size_t bv = align256((size_t)n_mat * max_states * 8), bz = evecs ? align256((size_t)n_mat * max_states * S * 8) : 0;
    DevMem mem;
    TX_CUDA(cudaMalloc(&mem.p, bd + be + 4 * bc + bv + bz + 256));
    char* q = mem.p;
    double* dd = (double*)q;  q += bd;
    double* de = (double*)q;  q += be;
    double* dcs = (double*)q; q += bc;
    double* dcc = (double*)q; q += bc;
    int* dns = (int*)q;       q += bc;
    int* dnb = (int*)q;       q += bc;
    double* dv = (double*)q;  q += bv;
    double* dz = evecs ? (double*)q : nullptr;
    TX_CUDA(cudaMemcpy(dd, d, (size_t)n_mat * S * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(de, e, (size_t)n_mat * (S - 1) * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dcs, cut_states, (size_t)n_mat * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dcc, cut_count, (size_t)n_mat * 8, cudaMemcpyHostToDevice));
    if (int rc = launch_count(dd, de, S, n_mat, dcs, dcc, dns, dnb, nullptr)) return rc;
    if (max_states > 0)
        if (int rc = launch_eig(dd, de, S, n_mat, dcs, dns, max_states, dv, dz, nullptr, nullptr)) return rc;
    TX_CUDA(cudaMemcpy(n_states, dns, (size_t)n_mat * 4, cudaMemcpyDeviceToHost));
    if (n_below) TX_CUDA(cudaMemcpy(n_below, dnb, (size_t)n_mat * 4, cudaMemcpyDeviceToHost));
    if (max_states > 0) {
        TX_CUDA(cudaMemcpy(evals, dv, (size_t)n_mat * max_states * 8, cudaMemcpyDeviceToHost));
        if (evecs) TX_CUDA(cudaMemcpy(evecs, dz, (size_t)n_mat * max_states * S * 8, cudaMemcpyDeviceToHost));
    }
    return TX_OK;
}

long long tx_bardeen_workspace_bytes(int P, int n, int S, int max_states) {
    // psiL, psiR, support ranges of Hdiff
    return (long long)(2 * align256((size_t)P * n * max_states * S * 8) + align256((size_t)P * n * n * 8) + 256);
}

namespace {
struct EigFork {
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
}  // namespace

static int wells_dev_impl(const double* dL, const double* eL, const double* dR, const double* eR,
                          const double* cutL, const double* cutR_states, const double* cutR_count,
                          const double* hd0, const double* hdU, const double* hdL, int P, int n, int S, int max_states,
                          double interval, int sign_fix, double* EL, int* nL, double* ER, int* nR_states,
                          int* nR_below, double* Mavg, void* workspace, void* stream, bool counted) {
    if (!dL || !eL || !dR || !eR || !cutL || !cutR_states || !cutR_count || !hd0 || !hdU || !hdL || !EL || !nL || !ER ||
        !nR_states || !nR_below || !Mavg || !workspace)
        return fail(TX_ERR_ARG, "null pointer argument");
    if (P < 1 || n < 1 || S < 2 || max_states < 1) return fail(TX_ERR_ARG, "bad sizes P=%d n=%d S=%d max_states=%d", P, n, S, max_states);
    if ((size_t)S * 8 > 200 * 1024) return fail(TX_ERR_UNSUPPORTED, "chain length S=%d too long", S);
    if ((long long)n * n > 65535 || P > 65535 || (long long)P * n * n > 0x3fffffffLL)
        return fail(TX_ERR_UNSUPPORTED, "grid too large (n=%d, P=%d)", n, P);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bz = align256((size_t)P * n * max_states * S * 8);
    double* psiL = (double*)workspace;
    double* psiR = (double*)((char*)workspace + bz);
    const int n_mat = P * n;
    if (!counted) {
        if (int rc = launch_count(dL, eL, S, n_mat, cutL, cutL, nL, nR_below /* overwritten below */, st)) return rc;
        if (int rc = launch_count(dR, eR, S, n_mat, cutR_states, cutR_count, nR_states, nR_below, st)) return rc;
    }
    // The eigenproblems of the two wells are independent and their grids (a few hundred small CTAs for a sweep of 256
    // geometries) leave most of the GPU idle: the right well runs on an auxiliary stream, forked from and joined to `st`
    // by events (legal under stream capture too).  One auxiliary stream per device; the enqueue is serialised.
    {
        static EigFork forks[16];
        static std::mutex fork_mu;
        int dev = 0;
        TX_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(fork_mu);
        EigFork& f = forks[dev & 15];
        if (!f.aux) {
            TX_CUDA(cudaStreamCreateWithFlags(&f.aux, cudaStreamNonBlocking));
            TX_CUDA(cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming));
            TX_CUDA(cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming));
        }
        TX_CUDA(cudaEventRecord(f.fork, st));
        TX_CUDA(cudaStreamWaitEvent(f.aux, f.fork, 0));
        if (int rc = launch_eig(dL, eL, S, n_mat, cutL, nL, max_states, EL, psiL, nullptr, st)) return rc;
        if (int rc = launch_eig(dR, eR, S, n_mat, cutR_states, nR_states, max_states, ER, psiR, nullptr, f.aux)) return rc;
        TX_CUDA(cudaEventRecord(f.join, f.aux));
        TX_CUDA(cudaStreamWaitEvent(st, f.join, 0));
    }
    static bool attr_set = false;
    if (!attr_set) {
        TX_CUDA(cudaFuncSetAttribute(bardeen_melem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    int* support = (int*)((char*)workspace + 2 * bz);
    hdiff_support_kernel<<<P * n * n, 128, 0, st>>>(hd0, hdU, hdL, n, S, support);
    TX_CUDA(cudaGetLastError());
    count_launch();
    dim3 grid(max_states, n * n, P);
    bardeen_melem_kernel<<<grid, 128, (size_t)S * 8, st>>>(EL, nL, psiL, ER, nR_states, psiR, hd0, hdU, hdL, support, n, S,
                                                          max_states, interval, sign_fix, Mavg);
    TX_CUDA(cudaGetLastError());
    count_launch();
    return TX_OK;
}

int tx_bardeen_wells_dev(const double* dL, const double* eL, const double* dR, const double* eR,
                         const double* cutL, const double* cutR_states, const double* cutR_count,
                         const double* hd0, const double* hdU, const double* hdL, int P, int n, int S, int max_states,
                         double interval, int sign_fix, double* EL, int* nL, double* ER, int* nR_states,
                         int* nR_below, double* Mavg, void* workspace, void* stream) {
    return wells_dev_impl(dL, eL, dR, eR, cutL, cutR_states, cutR_count, hd0, hdU, hdL, P, n, S, max_states, interval, sign_fix,
                          EL, nL, ER, nR_states, nR_below, Mavg, workspace, stream, false);
}

// Grow-only per-device workspace of tx_bardeen_region_sweep (serialised by a mutex: one sweep at a time per process).
// All inputs travel as ONE host-to-device copy and all outputs as ONE device-to-host copy through a pinned staging block:
// the call used to issue ~20 small copies from pageable memory and a mid-stream synchronisation, which cost three times
// the device pass at S = 451.
namespace {
struct RegionWs {
    char* p = nullptr;
    size_t cap = 0;
    char* host = nullptr;      // pinned staging: packed inputs, then packed outputs
    size_t host_cap = 0;
    cudaStream_t st = nullptr;
};
RegionWs g_region_ws[16];
std::mutex g_region_mutex;
}  // namespace

int tx_bardeen_region_sweep(const double* tinf, const double* tL, const double* tR, const double* Vinf,
                            const double* VL, const double* VLprime, const double* VR, const double* VRprime, int n_pot,
                            const double* hc0, const double* hcU, const double* hcL,
                            const double* hp0, const double* hpU, const double* hpL, int n_hc,
                            int Ninf, int NL, int NR, int NC, int n, int P,
                            const double* cutL, const double* cutR_states, const double* cutR_count,
                            double interval, int sign_fix, int row_cut, int max_states, int* max_needed,
                            double* EL, int* nL, double* ER, int* nR_states, int* nR_below, double* Mavg,
                            const double* Vb, int nVb, double muR, double kBT, double* I) {
    if (!tinf || !tL || !tR || !Vinf || !VL || !VLprime || !VR || !VRprime || !hc0 || !cutL || !cutR_states || !cutR_count ||
        !max_needed || !nL || !nR_states || !nR_below)
        return fail(TX_ERR_ARG, "null pointer argument");
    if (Ninf < 1 || NL < 1 || NR < 1 || NC < 1 || NC % 2 != 1 || n < 1 || P < 1 || max_states < 0)
        return fail(TX_ERR_ARG, "bad sizes Ninf=%d NL=%d NR=%d NC=%d n=%d P=%d", Ninf, NL, NR, NC, n, P);
    if (NC > 1 && (!hcU || !hcL)) return fail(TX_ERR_ARG, "centre hoppings missing");
    if (!(n_pot == 1 || n_pot == P) || !(n_hc == 1 || n_hc == P)) return fail(TX_ERR_ARG, "n_pot / n_hc must be 1 or P");
    if (max_states > 0 && (!EL || !ER || !Mavg)) return fail(TX_ERR_ARG, "null output buffer");
    if (I && (!Vb || nVb < 1)) return fail(TX_ERR_ARG, "current requested without a bias grid");
    if (int rc = have_device()) return rc;
    if (!hp0) { hp0 = hc0; hpU = hcU; hpL = hcL; }
    const int S = 2 * Ninf + NL + NR + NC;
    if ((size_t)S * 8 > 200 * 1024) return fail(TX_ERR_UNSUPPORTED, "chain length S=%d too long", S);
    const int n_mat = P * n;
    const size_t nn = (size_t)n * n;
    const int ms = max_states > 0 ? max_states : 1;
    const size_t bd = align256((size_t)n_mat * S * 8), bc = align256((size_t)n_mat * 8);
    const size_t bh = align256((size_t)P * S * nn * 8);
    const size_t bv = align256((size_t)n_mat * ms * 8), bm = align256((size_t)n_mat * ms * n * 8);
    const size_t bw = max_states > 0 ? (size_t)tx_bardeen_workspace_bytes(P, n, S, max_states) : 0;
    const size_t bn = (size_t)n * 8, bpn = (size_t)n_pot * n * 8;
    const size_t b0 = (size_t)n_hc * NC * nn * 8, b1 = (size_t)n_hc * (NC - 1) * nn * 8;
    const size_t bI = I ? align256((size_t)P * nn * nVb * 8) : 0;
    // packed input block (same layout on the host and on the device)
    size_t in_bytes = 0;
    auto in_take = [&](size_t b) { size_t o = in_bytes; in_bytes += align256(b); return o; };
    const size_t o_par = in_take(4 * bn + 4 * bpn);
    size_t o_hc[6];
    for (int i = 0; i < 6; ++i) o_hc[i] = in_take((i % 3 == 0) ? b0 : (b1 ? b1 : 8));
    const size_t o_cL = in_take((size_t)n_mat * 8), o_cRs = in_take((size_t)n_mat * 8), o_cRc = in_take((size_t)n_mat * 8);
    const size_t o_Vb = in_take(I ? (size_t)nVb * 8 : 8);
    // packed output block: header (needed), the three counts, then (with max_states > 0) EL, ER, Mavg, I
    size_t out_bytes = 0;
    auto out_take = [&](size_t b) { size_t o = out_bytes; out_bytes += align256(b); return o; };
    const size_t o_max = out_take(256);
    const size_t o_nL = out_take((size_t)n_mat * 4), o_nRs = out_take((size_t)n_mat * 4), o_nRb = out_take((size_t)n_mat * 4);
    const size_t out_counts = out_bytes;
    const size_t o_EL = out_take(bv), o_ER = out_take(bv), o_M = out_take(bm);
    const size_t o_I = out_take(bI ? bI : 8);
    const size_t total = 4 * bd + 3 * bh + bm + bw + in_bytes + out_bytes + 512;
    const size_t host_total = in_bytes + out_bytes;

    int dev = 0;
    TX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_region_mutex);
    RegionWs& ws = g_region_ws[dev & 15];
    if (!ws.st) TX_CUDA(cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking));
    if (total > ws.cap) {
        if (ws.p) cudaFree(ws.p);
        ws.p = nullptr;
        ws.cap = 0;
        TX_CUDA(cudaMalloc(&ws.p, total + total / 4));
        ws.cap = total + total / 4;
    }
    if (host_total > ws.host_cap) {
        if (ws.host) cudaFreeHost(ws.host);
        ws.host = nullptr;
        ws.host_cap = 0;
        TX_CUDA(cudaHostAlloc((void**)&ws.host, host_total + host_total / 4, cudaHostAllocDefault));
        ws.host_cap = host_total + host_total / 4;
    }
    cudaStream_t st = ws.st;
    char* q = ws.p;
    auto take = [&](size_t b) { char* r = q; q += b; return r; };
    double* ddL = (double*)take(bd); double* deL = (double*)take(bd);
    double* ddR = (double*)take(bd); double* deR = (double*)take(bd);
    double* dh0 = (double*)take(bh); double* dhU = (double*)take(bh); double* dhL = (double*)take(bh);
    double* dM2 = (double*)take(bm);
    char* wsp = take(bw);
    char* din = take(in_bytes);
    char* dout = take(out_bytes);
    char* hin = ws.host;
    char* hout = ws.host + in_bytes;

    // pack and upload
    {
        double* hp = (double*)(hin + o_par);
        std::memcpy(hp, tinf, bn); std::memcpy(hp + n, tL, bn); std::memcpy(hp + 2 * n, tR, bn); std::memcpy(hp + 3 * n, Vinf, bn);
        const size_t np = (size_t)n_pot * n;
        std::memcpy(hp + 4 * n, VL, bpn); std::memcpy(hp + 4 * n + np, VLprime, bpn);
        std::memcpy(hp + 4 * n + 2 * np, VR, bpn); std::memcpy(hp + 4 * n + 3 * np, VRprime, bpn);
        std::memcpy(hin + o_hc[0], hc0, b0); std::memcpy(hin + o_hc[3], hp0, b0);
        if (NC > 1) {
            std::memcpy(hin + o_hc[1], hcU, b1); std::memcpy(hin + o_hc[2], hcL, b1);
            std::memcpy(hin + o_hc[4], hpU, b1); std::memcpy(hin + o_hc[5], hpL, b1);
        }
        std::memcpy(hin + o_cL, cutL, (size_t)n_mat * 8); std::memcpy(hin + o_cRs, cutR_states, (size_t)n_mat * 8);
        std::memcpy(hin + o_cRc, cutR_count, (size_t)n_mat * 8);
        if (I) std::memcpy(hin + o_Vb, Vb, (size_t)nVb * 8);
    }
    TX_CUDA(cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, st));
    const bool full = max_states > 0;
    TX_CUDA(cudaMemsetAsync(dout, 0, full ? out_bytes : out_counts, st));   // needed = 0; EL, ER, Mavg zero beyond the counts

    double* dpar = (double*)(din + o_par);
    double* dcL = (double*)(din + o_cL); double* dcRs = (double*)(din + o_cRs); double* dcRc = (double*)(din + o_cRc);
    double* dVb = (double*)(din + o_Vb);
    int* dmax = (int*)(dout + o_max);
    int* dnL = (int*)(dout + o_nL); int* dnRs = (int*)(dout + o_nRs); int* dnRb = (int*)(dout + o_nRb);
    double* dEL = (double*)(dout + o_EL); double* dER = (double*)(dout + o_ER); double* dM = (double*)(dout + o_M);
    double* dI = (double*)(dout + o_I);

    RegionParams rp{};
    rp.tinf = dpar; rp.tL = dpar + n; rp.tR = dpar + 2 * n; rp.Vinf = dpar + 3 * n;
    rp.VL = dpar + 4 * n; rp.VLp = rp.VL + (size_t)n_pot * n; rp.VR = rp.VLp + (size_t)n_pot * n;
    rp.VRp = rp.VR + (size_t)n_pot * n;
    rp.v_stride = n_pot == 1 ? 0 : n;
    rp.hc0 = (double*)(din + o_hc[0]); rp.hcU = (double*)(din + o_hc[1]); rp.hcL = (double*)(din + o_hc[2]);
    rp.hp0 = (double*)(din + o_hc[3]); rp.hpU = (double*)(din + o_hc[4]); rp.hpL = (double*)(din + o_hc[5]);
    rp.hc_stride0 = n_hc == 1 ? 0 : (long long)NC * nn;
    rp.hc_stride1 = n_hc == 1 ? 0 : (long long)(NC - 1) * nn;
    rp.Ninf = Ninf; rp.NL = NL; rp.NR = NR; rp.NC = NC; rp.n = n; rp.P = P; rp.S = S;
    const long long items = (long long)P * S;
    bardeen_assemble_kernel<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(rp, ddL, deL, ddR, deR, dh0, dhU, dhL);
    TX_CUDA(cudaGetLastError());
    count_launch();
    if (int rc = launch_count(ddL, deL, S, n_mat, dcL, dcL, dnL, dnRb, st)) return rc;
    if (int rc = launch_count(ddR, deR, S, n_mat, dcRs, dcRc, dnRs, dnRb, st)) return rc;
    max_int_kernel<<<1, 256, 0, st>>>(dnL, dnRs, n_mat, dmax);
    TX_CUDA(cudaGetLastError());
    count_launch();
    // No host round trip for the counts: the eigensolver and the matrix elements run against the caller's capacity (they
    // clamp to max_states); when it turns out too small the caller is told through *max_needed and calls again.
    if (full) {
        if (int rc = wells_dev_impl(ddL, deL, ddR, deR, dcL, dcRs, dcRc, dh0, dhU, dhL, P, n, S, max_states, interval, sign_fix,
                                    dEL, dnL, dER, dnRs, dnRb, dM, wsp, st, true))
            return rc;
        if (I) {
            const long long cnt = (long long)n_mat * max_states * n;
            square_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(dM, dM2, cnt, max_states, n, dnRb, row_cut);
            TX_CUDA(cudaGetLastError());
            count_launch();
            if (int rc = tx_bardeen_current_dev(dEL, dM2, P, n, max_states, dVb, nVb, muR, kBT, dI, st)) return rc;
        }
    }
    TX_CUDA(cudaMemcpyAsync(hout, dout, full ? out_bytes : out_counts, cudaMemcpyDeviceToHost, st));
    TX_CUDA(cudaStreamSynchronize(st));
    const int needed = *(const int*)(hout + o_max);
    *max_needed = needed;
    std::memcpy(nL, hout + o_nL, (size_t)n_mat * 4);
    std::memcpy(nR_states, hout + o_nRs, (size_t)n_mat * 4);
    std::memcpy(nR_below, hout + o_nRb, (size_t)n_mat * 4);
    if (!full || needed > max_states || needed == 0) return TX_OK;   // counts only: size the buffers from *max_needed, call again
    std::memcpy(EL, hout + o_EL, (size_t)n_mat * max_states * 8);
    std::memcpy(ER, hout + o_ER, (size_t)n_mat * max_states * 8);
    std::memcpy(Mavg, hout + o_M, (size_t)n_mat * max_states * n * 8);
    if (I) std::memcpy(I, hout + o_I, (size_t)P * nn * nVb * 8);
    return TX_OK;
}

int tx_bardeen_wells(const double* dL, const double* eL, const double* dR, const double* eR, const double* cutL,
                     const double* cutR_states, const double* cutR_count, const double* hd0, const double* hdU,
                     const double* hdL, int P, int n, int S, int max_states, double interval, int sign_fix, double* EL,
                     int* nL, double* ER, int* nR_states, int* nR_below, double* Mavg, double* psiL, double* psiR) {
    if (!dL || !eL || !dR || !eR || !cutL || !cutR_states || !cutR_count || !hd0 || !hdU || !hdL || !nL || !nR_states ||
        !nR_below)
        return fail(TX_ERR_ARG, "null pointer argument");
    if (P < 1 || n < 1 || S < 2 || max_states < 0) return fail(TX_ERR_ARG, "bad sizes P=%d n=%d S=%d max_states=%d", P, n, S, max_states);
    if (max_states > 0 && (!EL || !ER || !Mavg)) return fail(TX_ERR_ARG, "null output buffer");
    if (int rc = have_device()) return rc;
    const int n_mat = P * n;
    const size_t bd = align256((size_t)n_mat * S * 8), be = align256((size_t)n_mat * (S - 1) * 8), bc = align256((size_t)n_mat * 8);
    const size_t nn = (size_t)n * n;
    const size_t bh0 = align256((size_t)P * S * nn * 8), bh1 = align256((size_t)P * (S - 1) * nn * 8);
    const int ms = max_states > 0 ? max_states : 1;
    const size_t bv = align256((size_t)n_mat * ms * 8), bm = align256((size_t)n_mat * ms * n * 8);
    const size_t bw = max_states > 0 ? (size_t)tx_bardeen_workspace_bytes(P, n, S, max_states) : 0;
    DevMem mem;
    TX_CUDA(cudaMalloc(&mem.p, 2 * bd + 2 * be + 6 * bc + bh0 + 2 * bh1 + 2 * bv + bm + bw + 256));
    char* q = mem.p;
    auto take = [&](size_t b) { char* r = q; q += b; return r; };
    double* ddL = (double*)take(bd); double* deL = (double*)take(be);
    double* ddR = (double*)take(bd); double* deR = (double*)take(be);
    double* dcL = (double*)take(bc); double* dcRs = (double*)take(bc); double* dcRc = (double*)take(bc);
    int* dnL = (int*)take(bc); int* dnRs = (int*)take(bc); int* dnRb = (int*)take(bc);
    double* dh0 = (double*)take(bh0); double* dhU = (double*)take(bh1); double* dhL = (double*)take(bh1);
    double* dEL = (double*)take(bv); double* dER = (double*)take(bv); double* dM = (double*)take(bm);
    char* ws = take(bw);
    auto up = [&](void* dst, const void* src, size_t b) { return cudaMemcpy(dst, src, b, cudaMemcpyHostToDevice); };
    TX_CUDA(up(ddL, dL, (size_t)n_mat * S * 8));
    TX_CUDA(up(deL, eL, (size_t)n_mat * (S - 1) * 8));
    TX_CUDA(up(ddR, dR, (size_t)n_mat * S * 8));
    TX_CUDA(up(deR, eR, (size_t)n_mat * (S - 1) * 8));
    TX_CUDA(up(dcL, cutL, (size_t)n_mat * 8));
    TX_CUDA(up(dcRs, cutR_states, (size_t)n_mat * 8));
    TX_CUDA(up(dcRc, cutR_count, (size_t)n_mat * 8));
    TX_CUDA(up(dh0, hd0, (size_t)P * S * nn * 8));
    TX_CUDA(up(dhU, hdU, (size_t)P * (S - 1) * nn * 8));
    TX_CUDA(up(dhL, hdL, (size_t)P * (S - 1) * nn * 8));
    if (max_states == 0) {
        // counting pass only: the caller sizes its buffers from the counts and calls again
        if (int rc = launch_count(ddL, deL, S, n_mat, dcL, dcL, dnL, dnRb, nullptr)) return rc;
        if (int rc = launch_count(ddR, deR, S, n_mat, dcRs, dcRc, dnRs, dnRb, nullptr)) return rc;
    } else {
        if (int rc = tx_bardeen_wells_dev(ddL, deL, ddR, deR, dcL, dcRs, dcRc, dh0, dhU, dhL, P, n, S, max_states, interval,
                                          sign_fix, dEL, dnL, dER, dnRs, dnRb, dM, ws, nullptr))
            return rc;
    }
    auto down = [&](void* dst, const void* src, size_t b) { return cudaMemcpy(dst, src, b, cudaMemcpyDeviceToHost); };
    TX_CUDA(down(nL, dnL, (size_t)n_mat * 4));
    TX_CUDA(down(nR_states, dnRs, (size_t)n_mat * 4));
    TX_CUDA(down(nR_below, dnRb, (size_t)n_mat * 4));
    if (max_states > 0) {
        TX_CUDA(down(EL, dEL, (size_t)n_mat * max_states * 8));
        TX_CUDA(down(ER, dER, (size_t)n_mat * max_states * 8));
        TX_CUDA(down(Mavg, dM, (size_t)n_mat * max_states * n * 8));
        const size_t bz = align256((size_t)n_mat * max_states * S * 8);
        if (psiL) TX_CUDA(down(psiL, ws, (size_t)n_mat * max_states * S * 8));
        if (psiR) TX_CUDA(down(psiR, ws + bz, (size_t)n_mat * max_states * S * 8));
    }
    return TX_OK;
}

int tx_bardeen_current_dev(const double* E, const double* M2, int P, int n, int nb, const double* Vb, int nVb, double muR,
                           double kBT, double* I, void* stream) {
    if (!E || !M2 || !Vb || !I) return fail(TX_ERR_ARG, "null pointer argument");
    if (P < 1 || n < 1 || nb < 1 || nVb < 1) return fail(TX_ERR_ARG, "bad sizes P=%d n=%d nb=%d nVb=%d", P, n, nb, nVb);
    const long long warps = (long long)P * n * n * nVb;
    bardeen_current_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        E, M2, P, n, nb, Vb, nVb, muR, kBT, I);
    TX_CUDA(cudaGetLastError());
    count_launch();
    return TX_OK;
}

int tx_bardeen_current(const double* E, const double* M2, int P, int n, int nb, const double* Vb, int nVb, double muR,
                       double kBT, double* I) {
    if (!E || !M2 || !Vb || !I) return fail(TX_ERR_ARG, "null pointer argument");
    if (P < 1 || n < 1 || nb < 1 || nVb < 1) return fail(TX_ERR_ARG, "bad sizes P=%d n=%d nb=%d nVb=%d", P, n, nb, nVb);
    if (int rc = have_device()) return rc;
    const size_t bE = align256((size_t)P * n * nb * 8), bM = align256((size_t)P * n * nb * n * 8), bV = align256((size_t)nVb * 8);
    const size_t bI = (size_t)P * n * n * nVb * 8;
    DevMem mem;
    TX_CUDA(cudaMalloc(&mem.p, bE + bM + bV + bI + 256));
    double* dE = (double*)mem.p;
    double* dM = (double*)(mem.p + bE);
    double* dV = (double*)(mem.p + bE + bM);
    double* dI = (double*)(mem.p + bE + bM + bV);
    TX_CUDA(cudaMemcpy(dE, E, (size_t)P * n * nb * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dM, M2, (size_t)P * n * nb * n * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dV, Vb, (size_t)nVb * 8, cudaMemcpyHostToDevice));
    if (int rc = tx_bardeen_current_dev(dE, dM, P, n, nb, dV, nVb, muR, kBT, dI, nullptr)) return rc;
    TX_CUDA(cudaMemcpy(I, dI, bI, cudaMemcpyDeviceToHost));
    return TX_OK;
}

int tx_real_gemm(const double* A, const double* B, const double* w, double* C, int M, int N, int K, int transA, int transB) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) return fail(TX_ERR_ARG, "bad arguments");
    if (int rc = have_device()) return rc;
    const size_t bA = align256((size_t)M * K * 8), bB = align256((size_t)N * K * 8), bC = align256((size_t)M * N * 8), bw = align256((size_t)K * 8);
    DevMem mem;
    TX_CUDA(cudaMalloc(&mem.p, bA + bB + bC + bw + 256));
    double* dA = (double*)mem.p;
    double* dB = (double*)(mem.p + bA);
    double* dC = (double*)(mem.p + bA + bB);
    double* dw = (double*)(mem.p + bA + bB + bC);
    TX_CUDA(cudaMemcpy(dA, A, (size_t)M * K * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dB, B, (size_t)N * K * 8, cudaMemcpyHostToDevice));
    if (w) TX_CUDA(cudaMemcpy(dw, w, (size_t)K * 8, cudaMemcpyHostToDevice));
    dim3 block(32, 8), grid((N + 31) / 32, (M + 7) / 8);
    real_gemm_kernel<<<grid, block>>>(dA, dB, w ? dw : nullptr, dC, M, N, K, transA, transB);
    TX_CUDA(cudaGetLastError());
    count_launch();
    TX_CUDA(cudaMemcpy(C, dC, (size_t)M * N * 8, cudaMemcpyDeviceToHost));
    return TX_OK;
}

int tx_bardeen_ts(const double* E, const double* M2, int n, int nb, const double* tL, const double* VL, const double* tR,
                  const double* VR, int NL, int NR, double* T, int* flags) {
    if (!E || !M2 || !tL || !VL || !tR || !VR || !T || !flags) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || nb < 1) return fail(TX_ERR_ARG, "bad sizes n=%d nb=%d", n, nb);
    if (int rc = have_device()) return rc;
    const size_t bE = align256((size_t)n * nb * 8), bM = align256((size_t)n * nb * n * 8), bp = align256((size_t)n * 8);
    DevMem mem;
    TX_CUDA(cudaMalloc(&mem.p, bE + 2 * bM + 4 * bp + (size_t)(2 + n * n) * 4 + 256));
    char* q = mem.p;
    double* dE = (double*)q;  q += bE;
    double* dM = (double*)q;  q += bM;
    double* dT = (double*)q;  q += bM;
    double* dp[4];
    const double* hp[4] = {tL, VL, tR, VR};
    for (int i = 0; i < 4; ++i) {
        dp[i] = (double*)q;
        q += bp;
        TX_CUDA(cudaMemcpy(dp[i], hp[i], (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    int* dflag = (int*)q;
    TX_CUDA(cudaMemset(dflag, 0, (size_t)(2 + n * n) * 4));
    TX_CUDA(cudaMemcpy(dE, E, (size_t)n * nb * 8, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemcpy(dM, M2, (size_t)n * nb * n * 8, cudaMemcpyHostToDevice));
    const int total = n * nb * n;
    bardeen_ts_kernel<<<(total + 127) / 128, 128>>>(dE, dM, n, nb, dp[0], dp[1], dp[2], dp[3], (double)NL, (double)NR, dT, dflag);
    TX_CUDA(cudaGetLastError());
    count_launch();
    TX_CUDA(cudaMemcpy(T, dT, (size_t)total * 8, cudaMemcpyDeviceToHost));
    TX_CUDA(cudaMemcpy(flags, dflag, (size_t)(2 + n * n) * 4, cudaMemcpyDeviceToHost));
    return TX_OK;
}

}  // extern "C"
