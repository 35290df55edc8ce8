// Shared device helpers: complex128 arithmetic on (re, im) double pairs, closed-form lead GF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/transport_b200.h"

namespace tx {

// Largest block edge (n_loc_dof) of the CTA-cooperative large-block paths, and the largest one whose inverse staging block
// (n (n + 1) 16 B = 175 KB at n = 104) still fits the shared memory of an SM next to the static arrays of the kernels.
constexpr int TX_MAX_N = 128;
constexpr int TX_STAGE_SMEM_MAX_N = 104;


struct __align__(16) cd {
    double x, y;
};

// Multi-GPU gather fused into the K3 epilogue: the per-point total transmission is ALSO stored into the result
// buffers of up to 8 ranks (peer memory mapped over NVLink with CUDA IPC, peer.cu) at element `off + point`, so the
// "NCCL all-gather of the T(E) arrays" of SURVEY.md §8e rides on the sweep itself instead of following it.
struct PeerGather {
    double* dst[8];
    long long off;
    int n;
};
__device__ __forceinline__ void store_total(double* ttot, const PeerGather& pg, long long pt, double v) {
    if (ttot) ttot[pt] = v;
    for (int q = 0; q < pg.n; ++q) pg.dst[q][pg.off + pt] = v;
}

__host__ __device__ __forceinline__ cd mk(double x, double y) {
    cd r;
    r.x = x;
    r.y = y;
    return r;
}
__device__ __forceinline__ cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd operator*(cd a, cd b) {
    return mk(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cd conj(cd a) { return mk(a.x, -a.y); }
__device__ __forceinline__ cd scale(cd a, double s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ double norm2(cd a) { return fma(a.x, a.x, a.y * a.y); }

// acc += a*b  (4 DFMA)
__device__ __forceinline__ void cfma(double& ar, double& ai, double br, double bi, double cr, double ci) {
    ar = fma(br, cr, ar);
    ar = fma(-bi, ci, ar);
    ai = fma(br, ci, ai);
    ai = fma(bi, cr, ai);
}
// acc -= a*b  (4 DFMA)
__device__ __forceinline__ void cfms(double& ar, double& ai, double br, double bi, double cr, double ci) {
    ar = fma(-br, cr, ar);
    ar = fma(bi, ci, ar);
    ai = fma(-br, ci, ai);
    ai = fma(-bi, cr, ai);
}

// 1/z.  Blocks here are O(1) in magnitude (energies and hoppings of order the band width), so the
// plain conj(z)/|z|^2 form is safe from over/underflow.
__device__ __forceinline__ cd crecip(cd z) {
    double d = 1.0 / norm2(z);
    return mk(z.x * d, -z.y * d);
}

// Complex division with Smith's algorithm — the form numpy and CPython use, so that the
// closed-form lead GF follows the reference's rounding (exact for a real divisor).
__device__ __forceinline__ cd cdiv(cd a, cd b) {
    if (fabs(b.x) >= fabs(b.y)) {
        if (b.x == 0.0 && b.y == 0.0) return mk(a.x / fabs(b.x), a.y / fabs(b.x));
        double rat = b.y / b.x;
        double scl = 1.0 / (b.x + b.y * rat);
        return mk((a.x + a.y * rat) * scl, (a.y - a.x * rat) * scl);
    } else {
        double rat = b.x / b.y;
        double scl = 1.0 / (b.x * rat + b.y);
        return mk((a.x * rat + a.y) * scl, (a.y * rat - a.x) * scl);
    }
}

// Principal complex square root (C99 csqrt without the over/underflow rescaling).
__device__ __forceinline__ cd csqrt_principal(cd z) {
    if (z.x == 0.0 && z.y == 0.0) return mk(0.0, z.y);
    double m = hypot(z.x, z.y);
    if (z.x >= 0.0) {
        double t = sqrt((z.x + m) * 0.5);
        return mk(t, z.y / (2.0 * t));
    }
    double t = sqrt((-z.x + m) * 0.5);
    return mk(fabs(z.y) / (2.0 * t), copysign(t, z.y));
}

// numpy >= 2 sign of a complex number: z/|z| (0 for 0); equals the real sign for Im z = 0,
// which is how every reference driver calls g_closed (SURVEY.md §7 hard part 2).
__device__ __forceinline__ cd csign(cd z) {
    if (z.y == 0.0) return mk((z.x > 0.0) - (z.x < 0.0), 0.0);
    double m = hypot(z.x, z.y);
    return mk(z.x / m, z.y / m);
}

// wfm.g_closed for one channel (wfm/__init__.py:337-340):
//   lam = (E - eps)/(2 t); g = Re(lam/t) - sign(E-eps) |Re sqrt(lam^2-1)| - i |Im sqrt(lam^2-1)|
__device__ __forceinline__ cd g_closed_channel(cd E, cd eps, cd t) {
    if (E.y == 0.0 && eps.y == 0.0 && t.y == 0.0) {
        // All reference drivers: real energy, real lead parameters.  Same formula, real arithmetic
        // (identical results: Smith division by a real divisor and csqrt of a real are exact maps).
        const double dr = E.x - eps.x;
        double lam, first;
        if ((__double2hiint(t.x) & 0x000fffff) == 0 && __double2loint(t.x) == 0 && fabs(t.x) > 1e-300 && fabs(t.x) < 1e300) {
            // |t| a power of two (t = -1 in every reference driver): the two divisions are exact products
            const double rt = 1.0 / t.x;
            lam = dr * (0.5 * rt);
            first = lam * rt;
        } else {
            lam = dr / (2.0 * t.x);
            first = lam / t.x;
        }
        const double z = __dsub_rn(__dmul_rn(lam, lam), 1.0);   // rounded product first, like numpy (no fma)
        const double sg = (dr > 0.0) - (dr < 0.0);
        const double root = sqrt(fabs(z));
        return (z >= 0.0) ? mk(first - sg * root, -0.0) : mk(first, -root);
    }
    cd d = E - eps;
    cd lam = cdiv(d, mk(2.0 * t.x, 2.0 * t.y));
    cd first = cdiv(lam, t);
    // numpy's complex product is (ac-bd, ad+bc) with every product rounded (no fma): near a band edge
    // lam^2-1 cancels, so the rounding order is kept identical to stay at the 1e-14 level of parity.
    cd sq = mk(__dsub_rn(__dmul_rn(lam.x, lam.x), __dmul_rn(lam.y, lam.y)),
               __dadd_rn(__dmul_rn(lam.x, lam.y), __dmul_rn(lam.y, lam.x)));
    sq.x = __dsub_rn(sq.x, 1.0);
    cd root = csqrt_principal(sq);
    cd sg = csign(d);
    double ar = fabs(root.x), ai = fabs(root.y);
    return mk(first.x - sg.x * ar, -sg.y * ar - ai);
}

}  // namespace tx
