// CTA-cooperative operations on n x n complex128 blocks held in shared (n <= 32) or global memory.
// These serve the generic-size paths (lead surface Green's functions, full Green's function,
// any block size) where one CTA owns one energy; the n <= 16 T(E) hot path has its own
// register-resident kernel in rgf_small.cuh.  All routines end with __syncthreads().
#pragma once
#include "tx_common.cuh"

namespace tx {

enum { OP_N = 0, OP_H = 1 };   // as is / conjugate transpose

__device__ __forceinline__ cd ld_op(const cd* A, int n, int r, int c, int op) {
    if (op == OP_N) return A[r * n + c];
    const cd v = A[c * n + r];
    return mk(v.x, -v.y);
}

// C = op(A) * op(B)            (C must not alias A or B)
__device__ inline void cta_gemm_dmma(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn, bool accumulate);

__device__ inline void cta_gemm(cd* C, const cd* A, int opA, const cd* B, int opB, int n) {
    if ((n & 7) == 0 && n >= 16) {
        cta_gemm_dmma(C, A, opA, B, opB, n, 1.0, false);
        return;
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        double ar = 0.0, ai = 0.0;
        for (int k = 0; k < n; ++k) {
            const cd a = ld_op(A, n, r, k, opA);
            const cd b = ld_op(B, n, k, c, opB);
            cfma(ar, ai, a.x, a.y, b.x, b.y);
        }
        C[idx] = mk(ar, ai);
    }
    __syncthreads();
}

// ---- FP64 tensor-core path (mma.sync.m8n8k4.f64) for n a multiple of 8 -------------------------------
// One warp per 8x8 output tile, k advanced 4 at a time; a complex product is four real DMMAs
// (re += ar*br, re += (-ai)*bi, im += ar*bi, im += ai*br).  Fragment ownership (lane = 4*g + t):
//   A (8x4, row): lane holds A[g][t]     B (4x8, col): lane holds B[t][g]     C (8x8): lane holds C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma844(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// C = sgn*op(A)*op(B) (+ C if accumulate).  C must not alias A or B.  Requires n % 8 == 0.
// A warp works on a strip of up to four 8x8 output tiles of one tile row: the A fragment of a k-step is loaded
// once and reused for the four column tiles (16 DMMA per 5 fragment loads), and the loads of the next k-step are
// issued before the DMMAs of the current one (operands usually sit in L2/L1, not in shared memory).
__device__ inline void cta_gemm_dmma(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn, bool accumulate) {
    constexpr int NT = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = n >> 3;
    const int strips = (tiles + NT - 1) / NT;
    for (int item = warp; item < tiles * strips; item += nwarps) {
        const int tr = (item / strips) << 3;
        const int tc0 = (item % strips) * NT;
        const int nt = min(NT, tiles - tc0);
        double re[NT][2], im[NT][2];
#pragma unroll
        for (int u = 0; u < NT; ++u) re[u][0] = re[u][1] = im[u][0] = im[u][1] = 0.0;
        cd a = ld_op(A, n, tr + g, t, opA);
        cd b[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) b[u] = (u < nt) ? ld_op(B, n, t, ((tc0 + u) << 3) + g, opB) : mk(0.0, 0.0);
        for (int k = 0; k < n; k += 4) {
            const cd ac = a;
            cd bc[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) bc[u] = b[u];
            if (k + 4 < n) {
                a = ld_op(A, n, tr + g, k + 4 + t, opA);
#pragma unroll
                for (int u = 0; u < NT; ++u)
                    if (u < nt) b[u] = ld_op(B, n, k + 4 + t, ((tc0 + u) << 3) + g, opB);
            }
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                dmma844(re[u], ac.x, bc[u].x);
                dmma844(im[u], ac.x, bc[u].y);
            }
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                dmma844(re[u], -ac.y, bc[u].y);
                dmma844(im[u], ac.y, bc[u].x);
            }
        }
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            if (u < nt) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int idx = (tr + g) * n + ((tc0 + u) << 3) + 2 * t + i;
                    cd v = mk(sgn * re[u][i], sgn * im[u][i]);
                    if (accumulate) {
                        const cd c0 = C[idx];
                        v.x += c0.x;
                        v.y += c0.y;
                    }
                    C[idx] = v;
                }
            }
        }
    }
    __syncthreads();
}

// C += sgn * op(A) * op(B)
__device__ inline void cta_gemm_acc(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn) {
    if ((n & 7) == 0 && n >= 16) {
        cta_gemm_dmma(C, A, opA, B, opB, n, sgn, true);
        return;
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        double ar = 0.0, ai = 0.0;
        for (int k = 0; k < n; ++k) {
            const cd a = ld_op(A, n, r, k, opA);
            const cd b = ld_op(B, n, k, c, opB);
            cfma(ar, ai, a.x, a.y, b.x, b.y);
        }
        cd v = C[idx];
        v.x = fma(sgn, ar, v.x);
        v.y = fma(sgn, ai, v.y);
        C[idx] = v;
    }
    __syncthreads();
}

__device__ inline void cta_copy(cd* dst, const cd* src, int count) {
    for (int idx = threadIdx.x; idx < count; idx += blockDim.x) dst[idx] = src[idx];
    __syncthreads();
}

// A = z*I - H
__device__ inline void cta_z_minus(cd* A, cd z, const cd* H, int n) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        const cd hv = H[idx];
        A[idx] = (r == c) ? mk(z.x - hv.x, z.y - hv.y) : mk(-hv.x, -hv.y);
    }
    __syncthreads();
}

// A -= B
__device__ inline void cta_sub(cd* A, const cd* B, int count) {
    for (int idx = threadIdx.x; idx < count; idx += blockDim.x) {
        cd a = A[idx];
        const cd b = B[idx];
        a.x -= b.x;
        a.y -= b.y;
        A[idx] = a;
    }
    __syncthreads();
}

// In-place inverse by Gauss-Jordan with partial (row) pivoting.  `colk` is scratch of n entries,
// `ipiv` scratch of n ints (both shared memory).  Sets *singular = 1 on a zero pivot.
__device__ inline void cta_inverse(cd* A, int n, cd* colk, int* ipiv, int* singular) {
    __shared__ int s_piv;
    __shared__ cd s_inv;
    for (int k = 0; k < n; ++k) {
        // pivot search by warp 0
        if (threadIdx.x < 32) {
            double best = -1.0;
            int brow = k;
            for (int r = k + threadIdx.x; r < n; r += 32) {
                const double m = norm2(A[r * n + k]);
                if (m > best || !(m == m)) {
                    best = m;
                    brow = r;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
                if (ob > best || (ob == best && orow < brow)) {
                    best = ob;
                    brow = orow;
                }
            }
            if (threadIdx.x == 0) {
                s_piv = brow;
                ipiv[k] = brow;
                if (best == 0.0) *singular = 1;
            }
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k) {
            for (int c = threadIdx.x; c < n; c += blockDim.x) {
                const cd a = A[k * n + c];
                A[k * n + c] = A[p * n + c];
                A[p * n + c] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) s_inv = crecip(A[k * n + k]);
        for (int r = threadIdx.x; r < n; r += blockDim.x) colk[r] = A[r * n + k];
        __syncthreads();
        const cd inv = s_inv;
        // scale the pivot row (its k-th entry becomes 1/pivot)
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            A[k * n + c] = (c == k) ? inv : A[k * n + c] * inv;
        }
        __syncthreads();
        // eliminate column k from the other rows (row/column indices advanced without a division)
        {
            int r = threadIdx.x / n, c = threadIdx.x - r * n;
            const int dr = blockDim.x / n, dc = blockDim.x - dr * n;
            for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
                if (r != k) {
                    const cd f = colk[r];
                    const cd pk = A[k * n + c];
                    cd a = (c == k) ? mk(0.0, 0.0) : A[idx];
                    cfms(a.x, a.y, f.x, f.y, pk.x, pk.y);
                    A[idx] = a;
                }
                r += dr;
                c += dc;
                if (c >= n) {
                    c -= n;
                    ++r;
                }
            }
        }
        __syncthreads();
    }
    // undo the row exchanges: column swaps in reverse order
    for (int k = n - 1; k >= 0; --k) {
        const int p = ipiv[k];
        if (p != k) {
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                const cd a = A[r * n + k];
                A[r * n + k] = A[r * n + p];
                A[r * n + p] = a;
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

// Inverse of X using the shared-memory staging block S when X itself lives in global memory: every
// elimination step rewrites the whole block, which must not go through L2 sixty-four times.
__device__ inline void cta_inverse_staged(cd* X, cd* S, int n, cd* colk, int* ipiv, int* singular) {
    if (S == nullptr || S == X) {
        cta_inverse(X, n, colk, ipiv, singular);
        return;
    }
    cta_copy(S, X, n * n);
    cta_inverse(S, n, colk, ipiv, singular);
    cta_copy(X, S, n * n);
}

// max |A_ij|^2 over the block, broadcast to all threads
__device__ inline double cta_maxnorm2(const cd* A, int count) {
    __shared__ double s_red[32];
    double m = 0.0;
    for (int idx = threadIdx.x; idx < count; idx += blockDim.x) {
        const double v = norm2(A[idx]);
        m = (v > m || !(v == v)) ? v : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = (o > m || !(o == o)) ? o : m;
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) {
        const double o = s_red[w];
        r = (o > r || !(o == o)) ? o : r;
    }
    __syncthreads();
    return r;
}

}  // namespace tx
