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
__device__ inline void cta_gemm(cd* C, const cd* A, int opA, const cd* B, int opB, int n) {
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

// C += sgn * op(A) * op(B)
__device__ inline void cta_gemm_acc(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn) {
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
        // eliminate column k from the other rows
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int r = idx / n, c = idx - r * n;
            if (r == k) continue;
            const cd f = colk[r];
            const cd pk = A[k * n + c];
            cd a = (c == k) ? mk(0.0, 0.0) : A[idx];
            cfms(a.x, a.y, f.x, f.y, pk.x, pk.y);
            A[idx] = a;
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
