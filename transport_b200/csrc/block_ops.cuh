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

// ---- TMA staging of a product's B operand (n = 64, blocks in the L2-resident global workspace) ------------------
// ncu on the direct-from-L2 product showed operand fetch as half of the stall samples: every B fragment is re-read by
// the eight tile rows and every A fragment by both strips, 640 KB of L2 traffic per 64 x 64 product and CTA, 7 TB/s
// over the chip.  Here the whole B block (64 KB) is brought into shared memory ONCE per product by the TMA engine
// (cp.async.bulk global -> shared, one 1 KB row per copy into rows padded to 66 elements so that the DMMA fragment
// reads are bank-conflict free, completion on an mbarrier), A fragments come straight from L2 (one tile row per warp,
// so each is read once): 128 KB per product.  The staging block is the one the blocked inverse uses between products.
struct TmaStage {
    cd* buf;                       // shared memory, TMA_LD * 64 elements
    unsigned long long* bar;       // mbarrier (shared memory), initialised with tma_stage_init
    unsigned parity;               // phase parity of the next wait (per thread, all threads in step)
};
constexpr int TMA_LD = 66;
// Optional fused epilogue of the TMA-staged product: a second accumulate destination (C2 += sgn A B, e.g. the two
// renormalised on-site blocks of the Sancho-Rubio step that receive the same product) and a per-thread running maximum of
// |stored value|^2 (the convergence test of the decimation without another pass over the blocks).
struct GemmEpi {
    cd* C2 = nullptr;
    cd* C3 = nullptr;              // third destination: C3 += c3s * sgn A B  (Hermitian-hopping decimation: bulk block += 2 alpha G alpha)
    double c3s = 1.0;
    double* tmax = nullptr;
    bool reuse_staged = false;     // B is the block the previous product staged: no copy, no wait
};
// staging block of an n x n operand: rows padded to n + 2 elements (n % 8 == 0: the DMMA fragment reads of the four k-rows
// of a step then fall into disjoint bank groups); n = 64: TMA_LD
__host__ __device__ inline size_t tma_stage_bytes(int n) { return (size_t)n * (n + 2) * sizeof(cd); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_stage_init(TmaStage& ts, cd* buf, unsigned long long* bar) {
    ts.buf = buf;
    ts.bar = bar;
    ts.parity = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// C = op(A) * op(B)            (C must not alias A or B)
// FIXN: block edge known at compile time (64: the cfg4 instantiations of the sweep and lead kernels, which then hold only
// the n = 64 product routines) or 0 (any n; multiples of 8 from 16 on take the tensor-core path).
template <int FIXN = 0>
__device__ inline void cta_gemm_dmma(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn, bool accumulate,
                                     double cscale = 1.0, TmaStage* ts = nullptr);

template <int FIXN = 0>
__device__ inline void cta_gemm(cd* C, const cd* A, int opA, const cd* B, int opB, int n, TmaStage* ts = nullptr) {
    if constexpr (FIXN == 64) {
        cta_gemm_dmma<FIXN>(C, A, opA, B, opB, n, 1.0, false, 1.0, ts);
        return;
    }
    if ((n & 7) == 0 && n >= 16) {
        cta_gemm_dmma<FIXN>(C, A, opA, B, opB, n, 1.0, false, 1.0, ts);
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

// Compile-time size, no transposes (every product of the Sancho-Rubio loop and of the telescoped sweep): per-lane
// operand pointers are formed once per strip and advanced by constants, the k loop is fully unrolled, so the address
// arithmetic that makes up a third of the generic routine's instructions disappears.
template <int N>
__device__ inline void cta_gemm_dmma_nn(cd* C, const cd* __restrict__ A, const cd* __restrict__ B, double sgn, bool accumulate,
                                        double cscale) {
    constexpr int NT = 4;
    constexpr int TILES = N / 8;
    constexpr int STRIPS = (TILES + NT - 1) / NT;
    static_assert(TILES % NT == 0, "strips of four tiles");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int item = warp; item < TILES * STRIPS; item += nwarps) {
        const int tr = (item / STRIPS) << 3;
        const int tc0 = (item % STRIPS) * NT;
        double re[NT][2], im[NT][2];
#pragma unroll
        for (int u = 0; u < NT; ++u) re[u][0] = re[u][1] = im[u][0] = im[u][1] = 0.0;
        const cd* pa = A + (tr + g) * N + t;                     // A[tr+g][k+t], k += 4
        const cd* pb = B + t * N + (tc0 << 3) + g;               // B[k+t][(tc0+u)*8+g], k += 4
        // fragments of the next k-step are in flight while the DMMAs of the current one issue (two k-steps ahead was
        // measured too: no gain, the products are not bound by operand latency)
        cd a = pa[0];
        cd b[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) b[u] = pb[u * 8];
#pragma unroll
        for (int k = 0; k < N; k += 4) {
            const cd ac = a;
            cd bc[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) bc[u] = b[u];
            if (k + 4 < N) {
                a = pa[k + 4];
#pragma unroll
                for (int u = 0; u < NT; ++u) b[u] = pb[(k + 4) * N + u * 8];
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
        cd* pc = C + (tr + g) * N + (tc0 << 3) + 2 * t;
#pragma unroll
        for (int u = 0; u < NT; ++u) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cd v = mk(sgn * re[u][i], sgn * im[u][i]);
                if (accumulate) {
                    const cd c0 = pc[u * 8 + i];
                    v.x = fma(cscale, c0.x, v.x);
                    v.y = fma(cscale, c0.y, v.y);
                }
                pc[u * 8 + i] = v;
            }
        }
    }
    __syncthreads();
}

// The same product with B staged in shared memory by TMA (see TmaStage); 8 warps, one tile row each, all eight column
// tiles of the row in registers (32 accumulators), so every A fragment is fetched once.
template <int N>
__device__ inline void cta_gemm_dmma_nn_tma(cd* C, const cd* __restrict__ A, const cd* __restrict__ B, TmaStage& ts, double sgn,
                                            bool accumulate, double cscale, const GemmEpi epi = GemmEpi()) {
    static_assert(N == 64, "one tile row per warp, eight warps");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // B (global, possibly written by this CTA a moment ago) and the staging block (shared, last touched by ordinary
    // loads / stores) are about to be accessed through the async proxy: order the generic-proxy accesses first
    if (!epi.reuse_staged) asm volatile("fence.proxy.async;" ::: "memory");
    if (!epi.reuse_staged) __syncthreads();
    if (!epi.reuse_staged && warp == 0) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(ts.bar)), "r"((unsigned)(N * N * sizeof(cd)))
                         : "memory");
        __syncwarp();
#pragma unroll
        for (int r = lane; r < N; r += 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ts.buf + r * TMA_LD)),
                         "l"(B + r * N), "r"((unsigned)(N * sizeof(cd))), "r"(smem_u32(ts.bar))
                         : "memory");
    }
    double re[8][2], im[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) re[u][0] = re[u][1] = im[u][0] = im[u][1] = 0.0;
    const int tr = warp << 3;
    const cd* pa = A + (tr + g) * N + t;                         // A[tr+g][k+t], k += 4
    cd a = pa[0];
    if (!epi.reuse_staged) {   // wait for the staged block (bounded: a lost copy must not hang the GPU)
        unsigned done = 0;
        for (int spin = 0; !done; ++spin) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(smem_u32(ts.bar)), "r"(ts.parity)
                         : "memory");
            if (spin > (1 << 24)) __trap();
        }
        ts.parity ^= 1u;
    }
    const cd* pb = ts.buf + t * TMA_LD + g;                      // B[k+t][u*8+g]
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        const cd ac = a;
        if (k + 4 < N) a = pa[k + 4];
#pragma unroll
        for (int h = 0; h < 8; h += 4) {
            cd bc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) bc[u] = pb[k * TMA_LD + (h + u) * 8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dmma844(re[h + u], ac.x, bc[u].x);
                dmma844(im[h + u], ac.x, bc[u].y);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dmma844(re[h + u], -ac.y, bc[u].y);
                dmma844(im[h + u], ac.y, bc[u].x);
            }
        }
    }
    // Epilogue in batches of two tiles (four elements): the old values a batch reads (C for an accumulating product, the
    // second or third destination) are loaded before its first store, one destination at a time (sixteen registers of loads
    // in flight).  Warps issue in order, so a read-modify-write per element is a chain of sixteen L2 latencies (ncu: 11 % of
    // the samples of the Hermitian decimation sat on these lines).
    cd* pc = C + (tr + g) * N + 2 * t;
    auto add_into = [&](cd* D, const double scale) {           // D += scale * sgn * (A B)
#pragma unroll
        for (int u0 = 0; u0 < 8; u0 += 2) {
            cd o[2][2];
#pragma unroll
            for (int du = 0; du < 2; ++du)
#pragma unroll
                for (int i = 0; i < 2; ++i) o[du][i] = D[(u0 + du) * 8 + i];
#pragma unroll
            for (int du = 0; du < 2; ++du)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double f = scale * sgn;
                    D[(u0 + du) * 8 + i] = mk(fma(f, re[u0 + du][i], o[du][i].x), fma(f, im[u0 + du][i], o[du][i].y));
                }
        }
    };
    if (epi.C2) add_into(epi.C2 + (tr + g) * N + 2 * t, 1.0);
    if (epi.C3) add_into(epi.C3 + (tr + g) * N + 2 * t, epi.c3s);
#pragma unroll
    for (int u0 = 0; u0 < 8; u0 += 2) {
        cd o[2][2];
        if (accumulate) {
#pragma unroll
            for (int du = 0; du < 2; ++du)
#pragma unroll
                for (int i = 0; i < 2; ++i) o[du][i] = pc[(u0 + du) * 8 + i];
        }
#pragma unroll
        for (int du = 0; du < 2; ++du)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int u = u0 + du;
                cd v = mk(sgn * re[u][i], sgn * im[u][i]);
                if (accumulate) {
                    v.x = fma(cscale, o[du][i].x, v.x);
                    v.y = fma(cscale, o[du][i].y, v.y);
                }
                pc[u * 8 + i] = v;
                if (epi.tmax) {
                    const double m = norm2(v);
                    *epi.tmax = (m > *epi.tmax || !(m == m)) ? m : *epi.tmax;
                }
            }
    }
    __syncthreads();
}

// Block maximum of a per-thread value (NaN wins), broadcast to all threads.
__device__ inline double cta_reduce_max(double m) {
    __shared__ double s_rmax[32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = (o > m || !(o == o)) ? o : m;
    }
    if ((threadIdx.x & 31) == 0) s_rmax[threadIdx.x >> 5] = m;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) {
        const double o = s_rmax[w];
        r = (o > r || !(o == o)) ? o : r;
    }
    __syncthreads();
    return r;
}

// Where the run-time-n product stages its B operand (cfg4 workload, Sancho-Rubio leads + sweep, staged vs fragments from L2):
// n = 56: 46.0 vs 45.2 ms, 72: 101.8 vs 97.4, 96: 208.5 vs 213.3, 104: 271.5 vs 283.8 -- the copy sits at the head of every
// product with nothing to overlap it (one CTA per SM), so it only pays once a product is long enough; beyond n = 104 the
// staging block does not fit next to the kernels' static shared memory.
__host__ __device__ inline bool tma_stage_pays(int n) { return (n & 7) == 0 && n >= 96 && n <= TX_STAGE_SMEM_MAX_N; }

// Bring the n x n block B (global memory, n % 8 == 0) into the staging block with rows padded to n + 2 elements: one bulk
// copy per row issued by the lanes of warp 0, completion on the mbarrier; every thread of the CTA returns once the block
// has landed.  (Run-time-n counterpart of the prologue of cta_gemm_dmma_nn_tma.)
__device__ inline void tma_stage_load(TmaStage& ts, const cd* __restrict__ B, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = n + 2;
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(ts.bar)), "r"((unsigned)(n * n * sizeof(cd)))
                         : "memory");
        __syncwarp();
        for (int r = lane; r < n; r += 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ts.buf + r * ld)),
                         "l"(B + (long long)r * n), "r"((unsigned)(n * sizeof(cd))), "r"(smem_u32(ts.bar))
                         : "memory");
    }
    unsigned done = 0;
    for (int spin = 0; !done; ++spin) {   // bounded: a lost copy must not hang the GPU
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_u32(ts.bar)), "r"(ts.parity)
                     : "memory");
        if (spin > (1 << 24)) __trap();
    }
    ts.parity ^= 1u;
}

// Run-time size, no transposes (every product of the telescoped sweep and of the decimation at n != 64): per-lane operand
// pointers advanced by constants instead of re-deriving r * n + c behind an op switch for every fragment (the index
// arithmetic was 39 % of the instructions of the generic routine), and the column tiles of a tile row are split EVENLY over
// the ceil(tiles / 4) strips (n = 72: 9 tiles as 3 + 3 + 3 instead of 4 + 4 + 1).  More, narrower strips that would fill the
// warps better were measured and lose: every strip re-reads the A fragments of its tile row from L2, and operand fetch, not
// warp balance, binds these products (n = 104 with strips of 2: 389 ms against 308 ms for the cfg4 workload).
// One strip of NTT tiles (compile time: ncu showed the `u < nt` predicates around the DMMAs and the guarded prefetch as a
// third of the instructions, each predicate a BSSY / BSYNC / WARPSYNC around an mma.sync); the last k-step is peeled, so
// the loop body prefetches unconditionally.
template <int NTT>
__device__ __forceinline__ void gemm_rt_strip(cd* pc, const cd* __restrict__ pa, const cd* __restrict__ pb, int n, long long kstep,
                                              double sgn, bool accumulate, double cscale) {
    double re[NTT][2], im[NTT][2];
#pragma unroll
    for (int u = 0; u < NTT; ++u) re[u][0] = re[u][1] = im[u][0] = im[u][1] = 0.0;
    cd a = *pa;
    cd b[NTT];
#pragma unroll
    for (int u = 0; u < NTT; ++u) b[u] = pb[u * 8];
    auto kstep_mma = [&](const cd ac, const cd (&bc)[NTT]) {
#pragma unroll
        for (int u = 0; u < NTT; ++u) {
            dmma844(re[u], ac.x, bc[u].x);
            dmma844(im[u], ac.x, bc[u].y);
        }
#pragma unroll
        for (int u = 0; u < NTT; ++u) {
            dmma844(re[u], -ac.y, bc[u].y);
            dmma844(im[u], ac.y, bc[u].x);
        }
    };
    for (int k = 4; k < n; k += 4) {
        const cd ac = a;
        cd bc[NTT];
#pragma unroll
        for (int u = 0; u < NTT; ++u) bc[u] = b[u];
        pa += 4;
        pb += kstep;
        a = *pa;
#pragma unroll
        for (int u = 0; u < NTT; ++u) b[u] = pb[u * 8];
        kstep_mma(ac, bc);
    }
    kstep_mma(a, b);
#pragma unroll
    for (int u = 0; u < NTT; ++u) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            cd v = mk(sgn * re[u][i], sgn * im[u][i]);
            if (accumulate) {
                const cd c0 = pc[u * 8 + i];
                v.x = fma(cscale, c0.x, v.x);
                v.y = fma(cscale, c0.y, v.y);
            }
            pc[u * 8 + i] = v;
        }
    }
}   // (loads batched ahead of the stores were measured here too: the extra registers spill, 5 % slower)

// `ts`: stage B in shared memory first (TMA), so that only the A fragments -- one load per k-step and lane instead of up to
// five -- come from L2 (ncu on the n = 72 lead kernel: 23 % of all stall samples were DMMAs waiting for L2 operands).
__device__ inline void cta_gemm_dmma_nn_rt(cd* C, const cd* __restrict__ A, const cd* __restrict__ B, int n, double sgn,
                                           bool accumulate, double cscale, TmaStage* ts = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = n >> 3;
    const int strips = (tiles + 3) >> 2;
    const int base = tiles / strips, rem = tiles - base * strips;   // strip s holds base + (s < rem) tiles
    long long kstep = 4LL * n;
    int ldb = n;
    if (ts != nullptr) {
        tma_stage_load(*ts, B, n);
        B = ts->buf;
        ldb = n + 2;
        kstep = 4LL * ldb;
    }
    for (int item = warp; item < tiles * strips; item += nwarps) {
        const int trow = item / strips, s = item - trow * strips;
        const int tr = trow << 3;
        const int tc0 = s * base + min(s, rem);
        const int nt = base + (s < rem ? 1 : 0);
        const cd* pa = A + (long long)(tr + g) * n + t;              // A[tr+g][k+t], k += 4
        const cd* pb = B + (long long)t * ldb + (tc0 << 3) + g;      // B[k+t][(tc0+u)*8+g], k += 4
        cd* pc = C + (long long)(tr + g) * n + (tc0 << 3) + 2 * t;
        switch (nt) {                                                // warp-uniform
            case 4: gemm_rt_strip<4>(pc, pa, pb, n, kstep, sgn, accumulate, cscale); break;
            case 3: gemm_rt_strip<3>(pc, pa, pb, n, kstep, sgn, accumulate, cscale); break;
            case 2: gemm_rt_strip<2>(pc, pa, pb, n, kstep, sgn, accumulate, cscale); break;
            default: gemm_rt_strip<1>(pc, pa, pb, n, kstep, sgn, accumulate, cscale); break;
        }
    }
    __syncthreads();
}

// C = sgn*op(A)*op(B) (+ C if accumulate).  C must not alias A or B.  Requires n % 8 == 0.
// A warp works on a strip of up to four 8x8 output tiles of one tile row: the A fragment of a k-step is loaded
// once and reused for the four column tiles (16 DMMA per 5 fragment loads), and the loads of the next k-step are
// issued before the DMMAs of the current one (operands usually sit in L2/L1, not in shared memory).
template <int FIXN>
__device__ inline void cta_gemm_dmma(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn, bool accumulate,
                                     double cscale, TmaStage* ts) {
    if constexpr (FIXN == 64) {
        if (ts != nullptr && opA == OP_N && opB == OP_N && blockDim.x == 256) {
            cta_gemm_dmma_nn_tma<64>(C, A, B, *ts, sgn, accumulate, cscale);
            return;
        }
        if (opA == OP_N && opB == OP_N) {
            cta_gemm_dmma_nn<64>(C, A, B, sgn, accumulate, cscale);
            return;
        }
    } else {
        if (opA == OP_N && opB == OP_N) {
            cta_gemm_dmma_nn_rt(C, A, B, n, sgn, accumulate, cscale, ts);
            return;
        }
    }
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
                        v.x = fma(cscale, c0.x, v.x);
                        v.y = fma(cscale, c0.y, v.y);
                    }
                    C[idx] = v;
                }
            }
        }
    }
    __syncthreads();
}

// C = cscale * C + sgn * op(A) * op(B)
template <int FIXN = 0>
__device__ inline void cta_gemm_acc(cd* C, const cd* A, int opA, const cd* B, int opB, int n, double sgn, double cscale = 1.0,
                                    TmaStage* ts = nullptr) {
    if constexpr (FIXN == 64) {
        cta_gemm_dmma<FIXN>(C, A, opA, B, opB, n, sgn, true, cscale, ts);
        return;
    }
    if ((n & 7) == 0 && n >= 16) {
        cta_gemm_dmma<FIXN>(C, A, opA, B, opB, n, sgn, true, cscale, ts);
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
        v.x = fma(sgn, ar, cscale * v.x);
        v.y = fma(sgn, ai, cscale * v.y);
        C[idx] = v;
    }
    __syncthreads();
}

// The elementwise passes below run on blocks that usually sit in L2 / HBM (large-block workspaces): every thread
// issues FOUR independent loads before its first store, otherwise a pass is a chain of 16 dependent memory latencies.
__device__ inline void cta_copy(cd* dst, const cd* src, int count) {
    const int bd = blockDim.x;
    int idx = threadIdx.x;
    for (; idx + 3 * bd < count; idx += 4 * bd) {
        const cd v0 = src[idx], v1 = src[idx + bd], v2 = src[idx + 2 * bd], v3 = src[idx + 3 * bd];
        dst[idx] = v0;
        dst[idx + bd] = v1;
        dst[idx + 2 * bd] = v2;
        dst[idx + 3 * bd] = v3;
    }
    for (; idx < count; idx += bd) dst[idx] = src[idx];
    __syncthreads();
}

// A = z*I - H
__device__ inline void cta_z_minus(cd* A, cd z, const cd* H, int n) {
    const int bd = blockDim.x, count = n * n;
    int idx = threadIdx.x;
    auto one = [&](int i, cd hv) {
        const int r = i / n, c = i - r * n;
        A[i] = (r == c) ? mk(z.x - hv.x, z.y - hv.y) : mk(-hv.x, -hv.y);
    };
    for (; idx + 3 * bd < count; idx += 4 * bd) {
        const cd v0 = H[idx], v1 = H[idx + bd], v2 = H[idx + 2 * bd], v3 = H[idx + 3 * bd];
        one(idx, v0);
        one(idx + bd, v1);
        one(idx + 2 * bd, v2);
        one(idx + 3 * bd, v3);
    }
    for (; idx < count; idx += bd) one(idx, H[idx]);
    __syncthreads();
}

// A -= B
__device__ inline void cta_sub(cd* A, const cd* B, int count) {
    const int bd = blockDim.x;
    int idx = threadIdx.x;
    for (; idx + 3 * bd < count; idx += 4 * bd) {
        cd a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = A[idx + u * bd];
            b[u] = B[idx + u * bd];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) A[idx + u * bd] = mk(a[u].x - b[u].x, a[u].y - b[u].y);
    }
    for (; idx < count; idx += bd) {
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

// X[k][perm[c]] <- X[perm[k]][c] in place on an n x n block in shared memory (n * n <= 9 * blockDim.x): every thread
// reads its entries, barrier, writes them to their destinations.  Not inlined: the nine staged entries must not add to
// the register pressure of the kernels that call the inverse (they are capped at 128 registers for two CTAs per SM).
static __device__ __noinline__ void unpermute_in_place(cd* X, int n, const int* perm) {
    constexpr int MAXE = 9;
    cd v[MAXE];
#pragma unroll
    for (int u = 0; u < MAXE; ++u) {
        const int idx = threadIdx.x + u * blockDim.x;
        if (idx < n * n) {
            const int k = idx / n, c = idx - k * n;
            v[u] = X[perm[k] * n + c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXE; ++u) {
        const int idx = threadIdx.x + u * blockDim.x;
        if (idx < n * n) {
            const int k = idx / n, c = idx - k * n;
            X[k * n + perm[c]] = v[u];
        }
    }
    __syncthreads();
}

// 1/z with the reciprocal of |z|^2 from MUFU.RCP64H + two Newton steps (~1 ulp, no IEEE slow path): the reciprocal
// sits on the critical path of every elimination step of the blocked inverse.
__device__ __forceinline__ cd crecip_fast(cd z) {
    const double d = norm2(z);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return mk(z.x * r, -z.y * r);
}

// Blocked Gauss-Jordan inverse for n a multiple of 8 (n <= NMAX, 4 n <= blockDim.x: NMAX = 64 with 256 threads, NMAX = 128 with
// 512), X in global memory, S a staging block (shared memory; global for the largest blocks)
// of n rows with leading dimension n + 1 (conflict-free fragment accesses).  Per panel of 8 columns:
//   1. the 8 elimination steps on the n x 8 panel run in registers, two entries per thread (4n threads), with
//      IMPLICIT partial pivoting over the rows not used so far: no row is moved, the pivot row of step k is recorded
//      in perm[k]; a step is two block barriers (column + 64-bit max key to shared memory; raw pivot row).  (One warp
//      owning the whole panel needs no barriers but 64 dependent DFMAs per lane and step, which starve behind the
//      co-resident CTA's DMMA stream on the shared FP64 pipe: measured 3.5k clk per step inside the kernels.)
//   2. the accumulated transformation of the 8 steps is applied to all other columns as ONE rank-8 update on the FP64
//      tensor cores: with P the pivot rows, R = A[P, others] (copied out, then zeroed in place) and G the finished
//      panel,   A[:, others] += G R      (in-place Gauss-Jordan leaves in column k exactly the column of the
//      transformation that belongs to pivot row perm[k]).
// Afterwards S holds (PA)^-1 with logical row k in physical row perm[k]:  X[k][perm[c]] = S[perm[k]][c].
// Per column: 2 barriers and 2 complex FMAs per thread instead of 4-5 barriers and a full read-modify-write of the block.
// S == X: in place on a block that already lives in shared memory (leading dimension n); the permutation is then undone
// through registers (n <= 48 with 256 threads: at most 9 entries per thread).
template <int NMAX = 64>
__device__ inline void cta_inverse_blocked(cd* X, cd* S, int n, int* singular, const cd* zminus_src = nullptr, cd zval = mk(0.0, 0.0)) {
    static_assert(NMAX == 64 || NMAX == 128, "row index in the low 7 or 8 bits of the pivot key");
    constexpr unsigned RB = NMAX - 1 + NMAX;          // low bits that carry NMAX - row (1 .. NMAX)
    __shared__ cd s_R[8 * NMAX];
    __shared__ cd s_prow[8];
    __shared__ cd s_colk2[2 * NMAX];
    __shared__ unsigned s_key[32];
    __shared__ int s_perm[NMAX];
    const bool in_place = S == X;
    const int ld = in_place ? n : n + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (!in_place) {
        // staged copy; with zminus_src the block to invert is z - H, formed on the fly (X is then output only).  Four
        // independent loads per thread before the first store: the source sits in L2, and a load-store-load chain of 16
        // elements per thread was 6 of the 48 us of a 64 x 64 inverse.
        const cd* src = zminus_src ? zminus_src : X;
        const int bd = blockDim.x, count = n * n;
        auto put = [&](int idx, cd hv) {
            const int r = idx / n, c = idx - r * n;
            if (zminus_src) hv = (r == c) ? mk(zval.x - hv.x, zval.y - hv.y) : mk(-hv.x, -hv.y);
            S[r * ld + c] = hv;
        };
        int idx = threadIdx.x;
        for (; idx + 3 * bd < count; idx += 4 * bd) {
            const cd v0 = src[idx], v1 = src[idx + bd], v2 = src[idx + 2 * bd], v3 = src[idx + 3 * bd];
            put(idx, v0);
            put(idx + bd, v1);
            put(idx + 2 * bd, v2);
            put(idx + 3 * bd, v3);
        }
        for (; idx < count; idx += bd) put(idx, src[idx]);
        __syncthreads();
    }
    // panel ownership: thread t holds row r = t % n, columns 2cp and 2cp + 1 (cp = t / n < 4) of the current panel
    const int cp = threadIdx.x / n, r = threadIdx.x - cp * n;
    const bool act = cp < 4;
    bool used = false;                                 // my row has already served as a pivot row
    const int tiles = n >> 3;
    for (int kb = 0; kb < n; kb += 8) {
        cd a0 = act ? S[r * ld + kb + 2 * cp] : mk(0.0, 0.0);
        cd a1 = act ? S[r * ld + kb + 2 * cp + 1] : mk(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cd* colk = s_colk2 + (j & 1) * NMAX;
            // stage 1: column j of the panel to shared memory, pivot = largest |.|^2 among the unused rows.  The key is
            // the high word of |a|^2 (sign 0, exponent, 20 mantissa bits: monotone, NaN above everything) with 64 - row
            // in its low 7 bits: one REDUX per warp picks the pivot (13 mantissa bits decide — ample for partial
            // pivoting), ties go to the lowest row, 0 = not a candidate.
            unsigned key = 0u;
            if (act && cp == (j >> 1)) {
                const cd v = (j & 1) ? a1 : a0;
                colk[r] = v;
                if (!used) key = ((unsigned)__double2hiint(norm2(v)) & ~RB) | (unsigned)(NMAX - r);
            }
            key = __reduce_max_sync(0xffffffffu, key);
            if (lane == 0) s_key[warp] = key;
            __syncthreads();
            // one load + one REDUX instead of a loop of eight dependent loads: the elimination step is a latency chain, and
            // this alone took the 64 x 64 inverse from 71 to 61 us (tried and measured without gain: publishing every
            // candidate's reciprocal ahead of the barrier; letting the four owners of the pivot row scale it for everybody)
            key = __reduce_max_sync(0xffffffffu, lane < nwarps ? s_key[lane] : 0u);
            const int pr = NMAX - (int)(key & RB);
            const cd pivot = colk[pr];
            if (threadIdx.x == 0) {
                s_perm[kb + j] = pr;
                if (!(norm2(pivot) > 0.0)) *singular = 1;
            }
            if (act && r == pr) {                      // raw pivot row of the panel
                s_prow[2 * cp] = a0;
                s_prow[2 * cp + 1] = a1;
                used = true;
            }
            __syncthreads();
            // stage 2: every thread updates its two entries (the pivot row becomes row/pivot with 1/pivot in column j)
            if (act) {
                const cd inv = crecip_fast(pivot);
                const cd pw0 = (2 * cp == j) ? inv : s_prow[2 * cp] * inv;
                const cd pw1 = (2 * cp + 1 == j) ? inv : s_prow[2 * cp + 1] * inv;
                if (r == pr) {
                    a0 = pw0;
                    a1 = pw1;
                } else {
                    const cd f = colk[r];
                    cd v0 = (2 * cp == j) ? mk(0.0, 0.0) : a0;
                    cd v1 = (2 * cp + 1 == j) ? mk(0.0, 0.0) : a1;
                    cfms(v0.x, v0.y, f.x, f.y, pw0.x, pw0.y);
                    cfms(v1.x, v1.y, f.x, f.y, pw1.x, pw1.y);
                    a0 = v0;
                    a1 = v1;
                }
            }
        }
        if (act) {
            S[r * ld + kb + 2 * cp] = a0;
            S[r * ld + kb + 2 * cp + 1] = a1;
        }
        __syncthreads();
        // R = old pivot rows (all columns outside the panel), zeroed in place
        for (int idx = threadIdx.x; idx < 8 * n; idx += blockDim.x) {
            const int j = idx / n, c = idx - j * n;
            if (c < kb || c >= kb + 8) {
                const int pr = s_perm[kb + j];
                s_R[j * NMAX + c] = S[pr * ld + c];
                S[pr * ld + c] = mk(0.0, 0.0);
            }
        }
        __syncthreads();
        // A[:, others] += G R : a warp takes a tile row, the G fragments are loaded once for its column tiles
        {
            const int g = lane >> 2, t = lane & 3;
            for (int tr = warp; tr < tiles; tr += nwarps) {
                const cd ga = S[(tr * 8 + g) * ld + kb + t];
                const cd gb = S[(tr * 8 + g) * ld + kb + 4 + t];
                for (int tc = 0; tc < tiles; ++tc) {
                    if (tc * 8 == kb) continue;
                    const cd ra = s_R[t * NMAX + tc * 8 + g];
                    const cd rb = s_R[(4 + t) * NMAX + tc * 8 + g];
                    cd* cp = S + (tr * 8 + g) * ld + tc * 8 + 2 * t;
                    const cd c0 = cp[0], c1 = cp[1];
                    double re[2] = {c0.x, c1.x}, im[2] = {c0.y, c1.y};
                    dmma844(re, ga.x, ra.x);
                    dmma844(im, ga.x, ra.y);
                    dmma844(re, -ga.y, ra.y);
                    dmma844(im, ga.y, ra.x);
                    dmma844(re, gb.x, rb.x);
                    dmma844(im, gb.x, rb.y);
                    dmma844(re, -gb.y, rb.y);
                    dmma844(im, gb.y, rb.x);
                    cp[0] = mk(re[0], im[0]);
                    cp[1] = mk(re[1], im[1]);
                }
            }
        }
        __syncthreads();
    }
    // undo the implicit row permutation while copying out:  X[k][perm[c]] = S[perm[k]][c]
    if (in_place) {
        unpermute_in_place(X, n, s_perm);
        return;
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int k = idx / n, c = idx - k * n;
        X[k * n + s_perm[c]] = S[s_perm[k] * ld + c];
    }
    __syncthreads();
}

// Bytes of staging the inverse of a global-memory block needs (leading dimension n + 1 for the blocked path).
__host__ __device__ inline size_t inverse_stage_bytes(int n) { return (size_t)n * (n + 1) * sizeof(cd); }
// TX_STAGE_SMEM_MAX_N (tx_common.cuh): beyond it the staging block is a per-CTA block of the global workspace (L2-resident).
__host__ __device__ inline bool inverse_stage_in_smem(int n) { return n <= TX_STAGE_SMEM_MAX_N; }

// Inverse of X using the shared-memory staging block S (inverse_stage_bytes(n)) when X itself lives in global memory:
// every elimination step rewrites the block, which must not go through L2 sixty-four times.
template <int NMAX = 64>
__device__ inline void cta_inverse_staged(cd* X, cd* S, int n, cd* colk, int* ipiv, int* singular) {
    if (S == nullptr || S == X) {
        // block already in shared memory: blocked in place while the un-permutation fits the registers
        if ((n & 7) == 0 && n >= 16 && n * n <= 9 * (int)blockDim.x && 4 * n <= (int)blockDim.x)   // (implies n <= NMAX)
            cta_inverse_blocked<NMAX>(X, X, n, singular);
        else cta_inverse(X, n, colk, ipiv, singular);
        return;
    }
    // (NMAX = 64: every caller runs 256 threads; the 512-thread instantiation checks its thread count)
    if ((n & 7) == 0 && n >= 16 && n <= NMAX && (NMAX == 64 || 4 * n <= (int)blockDim.x)) {
        cta_inverse_blocked<NMAX>(X, S, n, singular);
        return;
    }
    cta_copy(S, X, n * n);
    cta_inverse(S, n, colk, ipiv, singular);
    cta_copy(X, S, n * n);
}

// max |A_ij|^2 over the block, broadcast to all threads
__device__ inline double cta_maxnorm2(const cd* A, int count, const cd* B = nullptr) {
    // max over A (and B when given: one pass, one pair of barriers for both)
    __shared__ double s_red[32];
    double m = 0.0;
    const int bd = blockDim.x;
    auto take = [&](cd x) {
        const double v = norm2(x);
        m = (v > m || !(v == v)) ? v : m;
    };
    for (int pass = 0; pass < (B ? 2 : 1); ++pass) {
        const cd* P = pass ? B : A;
        int idx = threadIdx.x;
        for (; idx + 3 * bd < count; idx += 4 * bd) {
            const cd v0 = P[idx], v1 = P[idx + bd], v2 = P[idx + 2 * bd], v3 = P[idx + 3 * bd];
            take(v0);
            take(v1);
            take(v2);
            take(v3);
        }
        for (; idx < count; idx += bd) take(P[idx]);
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
