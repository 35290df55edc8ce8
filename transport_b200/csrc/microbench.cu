// Roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 FMA-pipe peak (DFMA),
// the legacy FP64 tensor path (mma.sync m8n8k4 f64), and the two exchange primitives the
// small-block sweep leans on (SHFL and partially-broadcast LDS.128).  Register-resident loops, no
// global traffic; timed with CUDA events.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/transport_b200.h"

namespace {

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) dfma_loop(double* out, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_loop(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// DMMA with CH independent accumulator chains per warp (latency / occupancy scan)
template <int CH>
__global__ void __launch_bounds__(128) dmma_chain_loop(double* out, double a, double b) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 16 / CH; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(a), "d"(b));
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// DFMA and DMMA issued back to back from the same warp: do the two FP64 paths share one pipe?
__global__ void __launch_bounds__(256) dfma_dmma_loop(double* out, double a, double b) {
    double c[4][2];
    double x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
            x[2 * i] = fma(x[2 * i], b, a);
            x[2 * i + 1] = fma(x[2 * i + 1], b, a);
            x[(2 * i + 2) & 7] = fma(x[(2 * i + 2) & 7], b, a);
            x[(2 * i + 3) & 7] = fma(x[(2 * i + 3) & 7], b, a);
            x[(2 * i + 4) & 7] = fma(x[(2 * i + 4) & 7], b, a);
            x[(2 * i + 5) & 7] = fma(x[(2 * i + 5) & 7], b, a);
            x[(2 * i + 6) & 7] = fma(x[(2 * i + 6) & 7], b, a);
            x[(2 * i + 7) & 7] = fma(x[(2 * i + 7) & 7], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) shfl_loop(double* out) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1) + 1;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345u) out[0] = s;
}

// LDS.128 where each group of 4 lanes reads one 16-byte word and the 8 groups of a warp read 8
// consecutive words (one 128-byte line) — the access shape of the pivot-row / inverse broadcast.
__global__ void __launch_bounds__(256) lds_bcast_loop(double* out, int stride) {
    __shared__ double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    const int grp = (threadIdx.x & 31) >> 2;
    const int base = (threadIdx.x >> 5) * 64;
    double2 acc = make_double2(0, 0);
    int idx = base + grp;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double2 v = buf[(idx + i * stride) & 2047];
            acc.x += v.x;
            acc.y += v.y;
        }
        idx += (int)acc.x & 8;
    }
    if (acc.x == 123.456) out[0] = acc.y;
}

// DFMA and broadcast LDS.128 interleaved 8:1, the mix of the sweep's inner loops.
__global__ void __launch_bounds__(256) mix_loop(double* out, double a, double b) {
    __shared__ double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    __syncthreads();
    const int grp = (threadIdx.x & 31) >> 2;
    int idx = (threadIdx.x >> 5) * 64 + grp;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 v = buf[(idx + i * 8) & 2047];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                x[4 * i + k] = fma(x[4 * i + k], v.x, v.y);
                x[(4 * i + k + 8) & 15] = fma(x[(4 * i + k + 8) & 15], v.y, v.x);
            }
        }
        idx += 32;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s + b;
}

thread_local char g_mb_err[256];

template <typename F>
int time_kernel(F launch, double* ms_out) {
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return TX_ERR_CUDA;
    launch();   // warm-up
    launch();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return TX_ERR_CUDA;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return TX_ERR_CUDA;
    *ms_out = best;
    return TX_OK;
}

int sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

extern "C" {

int tx_measure_fp64_peak(double* tflops, double* milliseconds) {
    if (!tflops) return TX_ERR_ARG;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return TX_ERR_CUDA;
    const int blocks = sm_count() * 8, threads = 256;
    double ms = 0;
    int rc = time_kernel([&] { dfma_loop<<<blocks, threads>>>(d, 1.0000001, 0.9999999); }, &ms);
    cudaFree(d);
    if (rc) return rc;
    const double flops = 2.0 * 16.0 * ITERS * (double)blocks * threads;
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (milliseconds) *milliseconds = ms;
    return TX_OK;
}

int tx_measure_dmma_peak(double* tflops, double* milliseconds) {
    if (!tflops) return TX_ERR_ARG;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return TX_ERR_CUDA;
    const int blocks = sm_count() * 8, threads = 256;
    double ms = 0;
    int rc = time_kernel([&] { dmma_loop<<<blocks, threads>>>(d, 1.0000001, 0.9999999); }, &ms);
    cudaFree(d);
    if (rc) return rc;
    // one m8n8k4 = 8*8*4 FMA per warp
    const double flops = 2.0 * 256.0 * 8.0 * ITERS * (double)blocks * (threads / 32);
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (milliseconds) *milliseconds = ms;
    return TX_OK;
}

/* which: 0 = SHFL.b32 warp-instructions/s (1e9), 1 = LDS.128 group-broadcast warp-instructions/s (1e9,
 * stride 8 = one 128-byte line per instruction), 2 = same with a 2-way bank conflict (stride 4),
 * 3 = DFMA:LDS 8:1 mix in TFLOP/s.  Internal, used by scripts/microbench.py. */
int tx_microbench(int which, double* value) {
    if (!value) return TX_ERR_ARG;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return TX_ERR_CUDA;
    const int blocks = sm_count() * 8, threads = 256;
    double ms = 0;
    int rc = TX_ERR_ARG;
    const double warps = (double)blocks * (threads / 32);
    if (which == 0) {
        rc = time_kernel([&] { shfl_loop<<<blocks, threads>>>(d); }, &ms);
        *value = 8.0 * ITERS * warps / (ms * 1e-3) / 1e9;
    } else if (which == 1 || which == 2) {
        const int stride = which == 1 ? 8 : 4;
        rc = time_kernel([&] { lds_bcast_loop<<<blocks, threads>>>(d, stride); }, &ms);
        *value = 8.0 * ITERS * warps / (ms * 1e-3) / 1e9;
    } else if (which == 4) {
        // 4 DMMA (256 FMA each per warp) + 32 DFMA (32 FMA each per warp) per iteration
        rc = time_kernel([&] { dfma_dmma_loop<<<blocks, threads>>>(d, 1.0000001, 0.9999999); }, &ms);
        *value = 2.0 * (4.0 * 256.0 + 32.0 * 32.0) * ITERS * warps / (ms * 1e-3) / 1e12;
    } else if (which >= 10 && which < 20) {
        // DFMA peak at reduced occupancy: (which-10+1) warps per scheduler, 16 independent chains per thread
        const int wps = which - 10 + 1;
        const int thr = 128, blk = sm_count() * wps;       // wps blocks of 4 warps per SM
        rc = time_kernel([&] { dfma_loop<<<blk, thr>>>(d, 1.0000001, 0.9999999); }, &ms);
        *value = 2.0 * 16.0 * ITERS * (double)blk * thr / (ms * 1e-3) / 1e12;
    } else if (which >= 100 && which < 1000) {
        // DMMA TFLOP/s with (which/100) warps per scheduler and (which%100) independent chains per warp (16 DMMA per iteration)
        const int wps = which / 100, ch = which % 100;
        const int thr = 128, blk = sm_count() * wps;
        auto go = [&](auto kern) { return time_kernel([&] { kern<<<blk, thr>>>(d, 1.0000001, 0.9999999); }, &ms); };
        if (ch == 1) rc = go(dmma_chain_loop<1>);
        else if (ch == 2) rc = go(dmma_chain_loop<2>);
        else if (ch == 4) rc = go(dmma_chain_loop<4>);
        else if (ch == 8) rc = go(dmma_chain_loop<8>);
        else if (ch == 16) rc = go(dmma_chain_loop<16>);
        *value = 2.0 * 256.0 * 16.0 * ITERS * (double)blk * (thr / 32) / (ms * 1e-3) / 1e12;
    } else if (which == 3) {
        rc = time_kernel([&] { mix_loop<<<blocks, threads>>>(d, 1.0000001, 0.5); }, &ms);
        *value = 2.0 * 32.0 * ITERS * warps * 32 / (ms * 1e-3) / 1e12;
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"
