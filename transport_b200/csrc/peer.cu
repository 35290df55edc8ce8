// Peer memory for the multi-GPU gather: result buffers allocated with cudaMalloc, exported with CUDA
// IPC and mapped by the other ranks of the node, so that the sweep epilogue can store the per-point totals straight
// into every rank's copy over NVLink (PeerGather in tx_common.cuh), plus a device-side barrier on flags living in the
// same kind of memory.  One process per GPU; handles travel over the host-side process group (64 opaque bytes).
#include <cuda_runtime.h>

#include <cstring>

#include "tx_common.cuh"
#include "tx_host.h"

namespace tx {

struct PeerFlags {
    int* flags[8];   // flags[q] = rank q's flag array (8 ints), mapped into this process
};

// Rank `rank` tells every peer that its earlier work on this stream is complete (kernel boundaries order the peer
// stores of the sweep before this kernel; the fence + system-scope store publish the flag after them) and waits
// until every peer has said the same.  A peer that never arrives trips the timeout instead of hanging the GPU.
__global__ void peer_barrier_kernel(PeerFlags pf, int n, int rank, int epoch, long long timeout_cycles, int* timed_out) {
    const int q = threadIdx.x;
    if (q >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pf.flags[q] + rank), "r"(epoch) : "memory");
    const int* mine = pf.flags[rank] + q;
    const long long t0 = clock64();
    int v;
    do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (v - epoch >= 0) break;
        if (clock64() - t0 > timeout_cycles) {
            *timed_out = 1;
            break;
        }
    } while (true);
}

}  // namespace tx

using namespace tx;

extern "C" {

int tx_peer_alloc(long long bytes, void** ptr) {
    if (!ptr || bytes <= 0) return fail(TX_ERR_ARG, "bad arguments");
    void* p = nullptr;
    TX_CUDA(cudaMalloc(&p, (size_t)bytes));
    TX_CUDA(cudaMemset(p, 0, (size_t)bytes));
    TX_CUDA(cudaDeviceSynchronize());
    *ptr = p;
    return TX_OK;
}

int tx_peer_free(void* ptr) {
    if (ptr) TX_CUDA(cudaFree(ptr));
    return TX_OK;
}

int tx_peer_export(const void* ptr, unsigned char* handle64) {
    if (!ptr || !handle64) return fail(TX_ERR_ARG, "null pointer argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    TX_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle64, &h, 64);
    return TX_OK;
}

int tx_peer_open(const unsigned char* handle64, void** ptr) {
    if (!ptr || !handle64) return fail(TX_ERR_ARG, "null pointer argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    TX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return TX_OK;
}

int tx_peer_close(void* ptr) {
    if (ptr) TX_CUDA(cudaIpcCloseMemHandle(ptr));
    return TX_OK;
}

int tx_peer_barrier_dev(int* const* flag_peers, int n_peers, int rank, int epoch, int* timed_out, void* stream) {
    if (!flag_peers || !timed_out || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers)
        return fail(TX_ERR_ARG, "bad arguments (1 <= n_peers <= 8)");
    PeerFlags pf{};
    for (int q = 0; q < n_peers; ++q) {
        if (!flag_peers[q]) return fail(TX_ERR_ARG, "null flag pointer for rank %d", q);
        pf.flags[q] = flag_peers[q];
    }
    // ~2 s at 2 GHz: far beyond any legitimate skew between ranks of one step
    peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pf, n_peers, rank, epoch, 4000000000LL, timed_out);
    count_launch();
    TX_CUDA(cudaGetLastError());
    return TX_OK;
}

}  // extern "C"
