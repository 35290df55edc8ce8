// Host-side helpers shared by the translation units of libtransport_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace tx {
// Records a printf-style message for tx_last_error() (thread local) and returns `code`.
int fail(int code, const char* fmt, ...);
// Counts kernel launches issued by the library (tx_launch_counter()).
void count_launch(int n = 1);
// bardeen.cu: 0 auto, 1 warp per state, 2 lane per state (tx_set_option(3, mode)).
int set_eig_mode(int mode);
// surface_gf.cu: TMA staging of the n = 64 products' B operand (tx_set_option(5, on)); read by rgf_generic.cu too.
int set_tma_mode(int on);
int tma_mode();
int set_hermitian_shortcut(int on);   // tx_set_option(6, .): two-product decimation for Hermitian hopping blocks
}  // namespace tx

#define TX_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::tx::fail(TX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                 \
    } while (0)
