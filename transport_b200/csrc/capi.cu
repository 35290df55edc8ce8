// C ABI of libtransport_b200.so (declared in include/transport_b200.h).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "rgf_generic.h"
#include "rgf_n8.cuh"
#include "rgf_n8t.cuh"
#include "rgf_small.cuh"
#include "tx_host.h"

using namespace tx;

namespace tx {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
}  // namespace tx

namespace {

// ---- grow-only device buffer ---------------------------------------------------------------
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return TX_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) return fail(TX_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return TX_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

// Tuning state (tx_set_variant / tx_set_option set the process defaults; a context carries its own copy).
struct Options {
    int variant = 13;       // 13 telescoped sweep, 6 plain sweep (DMMA products), 0 register-resident template
    int chunk_kmax = 32;    // longest telescoped chunk of rgf_sweep_n8t
    int chunk_gexp = 4;     // log2 of the growth bound that closes a chunk early
    bool diag_tail = true;  // structure scan of the on-site blocks
};
Options g_default_opts;

}  // namespace

// A context owns everything a sweep call needs besides its arguments: the device workspace of the host-pointer
// entry, the scratch of the device-pointer entry (expanded affine families, site classes), streams, events and
// the tuning options.  Calls on one context are serialised (host mutex; scratch reuse is ordered across streams
// by an event); calls on distinct contexts are independent, so threads / streams that want concurrency create
// their own.  Entry points without a context argument use the per-device default context.
struct tx_context {
    int device = 0;
    std::mutex mu;
    Options opt;
    DevBuf h, h1, par, tnn, E, sigL, sigR, src, R, T, resid, ttot, caroli, spin, big, flag;
    DevBuf expand, nondiag, restart;      // restart: g_{b+1} of the chunk in flight, per resident warp (rgf_sweep_n8t)
    cudaStream_t own_stream = nullptr;    // created on first use
    cudaStream_t user_stream = nullptr;   // tx_context_set_stream
    bool has_user_stream = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t scratch_ev = nullptr;     // recorded after the last kernels that read expand / nondiag
    cudaStream_t scratch_stream = nullptr;
    bool scratch_used = false;
    bool follows_defaults = false;        // default contexts track tx_set_variant / tx_set_option
    PeerGather gather{};                  // tx_context_set_gather (n = 0: none)
    double lead_imE = 0.0, lead_tol = 1e-14;   // tx_context_set_lead_params (TX_LEAD_SANCHO / TX_LEAD_SEPARABLE)
    int lead_max_iter = 200;
    double last_lead_iters = 0.0;         // mean Sancho-Rubio iterations of the last host sweep (both leads)
    DevBuf nit, okb;
};

namespace {

std::mutex g_ctx_mutex;
tx_context* g_default_ctx[16] = {nullptr};

int default_context(tx_context** out) {
    int dev = 0;
    TX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    tx_context*& c = g_default_ctx[dev & 15];
    if (!c) {
        c = new tx_context();
        c->device = dev;
        c->follows_defaults = true;
    }
    *out = c;
    return TX_OK;
}

int round_up_block(int n) {
    if (n <= 1) return 1;
    if (n <= 2) return 2;
    if (n <= 4) return 4;
    if (n <= 8) return 8;
    if (n <= 16) return 16;
    return 0;
}

template <int N, int LP, int HOP, int LEAD, int MINB, int PIVX = 0>
int launch_small(const SweepParams& p, cudaStream_t st) {
    using Cfg = SmallCfg<N, LP, HOP>;
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * Cfg::PER_WARP_BYTES;
    static bool attr_done[16] = {false};
    int dev = 0;
    TX_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 15]) {
        TX_CUDA(cudaFuncSetAttribute(rgf_sweep_small<N, LP, HOP, LEAD, MINB, PIVX>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[dev & 15] = true;
    }
    const long long per_block = (long long)WARPS * Cfg::GPW;
    const long long blocks = (p.n_pts + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return fail(TX_ERR_ARG, "too many points for one launch");
    rgf_sweep_small<N, LP, HOP, LEAD, MINB, PIVX><<<(unsigned)blocks, WARPS * 32, smem, st>>>(p);
    count_launch();
    TX_CUDA(cudaGetLastError());
    return TX_OK;
}

// h0 + J*h1 families are expanded once into the context's grow-only buffer (8 KB per Hamiltonian at n = M = 8).
int materialise_affine(tx_context& cx, SweepParams& p, cudaStream_t st) {
    if (!p.h1) return TX_OK;
    const long long per_ham = (long long)p.M * p.n * p.n;
    const long long n_ham = p.n_pts / p.nE;
    const long long total = per_ham * n_ham;
    DevBuf& buf = cx.expand;
    if ((size_t)total * sizeof(cd) > buf.cap) TX_CUDA(cudaDeviceSynchronize());   // a regrow frees the old buffer
    int rc = buf.ensure((size_t)total * sizeof(cd));
    if (rc) return rc;
    affine_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.h, p.h1, p.params, (cd*)buf.ptr, per_ham, total);
    count_launch();
    TX_CUDA(cudaGetLastError());
    p.h = (const cd*)buf.ptr;
    p.h_stride = per_ham;
    p.h1 = nullptr;
    p.params = nullptr;
    cx.scratch_used = true;
    return TX_OK;
}

// Structure scan of the on-site blocks (structure_scan_kernel): per Hamiltonian the last non-diagonal site and
// per site a class (2 scalar multiple of I, 1 diagonal, 0 general), plus the batch-wide count of diagonal interior
// sites.  Must run before materialise_affine.
int scan_structure(tx_context& cx, SweepParams& p, cudaStream_t st, bool want_tail) {
    const int n_hs = (p.h_stride == 0) ? 1 : (int)(p.n_pts / p.nE);
    DevBuf& nb = cx.nondiag;
    const size_t need = ((size_t)n_hs * (1 + (size_t)p.M) + 1) * sizeof(int);
    if (need > nb.cap) TX_CUDA(cudaDeviceSynchronize());
    int rc = nb.ensure(need);
    if (rc) return rc;
    int* nd = (int*)nb.ptr;
    int* cl = nd + n_hs;
    const long long ncl = (long long)n_hs * p.M;
    int* cnt = cl + ncl;
    // cnt = 0 and nd = -1 are set by one memset each; the classes are written (not accumulated) by the scan
    TX_CUDA(cudaMemsetAsync(nd, 0xFF, (size_t)n_hs * sizeof(int), st));
    TX_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
    // one warp per (Hamiltonian, site)
    const long long warps = ncl;
    structure_scan_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(p.h, p.h1, n_hs, p.M, p.n, nd, cl, cnt);
    TX_CUDA(cudaGetLastError());
    count_launch(1);
    p.class_count = cnt;
    p.class_sites = (long long)n_hs * (p.M > 2 ? p.M - 2 : 0);
    if (want_tail) {
        p.nondiag_max = nd;
        p.nondiag_stride = (p.h_stride == 0) ? 0 : 1;
    }
    p.site_class = cl;
    p.class_stride = (p.h_stride == 0) ? 0 : p.M;
    cx.scratch_used = true;
    return TX_OK;
}

// Telescoped sweep (rgf_n8t.cuh): the default for 5 <= n <= 8 with scalar hopping.  Persistent grid: as many CTAs
// as are resident on the device, each warp striding over its items.
template <int LEAD, int MINB, bool FULL, bool DIAG, bool GUARD>
int launch_n8t_guard(tx_context& cx, const SweepParams& p, cudaStream_t st) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * N8TCfg::PER_WARP_BYTES;
    static int resident[16] = {0};
    int dev = 0;
    TX_CUDA(cudaGetDevice(&dev));
    if (!resident[dev & 15]) {
        TX_CUDA(cudaFuncSetAttribute(rgf_sweep_n8t<LEAD, MINB, FULL, DIAG, GUARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        TX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rgf_sweep_n8t<LEAD, MINB, FULL, DIAG, GUARD>, WARPS * 32, smem));
        TX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        resident[dev & 15] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    // a warp = eight consecutive energies of one Hamiltonian (energy axis padded to a multiple of eight)
    const int warps_per_ham = (p.nE + 7) / 8;
    const long long n_warps = (p.n_pts / p.nE) * warps_per_ham;
    long long blocks = (n_warps + WARPS - 1) / WARPS;
    if (blocks > resident[dev & 15]) blocks = resident[dev & 15];
    if (GUARD) {
        const size_t scratch = (size_t)resident[dev & 15] * WARPS * 8 * 64 * sizeof(cd);
        if (scratch > cx.restart.cap) TX_CUDA(cudaDeviceSynchronize());
        if (int rc = cx.restart.ensure(scratch)) return rc;
        cx.scratch_used = true;
    }
    rgf_sweep_n8t<LEAD, MINB, FULL, DIAG, GUARD><<<(unsigned)blocks, WARPS * 32, smem, st>>>(p, cx.opt.chunk_kmax, 2 * cx.opt.chunk_gexp, warps_per_ham, n_warps,
                                                                                  reinterpret_cast<double2*>(cx.restart.ptr));
    count_launch();
    TX_CUDA(cudaGetLastError());
    return TX_OK;
}

// chains longer than one chunk get the conditioning guard (rgf_n8t.cuh)
template <int LEAD, int MINB, bool FULL, bool DIAG>
int launch_n8t_inst(tx_context& cx, const SweepParams& p, cudaStream_t st) {
    if (p.M > cx.opt.chunk_kmax) return launch_n8t_guard<LEAD, MINB, FULL, DIAG, true>(cx, p, st);
    return launch_n8t_guard<LEAD, MINB, FULL, DIAG, false>(cx, p, st);
}

// Once per sweep (not per chunk of the Hamiltonian axis): structure scan of the on-site blocks and expansion of an
// affine family, both into the context's scratch; `p` is rewritten to point at the results.
int prepare_n8(tx_context& cx, SweepParams& p, cudaStream_t st, bool want_tail) {
    p.nondiag_max = nullptr;
    p.nondiag_stride = 0;
    p.site_class = nullptr;
    p.class_stride = 0;
    p.class_count = nullptr;
    p.class_sites = 0;
    if (cx.opt.diag_tail) {
        int rc = scan_structure(cx, p, st, want_tail);
        if (rc) return rc;
    }
    return materialise_affine(cx, p, st);
}

template <int LEAD, int MINB, bool FULL>
int launch_n8t_full(tx_context& cx, const SweepParams& p, cudaStream_t st) {
    if (!cx.opt.diag_tail) return launch_n8t_inst<LEAD, MINB, FULL, false>(cx, p, st);
    // both instantiations; the one the device-side site statistics do not select returns at once
    int rc = launch_n8t_inst<LEAD, MINB, FULL, true>(cx, p, st);
    if (rc) return rc;
    return launch_n8t_inst<LEAD, MINB, FULL, false>(cx, p, st);
}

template <int LEAD, int MINB>
int launch_n8t(tx_context& cx, const SweepParams& p, cudaStream_t st) {
    return p.n == 8 ? launch_n8t_full<LEAD, MINB, true>(cx, p, st) : launch_n8t_full<LEAD, MINB, false>(cx, p, st);
}

template <int LEAD, int MINB, int PIVX, int DMB = 1>
int launch_n8(tx_context& cx, const SweepParams& p, cudaStream_t st) {
    constexpr int WARPS = 4;
    (void)cx;
    const size_t smem = (size_t)WARPS * N8Cfg::PER_WARP_BYTES;
    static bool attr_done[16] = {false};
    int dev = 0;
    TX_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 15]) {
        TX_CUDA(cudaFuncSetAttribute(rgf_sweep_n8<LEAD, MINB, PIVX, DMB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[dev & 15] = true;
    }
    const long long per_block = (long long)WARPS * 8;
    const long long blocks = (p.n_pts + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return fail(TX_ERR_ARG, "too many points for one launch");
    rgf_sweep_n8<LEAD, MINB, PIVX, DMB><<<(unsigned)blocks, WARPS * 32, smem, st>>>(p);
    count_launch();
    TX_CUDA(cudaGetLastError());
    return TX_OK;
}

template <int N, int LP, int MB_SCALAR, int MB_GENERAL>
int dispatch_modes(const SweepParams& p, int hop, int lead, cudaStream_t st) {
    if (hop == HOP_SCALAR) {
        if (lead == LEAD_CLOSED) return launch_small<N, LP, HOP_SCALAR, LEAD_CLOSED, MB_SCALAR>(p, st);
        return launch_small<N, LP, HOP_SCALAR, LEAD_TABLE, MB_SCALAR>(p, st);
    }
    if (lead == LEAD_CLOSED) return launch_small<N, LP, HOP_GENERAL, LEAD_CLOSED, MB_GENERAL>(p, st);
    return launch_small<N, LP, HOP_GENERAL, LEAD_TABLE, MB_GENERAL>(p, st);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Output arrays of a sweep (any may be null; R and T go together).
struct SweepOut {
    double *R, *T, *resid, *t_total, *caroli, *spin;
    int n_up, spin_n;
};

int check_spin(const SweepOut& o, int n, bool has_gather = false) {
    if ((o.R == nullptr) != (o.T == nullptr)) return fail(TX_ERR_ARG, "R and T must be given together (or both NULL)");
    if (o.spin) {
        if (o.spin_n != 2 && o.spin_n != 4) return fail(TX_ERR_ARG, "spin_n must be 2 or 4");
        if (o.n_up < 0 || o.n_up > n) return fail(TX_ERR_ARG, "n_up=%d outside 0..n", o.n_up);
    }
    if (!o.R && !o.resid && !o.t_total && !o.caroli && !o.spin && !has_gather) return fail(TX_ERR_ARG, "no output array given");
    return TX_OK;
}

}  // namespace

extern "C" {

const char* tx_version(void) { return "transport_b200 0.2 (sm_100a)"; }
const char* tx_last_error(void) { return g_err; }

int tx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int tx_set_device(int device) {
    TX_CUDA(cudaSetDevice(device));
    return TX_OK;
}

int tx_device_synchronize(void) {
    TX_CUDA(cudaDeviceSynchronize());
    return TX_OK;
}

long long tx_launch_counter(void) { return g_launches.load(); }

static int set_option_on(Options& o, int option, int value) {
    if (option == 0) {
        o.diag_tail = value != 0;
        return TX_OK;
    }
    if (option == 1) {
        if (value < 1 || value > 64) return fail(TX_ERR_ARG, "chunk length must be 1..64");
        o.chunk_kmax = value;
        return TX_OK;
    }
    if (option == 2) {
        if (value < 0 || value > 500) return fail(TX_ERR_ARG, "growth exponent must be 0..500");
        o.chunk_gexp = value;
        return TX_OK;
    }
    if (option == 4) {
        if (value != 0 && value != 6 && value != 13) return fail(TX_ERR_ARG, "variant must be 13, 6 or 0");
        o.variant = value;
        return TX_OK;
    }
    return fail(TX_ERR_ARG, "unknown option %d", option);
}

int tx_set_option(int option, int value) {
    if (option == 3) return set_eig_mode(value);
    if (option == 5) return set_tma_mode(value);
    if (option == 6) return set_hermitian_shortcut(value);
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    return set_option_on(g_default_opts, option, value);
}

int tx_set_variant(int v) { return tx_set_option(4, v); }

int tx_context_create(int device, tx_context** out) {
    if (!out) return fail(TX_ERR_ARG, "null pointer argument");
    int ndev = tx_device_count();
    if (ndev < 1) return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TX_ERR_ARG, "device %d outside 0..%d", device, ndev - 1);
    tx_context* c = new tx_context();
    c->device = device;
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        c->opt = g_default_opts;
    }
    *out = c;
    return TX_OK;
}

int tx_context_destroy(tx_context* c) {
    if (!c) return TX_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->scratch_ev) cudaEventSynchronize(c->scratch_ev);
    DevBuf* bufs[] = {&c->h, &c->h1, &c->par, &c->tnn, &c->E, &c->sigL, &c->sigR, &c->src, &c->R, &c->T, &c->resid,
                      &c->ttot, &c->caroli, &c->spin, &c->big, &c->flag, &c->expand, &c->nondiag, &c->restart, &c->nit, &c->okb};
    for (DevBuf* b : bufs) b->release();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->scratch_ev) cudaEventDestroy(c->scratch_ev);
    cudaSetDevice(prev);
    delete c;
    return TX_OK;
}

int tx_context_set_stream(tx_context* c, void* stream, int use_stream) {
    if (!c) return fail(TX_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lock(c->mu);
    c->user_stream = reinterpret_cast<cudaStream_t>(stream);
    c->has_user_stream = use_stream != 0;
    return TX_OK;
}

int tx_context_set_lead_params(tx_context* c, double imE, double conv_tol, int max_iter) {
    if (!(conv_tol >= 0.0) || max_iter < 1) return fail(TX_ERR_ARG, "bad lead parameters");
    if (!c) {
        if (tx_device_count() < 1) return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
        if (int rc = default_context(&c)) return rc;
    }
    std::lock_guard<std::mutex> lock(c->mu);
    c->lead_imE = imE;
    c->lead_tol = conv_tol;
    c->lead_max_iter = max_iter;
    return TX_OK;
}

int tx_context_last_lead_iterations(tx_context* c, double* mean_iterations) {
    if (!mean_iterations) return fail(TX_ERR_ARG, "null pointer argument");
    if (!c) {
        if (tx_device_count() < 1) return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
        if (int rc = default_context(&c)) return rc;
    }
    std::lock_guard<std::mutex> lock(c->mu);
    *mean_iterations = c->last_lead_iters;
    return TX_OK;
}

int tx_context_set_gather(tx_context* c, double* const* peers, int n_peers, long long offset) {
    if (!c) return fail(TX_ERR_ARG, "null context");
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !peers) || offset < 0) return fail(TX_ERR_ARG, "bad gather spec (n_peers <= 8)");
    std::lock_guard<std::mutex> lock(c->mu);
    c->gather = PeerGather{};
    for (int q = 0; q < n_peers; ++q) {
        if (!peers[q]) return fail(TX_ERR_ARG, "null peer pointer %d", q);
        c->gather.dst[q] = peers[q];
    }
    c->gather.n = n_peers;
    c->gather.off = offset;
    return TX_OK;
}

int tx_context_set_option(tx_context* c, int option, int value) {
    if (!c) return fail(TX_ERR_ARG, "null context");
    if (option == 3) return fail(TX_ERR_ARG, "option 3 (eigenvalue kernel) is process-wide: use tx_set_option");
    std::lock_guard<std::mutex> lock(c->mu);
    c->follows_defaults = false;
    return set_option_on(c->opt, option, value);
}

// Unit-test hook: batched inverse of n x n blocks with the sweep's own device routine (host pointers).
int tx_debug_invert(const tx_cplx* in, tx_cplx* out, int count, int n) {
    if (!in || !out || count < 1 || n < 1 || n > 16) return fail(TX_ERR_ARG, "bad arguments");
    if (tx_device_count() < 1) return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    const size_t bytes = (size_t)count * n * n * sizeof(cd);
    cd *din = nullptr, *dout = nullptr;
    int* dflag = nullptr;
    TX_CUDA(cudaMalloc(&din, bytes));
    TX_CUDA(cudaMalloc(&dout, bytes));
    TX_CUDA(cudaMalloc(&dflag, sizeof(int)));
    TX_CUDA(cudaMemcpy(din, in, bytes, cudaMemcpyHostToDevice));
    TX_CUDA(cudaMemset(dflag, 0, sizeof(int)));
    const int N = round_up_block(n);
    auto go = [&](auto kern, int NN, int LP) -> int {
        const int GPW = 32 / LP;
        const size_t smem = (size_t)4 * (2 * NN + NN * NN) * GPW * 16;
        TX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long per_block = 4LL * GPW;
        kern<<<(unsigned)((count + per_block - 1) / per_block), 128, smem>>>(din, dout, count, n, dflag);
        TX_CUDA(cudaGetLastError());
        return TX_OK;
    };
    int rc = TX_OK;
    switch (N) {
        case 1: rc = go(invert_batch_small<1, 1>, 1, 1); break;
        case 2: rc = go(invert_batch_small<2, 1>, 2, 1); break;
        case 4: rc = go(invert_batch_small<4, 2>, 4, 2); break;
        case 8: rc = go(invert_batch_small<8, 4>, 8, 4); break;
        case 16: rc = go(invert_batch_small<16, 16>, 16, 16); break;
    }
    if (rc == TX_OK) {
        TX_CUDA(cudaDeviceSynchronize());
        TX_CUDA(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    }
    cudaFree(din);
    cudaFree(dout);
    cudaFree(dflag);
    return rc;
}

int tx_sweep_launch_count(int n, int M, int hop_mode) {
    (void)M;
    const int N = round_up_block(n);
    if (N == 0) return 0;
    // 5 <= n <= 8 with scalar hopping: structure scan + the two instantiations of the telescoped sweep (one of
    // them returns at once); an affine family adds its expansion kernel
    if (N == 8 && hop_mode == TX_HOP_SCALAR) return g_default_opts.diag_tail ? 3 : 1;
    return 1;
}

}  // extern "C"

namespace {

// Chunk c of the Hamiltonian axis [p0, p0 + np_) of a prepared sweep: every per-Hamiltonian pointer moves, the batch-wide
// site statistics stay.
SweepParams slice_params(const SweepParams& p, long long p0, long long np_) {
    SweepParams q = p;
    const long long o_pts = p0 * p.nE;
    q.h += p0 * p.h_stride;
    q.tnn += p0 * p.t_stride;
    if (q.sigL) q.sigL += p0 * p.sig_stride;
    if (q.sigR) q.sigR += p0 * p.sig_stride;
    if (q.params) q.params += p0;
    if (q.R) q.R += o_pts * p.n_src * p.n;
    if (q.T) q.T += o_pts * p.n_src * p.n;
    if (q.resid) q.resid += o_pts * p.n_src;
    if (q.ttot) q.ttot += o_pts;
    if (q.spin) q.spin += o_pts * p.n_src * p.spin_n;
    q.gather.off += o_pts;
    if (q.nondiag_max) q.nondiag_max += p0 * p.nondiag_stride;
    if (q.site_class) q.site_class += p0 * p.class_stride;
    q.n_pts = np_ * p.nE;
    return q;
}

// Called after the launches of every chunk (the host entry records an event and queues the chunk's device->host copies).
struct ChunkHook {
    virtual int after_chunk(int c, long long p0, long long np_, cudaStream_t st) = 0;
    virtual ~ChunkHook() {}
};

// Argument checks of the device-pointer sweep (run before any CUDA call, so that bad arguments are reported as
// such even without a device).
int check_small_args(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                    const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE, int lead_mode,
                    const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode, const double* sources, int n_src,
                    const SweepOut& o, const int* singular_flag) {
    if (!h || !tnn || !E || !singular_flag) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || M < 2 || n_ham < 1 || nE < 1 || n_src < 1)
        return fail(TX_ERR_ARG, "bad sizes n=%d M=%d n_ham=%d nE=%d n_src=%d", n, M, n_ham, nE, n_src);
    if (!(n_h == 1 || n_h == n_ham) || !(n_t == 1 || n_t == n_ham))
        return fail(TX_ERR_ARG, "n_h/n_t must be 1 or n_ham");
    if (h1 && !params) return fail(TX_ERR_ARG, "h1 given without params");
    if (!sources && n_src != n) return fail(TX_ERR_ARG, "sources == NULL requires n_src == n");
    if (int rc = check_spin(o, n)) return rc;
    if (o.caroli) return fail(TX_ERR_ARG, "caroli is produced by the large-block sweep only");
    if (lead_mode == TX_LEAD_TABLE) {
        if (!sigL || !sigR) return fail(TX_ERR_ARG, "TX_LEAD_TABLE needs sigL and sigR");
        if (!(n_sig == 1 || n_sig == n_ham)) return fail(TX_ERR_ARG, "n_sig must be 1 or n_ham");
    } else if (lead_mode != TX_LEAD_CLOSED) {
        return fail(TX_ERR_ARG, "unknown lead_mode %d", lead_mode);
    }
    if (!aligned16(h) || !aligned16(tnn) || !aligned16(E) || (h1 && !aligned16(h1)) ||
        (sigL && !aligned16(sigL)) || (sigR && !aligned16(sigR)))
        return fail(TX_ERR_ARG, "complex arrays must be 16-byte aligned");
    if (!aligned16(o.R) || !aligned16(o.T) || !aligned16(o.spin)) return fail(TX_ERR_ARG, "R, T and spin must be 16-byte aligned");
    const int N = round_up_block(n);
    if (N == 0) return fail(TX_ERR_UNSUPPORTED, "block size n=%d > 16 not covered by the small-block sweep", n);
    return TX_OK;
}

// Device-pointer sweep for n <= 16 on stream `st`; the caller holds cx.mu.
int sweep_small_dev(tx_context& cx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                    const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE, int lead_mode,
                    const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode, const double* sources, int n_src,
                    const SweepOut& o, int* singular_flag, cudaStream_t st, int n_chunks = 1, ChunkHook* hook = nullptr) {
    if (int rc = check_small_args(h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, lead_mode, sigL, sigR, n_sig, hop_mode,
                                  sources, n_src, o, singular_flag))
        return rc;
    const int N = round_up_block(n);
    if (cx.follows_defaults) {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        cx.opt = g_default_opts;
    }

    SweepParams p;
    p.h = reinterpret_cast<const cd*>(h);
    p.h_stride = (n_h == 1) ? 0 : (long long)M * n * n;
    p.h1 = reinterpret_cast<const cd*>(h1);
    p.params = params;
    p.tnn = reinterpret_cast<const cd*>(tnn);
    p.t_stride = (n_t == 1) ? 0 : (long long)(M - 1) * n * n;
    p.E = reinterpret_cast<const cd*>(E);
    p.sigL = reinterpret_cast<const cd*>(sigL);
    p.sigR = reinterpret_cast<const cd*>(sigR);
    p.sig_stride = (n_sig == 1) ? 0 : (long long)nE * n * n;
    p.sources = sources;
    p.R = o.R;
    p.T = o.T;
    p.resid = o.resid;
    p.ttot = o.t_total;
    p.spin = o.spin;
    p.n_up = o.n_up;
    p.spin_n = o.spin_n;
    p.gather = cx.gather;
    p.singular = singular_flag;
    p.nondiag_max = nullptr;
    p.nondiag_stride = 0;
    p.site_class = nullptr;
    p.class_stride = 0;
    p.class_count = nullptr;
    p.class_sites = 0;
    p.n_pts = (long long)n_ham * nE;
    p.n = n;
    p.M = M;
    p.nE = nE;
    p.n_src = n_src;

    const int hop = (hop_mode == TX_HOP_SCALAR) ? HOP_SCALAR : HOP_GENERAL;
    const int lead = (lead_mode == TX_LEAD_TABLE) ? LEAD_TABLE : LEAD_CLOSED;
    // the context's scratch (expanded Hamiltonians, site classes) may still be read by kernels of an earlier call
    // on another stream: order this call behind them
    if (cx.scratch_used && cx.scratch_stream != st && cx.scratch_ev) TX_CUDA(cudaStreamWaitEvent(st, cx.scratch_ev, 0));
    cx.scratch_used = false;
    // kernel family; the n = 8 scalar-hopping kernels share one preparation per sweep
    // default: telescoped sweep; its warps take eight energies of one Hamiltonian, so very short energy axes that are
    // not a multiple of eight go to the plain sweep (one point per four lanes) instead
    const bool fam_n8 = N == 8 && hop == HOP_SCALAR && cx.opt.variant != 0;
    const bool fam_n8t = fam_n8 && cx.opt.variant == 13 && (nE >= 64 || nE % 8 == 0);
    int rc = TX_OK;
    if (fam_n8) rc = prepare_n8(cx, p, st, !fam_n8t && lead == LEAD_CLOSED);
    if (rc) return rc;
    if (n_chunks < 1) n_chunks = 1;
    for (int c = 0; c < n_chunks; ++c) {
        const long long p0 = (long long)n_ham * c / n_chunks, p1 = (long long)n_ham * (c + 1) / n_chunks;
        if (p1 <= p0) continue;
        const SweepParams q = n_chunks == 1 ? p : slice_params(p, p0, p1 - p0);
        rc = TX_ERR_UNSUPPORTED;
        switch (N) {
            case 1: rc = dispatch_modes<1, 1, 4, 4>(q, hop, lead, st); break;
            case 2: rc = dispatch_modes<2, 1, 4, 4>(q, hop, lead, st); break;
            case 4: rc = dispatch_modes<4, 2, 3, 3>(q, hop, lead, st); break;
            case 8:
                if (fam_n8t)
                    rc = (lead == LEAD_TABLE) ? launch_n8t<LEAD_TABLE, 2>(cx, q, st) : launch_n8t<LEAD_CLOSED, 2>(cx, q, st);
                else if (fam_n8)
                    rc = (lead == LEAD_TABLE) ? launch_n8<LEAD_TABLE, 3, 0>(cx, q, st) : launch_n8<LEAD_CLOSED, 3, 0>(cx, q, st);
                else
                    rc = dispatch_modes<8, 4, 3, 2>(q, hop, lead, st);
                break;
            case 16: rc = dispatch_modes<16, 16, 3, 2>(q, hop, lead, st); break;
        }
        if (rc) return rc;
        if (hook && (rc = hook->after_chunk(c, p0, p1 - p0, st))) return rc;
    }
    if (rc == TX_OK && cx.scratch_used) {
        if (!cx.scratch_ev) TX_CUDA(cudaEventCreateWithFlags(&cx.scratch_ev, cudaEventDisableTiming));
        TX_CUDA(cudaEventRecord(cx.scratch_ev, st));
        cx.scratch_stream = st;
    }
    return rc;
}

int sweep_large_dev(tx_context& cx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params, const tx_cplx* tnn,
                    int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE, const tx_cplx* sigL, const tx_cplx* sigR,
                    int n_sig, int hop_mode, const double* sources, int n_src, const SweepOut& o, tx_cplx* workspace,
                    int* singular_flag, cudaStream_t st) {
    if (!h || !tnn || !E || !sigL || !sigR || !singular_flag) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || n > TX_MAX_N) return fail(TX_ERR_UNSUPPORTED, "block size n=%d outside 1..%d", n, TX_MAX_N);
    if (M < 2 || n_ham < 1 || nE < 1 || n_src < 1) return fail(TX_ERR_ARG, "bad sizes");
    if (!(n_h == 1 || n_h == n_ham) || !(n_t == 1 || n_t == n_ham) || !(n_sig == 1 || n_sig == n_ham))
        return fail(TX_ERR_ARG, "n_h/n_t/n_sig must be 1 or n_ham");
    if (!sources && n_src != n) return fail(TX_ERR_ARG, "sources == NULL requires n_src == n");
    if (h1 && !params) return fail(TX_ERR_ARG, "h1 given without params");
    if (int rc = check_spin(o, n)) return rc;
    if (cx.follows_defaults) {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        cx.opt = g_default_opts;
    }
    GenericParams p{};
    p.h = reinterpret_cast<const cd*>(h);
    p.h_stride = (n_h == 1) ? 0 : (long long)M * n * n;
    p.h1 = reinterpret_cast<const cd*>(h1);
    p.params = params;
    p.tnn = reinterpret_cast<const cd*>(tnn);
    p.t_stride = (n_t == 1) ? 0 : (long long)(M - 1) * n * n;
    p.E = reinterpret_cast<const cd*>(E);
    p.sigL = reinterpret_cast<const cd*>(sigL);
    p.sigR = reinterpret_cast<const cd*>(sigR);
    p.sig_stride = (n_sig == 1) ? 0 : (long long)nE * n * n;
    p.sources = sources;
    p.R = o.R;
    p.T = o.T;
    p.resid = o.resid;
    p.ttot = o.t_total;
    p.caroli = o.caroli;
    p.spin = o.spin;
    p.n_up = o.n_up;
    p.spin_n = o.spin_n;
    p.gather = cx.gather;
    p.work = reinterpret_cast<cd*>(workspace);
    p.singular = singular_flag;
    p.n_pts = (long long)n_ham * nE;
    p.n = n;
    p.M = M;
    p.nE = nE;
    p.n_src = n_src;
    p.hop_scalar = (hop_mode == TX_HOP_SCALAR) ? 1 : 0;
    p.kmax = cx.opt.chunk_kmax;
    p.gexp = cx.opt.chunk_gexp;
    // The structure scan's buffer and the caller's workspace of a context are reused from call to call: same ordering
    // protocol across streams as in sweep_small_dev (an event behind the last user, waited for on the device)
    if (cx.scratch_used && cx.scratch_stream != st && cx.scratch_ev) TX_CUDA(cudaStreamWaitEvent(st, cx.scratch_ev, 0));
    cx.scratch_used = false;
    if (p.hop_scalar && p.kmax > 1 && cx.opt.diag_tail && M > 2) {
        // structure scan (one warp per site): diagonal on-site blocks become row scalings in the telescoped sweep
        SweepParams sp{};
        sp.h = p.h;
        sp.h_stride = p.h_stride;
        sp.h1 = p.h1;
        sp.n_pts = p.n_pts;
        sp.nE = nE;
        sp.M = M;
        sp.n = n;
        if (int rc = scan_structure(cx, sp, st, false)) return rc;
        p.site_class = sp.site_class;
        p.class_stride = sp.class_stride;
    }
    const int rc = launch_rgf_generic(p, st);
    if (rc == TX_OK) {
        if (!cx.scratch_ev) TX_CUDA(cudaEventCreateWithFlags(&cx.scratch_ev, cudaEventDisableTiming));
        TX_CUDA(cudaEventRecord(cx.scratch_ev, st));
        cx.scratch_stream = st;
        cx.scratch_used = true;
    }
    return rc;
}

int resolve_context(tx_context* ctx, tx_context** out) {
    if (ctx) {
        *out = ctx;
        return TX_OK;
    }
    return default_context(out);
}

}  // namespace

extern "C" {

int tx_transmission_sweep_spin_dev(tx_context* ctx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                                   const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE,
                                   int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                                   const double* sources, const int* src_idx, int n_src, double* R, double* T,
                                   double* resid, double* t_total, int n_up, int spin_n, double* spin,
                                   int* singular_flag, void* stream) {
    if (src_idx) return fail(TX_ERR_ARG, "src_idx is resolved by the host entry point; pass sources to the device entry");
    const SweepOut o{R, T, resid, t_total, nullptr, spin, n_up, spin_n};
    if (int rc = check_small_args(h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, lead_mode, sigL, sigR, n_sig, hop_mode,
                                  sources, n_src, o, singular_flag))
        return rc;
    tx_context* cx = nullptr;
    if (int rc = resolve_context(ctx, &cx)) return rc;
    std::lock_guard<std::mutex> lock(cx->mu);
    return sweep_small_dev(*cx, h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, lead_mode, sigL, sigR, n_sig, hop_mode,
                           sources, n_src, o, singular_flag, reinterpret_cast<cudaStream_t>(stream));
}

int tx_transmission_sweep_dev(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                              const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                              const tx_cplx* E, int nE,
                              int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig,
                              int hop_mode,
                              const double* sources, const int* src_idx, int n_src,
                              double* R, double* T, double* resid, double* t_total,
                              int* singular_flag, void* stream) {
    return tx_transmission_sweep_spin_dev(nullptr, h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, lead_mode, sigL, sigR,
                                          n_sig, hop_mode, sources, src_idx, n_src, R, T, resid, t_total, 0, 0, nullptr,
                                          singular_flag, stream);
}

long long tx_sweep_workspace_elems(int n, long long n_pts) { return (long long)rgf_generic_workspace_elems(n, n_pts); }

int tx_transmission_sweep_large_dev(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                                    const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                                    const tx_cplx* E, int nE,
                                    const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                                    const double* sources, int n_src,
                                    double* R, double* T, double* resid, double* t_total, double* caroli,
                                    tx_cplx* workspace, int* singular_flag, void* stream) {
    tx_context* cx = nullptr;
    if (int rc = resolve_context(nullptr, &cx)) return rc;
    std::lock_guard<std::mutex> lock(cx->mu);
    const SweepOut o{R, T, resid, t_total, caroli, nullptr, 0, 0};
    return sweep_large_dev(*cx, h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, sigL, sigR, n_sig, hop_mode, sources, n_src,
                           o, workspace, singular_flag, reinterpret_cast<cudaStream_t>(stream));
}

int tx_transmission_sweep_spin(tx_context* ctx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                               const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE,
                               int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                               const double* sources, const int* src_idx, int n_src, double* R, double* T,
                               double* resid, double* t_total, double* caroli, int n_up, int spin_n, double* spin) {
    if (!h || !tnn || !E) return fail(TX_ERR_ARG, "null pointer argument");
    if (n < 1 || M < 2 || n_ham < 1 || nE < 1 || n_src < 1)
        return fail(TX_ERR_ARG, "bad sizes n=%d M=%d n_ham=%d nE=%d n_src=%d", n, M, n_ham, nE, n_src);
    if (!(n_h == 1 || n_h == n_ham) || !(n_t == 1 || n_t == n_ham))
        return fail(TX_ERR_ARG, "n_h/n_t must be 1 or n_ham");
    if (sources && src_idx) return fail(TX_ERR_ARG, "give sources or src_idx, not both");
    if (lead_mode < TX_LEAD_CLOSED || lead_mode > TX_LEAD_SEPARABLE) return fail(TX_ERR_ARG, "unknown lead_mode %d", lead_mode);
    {
        const SweepOut chk{R, T, resid, t_total, caroli, spin, n_up, spin_n};
        if (int rc = check_spin(chk, n)) return rc;
    }
    if (tx_device_count() < 1) return fail(TX_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");

    const size_t nn = (size_t)n * n;
    // hopping structure (TX_HOP_AUTO): scalar multiples of the identity -> cheaper kernel
    if (hop_mode == TX_HOP_AUTO) {
        hop_mode = TX_HOP_SCALAR;
        const size_t blocks = (size_t)n_t * (M - 1);
        for (size_t b = 0; b < blocks && hop_mode == TX_HOP_SCALAR; ++b) {
            const tx_cplx* t = tnn + b * nn;
            for (int r = 0; r < n && hop_mode == TX_HOP_SCALAR; ++r)
                for (int c = 0; c < n; ++c) {
                    const tx_cplx v = t[(size_t)r * n + c];
                    if (r == c) {
                        if (v.re != t[0].re || v.im != t[0].im) hop_mode = TX_HOP_GENERAL;
                    } else if (v.re != 0.0 || v.im != 0.0) {
                        hop_mode = TX_HOP_GENERAL;
                    }
                }
        }
    }
    // one-hot index lists become amplitude vectors
    std::vector<double> src_host;
    if (src_idx) {
        src_host.assign((size_t)n_src * n, 0.0);
        for (int i = 0; i < n_src; ++i) {
            if (src_idx[i] < 0 || src_idx[i] >= n) return fail(TX_ERR_ARG, "src_idx[%d]=%d out of range", i, src_idx[i]);
            src_host[(size_t)i * n + src_idx[i]] = 1.0;
        }
        sources = src_host.data();
    }
    if (!sources && n_src != n) return fail(TX_ERR_ARG, "sources == NULL requires n_src == n");

    tx_context* cxp = nullptr;
    if (int rc = resolve_context(ctx, &cxp)) return rc;
    tx_context& ws = *cxp;
    std::lock_guard<std::mutex> lock(ws.mu);
    if (ctx) {
        int cur = 0;
        TX_CUDA(cudaGetDevice(&cur));
        if (cur != ws.device) return fail(TX_ERR_ARG, "context belongs to device %d, current device is %d", ws.device, cur);
    }
    cudaStream_t st;
    if (ws.has_user_stream) {
        st = ws.user_stream;
    } else {
        if (!ws.own_stream) TX_CUDA(cudaStreamCreateWithFlags(&ws.own_stream, cudaStreamNonBlocking));
        st = ws.own_stream;
    }

    const size_t n_pts = (size_t)n_ham * nE;
    const size_t bytes_h = (size_t)n_h * M * nn * sizeof(tx_cplx);
    const size_t bytes_h1 = h1 ? (size_t)M * nn * sizeof(tx_cplx) : 0;
    const size_t bytes_t = (size_t)n_t * (M - 1) * nn * sizeof(tx_cplx);
    const size_t bytes_E = (size_t)nE * sizeof(tx_cplx);
    const size_t bytes_sig = (lead_mode == TX_LEAD_TABLE) ? (size_t)n_sig * nE * nn * sizeof(tx_cplx) : 0;
    const size_t bytes_src = sources ? (size_t)n_src * n * sizeof(double) : 0;
    const size_t bytes_out = R ? n_pts * n_src * n * sizeof(double) : 0;
    const size_t bytes_res = resid ? n_pts * n_src * sizeof(double) : 0;
    const size_t bytes_spin = spin ? n_pts * n_src * spin_n * sizeof(double) : 0;

    int rc;
    if ((rc = ws.h.ensure(bytes_h)) || (rc = ws.tnn.ensure(bytes_t)) || (rc = ws.E.ensure(bytes_E)) ||
        (rc = ws.flag.ensure(sizeof(int))))
        return rc;
    if (bytes_out && ((rc = ws.R.ensure(bytes_out)) || (rc = ws.T.ensure(bytes_out)))) return rc;
    if (h1 && ((rc = ws.h1.ensure(bytes_h1)) || (rc = ws.par.ensure((size_t)n_ham * sizeof(double))))) return rc;
    if (bytes_sig && ((rc = ws.sigL.ensure(bytes_sig)) || (rc = ws.sigR.ensure(bytes_sig)))) return rc;
    if (bytes_src && (rc = ws.src.ensure(bytes_src))) return rc;
    if (bytes_res && (rc = ws.resid.ensure(bytes_res))) return rc;
    if (bytes_spin && (rc = ws.spin.ensure(bytes_spin))) return rc;
    if (t_total && (rc = ws.ttot.ensure(n_pts * sizeof(double)))) return rc;

    TX_CUDA(cudaMemcpyAsync(ws.h.ptr, h, bytes_h, cudaMemcpyHostToDevice, st));
    TX_CUDA(cudaMemcpyAsync(ws.tnn.ptr, tnn, bytes_t, cudaMemcpyHostToDevice, st));
    TX_CUDA(cudaMemcpyAsync(ws.E.ptr, E, bytes_E, cudaMemcpyHostToDevice, st));
    if (h1) {
        if (!params) return fail(TX_ERR_ARG, "h1 given without params");
        TX_CUDA(cudaMemcpyAsync(ws.h1.ptr, h1, bytes_h1, cudaMemcpyHostToDevice, st));
        TX_CUDA(cudaMemcpyAsync(ws.par.ptr, params, (size_t)n_ham * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (bytes_sig) {
        if (!sigL || !sigR) return fail(TX_ERR_ARG, "TX_LEAD_TABLE needs sigL and sigR");
        TX_CUDA(cudaMemcpyAsync(ws.sigL.ptr, sigL, bytes_sig, cudaMemcpyHostToDevice, st));
        TX_CUDA(cudaMemcpyAsync(ws.sigR.ptr, sigR, bytes_sig, cudaMemcpyHostToDevice, st));
    }
    if (bytes_src) TX_CUDA(cudaMemcpyAsync(ws.src.ptr, sources, bytes_src, cudaMemcpyHostToDevice, st));
    TX_CUDA(cudaMemsetAsync(ws.flag.ptr, 0, sizeof(int), st));

    const bool large = (n > 16) || (caroli != nullptr);
    bool copied_out = false;
    // Lead stage (K1) for the modes that need device tables: iterative / separable multi-channel leads for any n, and
    // closed-form leads ahead of the large-block kernel.  The tables never leave the device.
    int n_sig_use = n_sig;
    const bool build_tables = lead_mode == TX_LEAD_SANCHO || lead_mode == TX_LEAD_SEPARABLE || (large && lead_mode == TX_LEAD_CLOSED);
    bool check_conv = false;
    if (build_tables) {
        // the lead blocks must be common to all Hamiltonians of the batch
        const size_t lead_off_R = (size_t)(M - 1) * nn, hop_off_R = (size_t)(M - 2) * nn;
        for (int q = 1; q < n_h; ++q)
            if (memcmp(h + (size_t)q * M * nn, h, nn * sizeof(tx_cplx)) ||
                memcmp(h + (size_t)q * M * nn + lead_off_R, h + lead_off_R, nn * sizeof(tx_cplx)))
                return fail(TX_ERR_UNSUPPORTED, "device-built lead tables need identical lead blocks in every Hamiltonian; pass tables");
        for (int q = 1; q < n_t; ++q)
            if (memcmp(tnn + (size_t)q * (M - 1) * nn, tnn, nn * sizeof(tx_cplx)) ||
                memcmp(tnn + (size_t)q * (M - 1) * nn + hop_off_R, tnn + hop_off_R, nn * sizeof(tx_cplx)))
                return fail(TX_ERR_UNSUPPORTED, "device-built lead tables need identical lead hoppings in every Hamiltonian; pass tables");
        if (h1)
            for (size_t i = 0; i < nn; ++i)
                if (h1[i].re != 0 || h1[i].im != 0 || h1[lead_off_R + i].re != 0 || h1[lead_off_R + i].im != 0)
                    return fail(TX_ERR_UNSUPPORTED, "affine term on the lead blocks is not supported with device-built lead tables");
        const size_t bs = (size_t)nE * nn * sizeof(tx_cplx);
        if ((rc = ws.sigL.ensure(bs)) || (rc = ws.sigR.ensure(bs))) return rc;
        const int gf_mode = lead_mode == TX_LEAD_SANCHO ? TX_GF_SANCHO : (lead_mode == TX_LEAD_SEPARABLE ? TX_GF_SEPARABLE : TX_GF_CLOSED);
        check_conv = gf_mode == TX_GF_SANCHO;
        int *nitL = nullptr, *nitR = nullptr, *okL = nullptr, *okR = nullptr;
        if (check_conv) {
            if ((rc = ws.nit.ensure(2 * (size_t)nE * sizeof(int))) || (rc = ws.okb.ensure(2 * (size_t)nE * sizeof(int)))) return rc;
            nitL = (int*)ws.nit.ptr;  nitR = nitL + nE;
            okL = (int*)ws.okb.ptr;   okR = okL + nE;
        }
        const double imE = gf_mode == TX_GF_CLOSED ? 0.0 : ws.lead_imE;
        const tx_cplx* dh = (const tx_cplx*)ws.h.ptr;
        const tx_cplx* dt = (const tx_cplx*)ws.tnn.ptr;
        // both leads the same material (the usual junction): one decimation yields both surfaces
        const bool same_leads = gf_mode == TX_GF_SANCHO && !memcmp(h, h + lead_off_R, nn * sizeof(tx_cplx)) &&
                                !memcmp(tnn, tnn + hop_off_R, nn * sizeof(tx_cplx));
        if (same_leads) {
            rc = tx_surface_gf_pair_dev(dh, dt, n, (const tx_cplx*)ws.E.ptr, nE, imE, ws.lead_tol, ws.lead_max_iter, nullptr,
                                        (tx_cplx*)ws.sigL.ptr, nullptr, (tx_cplx*)ws.sigR.ptr, nitL, nullptr, okL, (int*)ws.flag.ptr, st);
            if (rc) return rc;
            TX_CUDA(cudaMemcpyAsync(nitR, nitL, (size_t)nE * sizeof(int), cudaMemcpyDeviceToDevice, st));
            TX_CUDA(cudaMemcpyAsync(okR, okL, (size_t)nE * sizeof(int), cudaMemcpyDeviceToDevice, st));
        } else {
            rc = tx_surface_gf_dev(dh, dt, n, (const tx_cplx*)ws.E.ptr, nE, -1, gf_mode, imE, ws.lead_tol, 0, 0, ws.lead_max_iter, nullptr,
                                   (tx_cplx*)ws.sigL.ptr, nitL, nullptr, okL, (int*)ws.flag.ptr, st);
            if (rc) return rc;
            rc = tx_surface_gf_dev(dh + lead_off_R, dt + hop_off_R, n, (const tx_cplx*)ws.E.ptr, nE, 1, gf_mode, imE, ws.lead_tol, 0, 0,
                                   ws.lead_max_iter, nullptr, (tx_cplx*)ws.sigR.ptr, nitR, nullptr, okR, (int*)ws.flag.ptr, st);
            if (rc) return rc;
        }
        n_sig_use = 1;
        lead_mode = TX_LEAD_TABLE;
    }
    const bool have_tables = lead_mode == TX_LEAD_TABLE;
    if (!large) {
        // Chunk the Hamiltonian axis so that the device->host copy of chunk c overlaps the sweep of
        // chunk c+1 (two streams, one event per chunk).  The results dominate the PCIe traffic.
        const size_t per_ham_out = (size_t)nE * n_src * (R ? 2 * n : (spin ? spin_n : 0)) * sizeof(double);
        int n_chunks = 1;
        if (n_ham >= 8 && (size_t)n_ham * per_ham_out > ((size_t)32 << 20)) n_chunks = n_ham >= 64 ? 16 : 4;
        if (!ws.copy_stream) TX_CUDA(cudaStreamCreateWithFlags(&ws.copy_stream, cudaStreamNonBlocking));
        if (!ws.ev[0])
            for (int i = 0; i < 2; ++i) TX_CUDA(cudaEventCreateWithFlags(&ws.ev[i], cudaEventDisableTiming));
        // one preparation (structure scan, affine expansion) for the whole batch; per chunk only the sweep kernels
        struct CopyOut : ChunkHook {
            tx_context& ws; double *R, *T, *spin; size_t row, srow; int nE, n_chunks; bool* copied;
            CopyOut(tx_context& w, double* r, double* t, double* s, size_t row_, size_t srow_, int nE_, int nc, bool* cp)
                : ws(w), R(r), T(t), spin(s), row(row_), srow(srow_), nE(nE_), n_chunks(nc), copied(cp) {}
            int after_chunk(int c, long long p0, long long np_, cudaStream_t st) override {
                if (n_chunks == 1) return TX_OK;          // single chunk: the copies follow on the same stream
                cudaEvent_t ev = ws.ev[c & 1];
                TX_CUDA(cudaEventRecord(ev, st));
                TX_CUDA(cudaStreamWaitEvent(ws.copy_stream, ev, 0));
                const size_t o_pts = (size_t)p0 * nE, pts = (size_t)np_ * nE;
                if (R) {
                    TX_CUDA(cudaMemcpyAsync(R + o_pts * row, (double*)ws.R.ptr + o_pts * row, pts * row * sizeof(double), cudaMemcpyDeviceToHost, ws.copy_stream));
                    TX_CUDA(cudaMemcpyAsync(T + o_pts * row, (double*)ws.T.ptr + o_pts * row, pts * row * sizeof(double), cudaMemcpyDeviceToHost, ws.copy_stream));
                }
                if (spin)
                    TX_CUDA(cudaMemcpyAsync(spin + o_pts * srow, (double*)ws.spin.ptr + o_pts * srow, pts * srow * sizeof(double),
                                            cudaMemcpyDeviceToHost, ws.copy_stream));
                *copied = true;
                return TX_OK;
            }
        } hook(ws, R, T, spin, (size_t)n_src * n, (size_t)n_src * spin_n, nE, n_chunks, &copied_out);
        const SweepOut o{R ? (double*)ws.R.ptr : nullptr, R ? (double*)ws.T.ptr : nullptr, bytes_res ? (double*)ws.resid.ptr : nullptr,
                         t_total ? (double*)ws.ttot.ptr : nullptr, nullptr, spin ? (double*)ws.spin.ptr : nullptr, n_up, spin_n};
        rc = sweep_small_dev(ws, (const tx_cplx*)ws.h.ptr, n_h, h1 ? (const tx_cplx*)ws.h1.ptr : nullptr,
                             h1 ? (const double*)ws.par.ptr : nullptr, (const tx_cplx*)ws.tnn.ptr, n_t, n, M, n_ham,
                             (const tx_cplx*)ws.E.ptr, nE, lead_mode, have_tables ? (const tx_cplx*)ws.sigL.ptr : nullptr,
                             have_tables ? (const tx_cplx*)ws.sigR.ptr : nullptr, n_sig_use, hop_mode,
                             bytes_src ? (const double*)ws.src.ptr : nullptr, n_src, o, (int*)ws.flag.ptr, st, n_chunks, &hook);
        if (rc) return rc;
        if (copied_out) TX_CUDA(cudaStreamSynchronize(ws.copy_stream));
    } else {
        // generic-size path (n <= 128): one CTA per point on the device tables
        const size_t wse = rgf_generic_workspace_elems(n, (long long)n_pts);
        if (wse && (rc = ws.big.ensure(wse * sizeof(tx_cplx)))) return rc;
        if (caroli && (rc = ws.caroli.ensure(n_pts * sizeof(double)))) return rc;
        const SweepOut o{R ? (double*)ws.R.ptr : nullptr, R ? (double*)ws.T.ptr : nullptr, bytes_res ? (double*)ws.resid.ptr : nullptr,
                         t_total ? (double*)ws.ttot.ptr : nullptr, caroli ? (double*)ws.caroli.ptr : nullptr,
                         spin ? (double*)ws.spin.ptr : nullptr, n_up, spin_n};
        rc = sweep_large_dev(ws, (const tx_cplx*)ws.h.ptr, n_h, h1 ? (const tx_cplx*)ws.h1.ptr : nullptr,
                             h1 ? (const double*)ws.par.ptr : nullptr, (const tx_cplx*)ws.tnn.ptr, n_t, n, M, n_ham,
                             (const tx_cplx*)ws.E.ptr, nE, (const tx_cplx*)ws.sigL.ptr, (const tx_cplx*)ws.sigR.ptr, n_sig_use,
                             hop_mode, bytes_src ? (const double*)ws.src.ptr : nullptr, n_src, o,
                             wse ? (tx_cplx*)ws.big.ptr : nullptr, (int*)ws.flag.ptr, st);
        if (rc) return rc;
        if (caroli) TX_CUDA(cudaMemcpyAsync(caroli, ws.caroli.ptr, n_pts * sizeof(double), cudaMemcpyDeviceToHost, st));
    }

    int flag = 0;
    if (!copied_out) {
        if (R) {
            TX_CUDA(cudaMemcpyAsync(R, ws.R.ptr, bytes_out, cudaMemcpyDeviceToHost, st));
            TX_CUDA(cudaMemcpyAsync(T, ws.T.ptr, bytes_out, cudaMemcpyDeviceToHost, st));
        }
        if (spin) TX_CUDA(cudaMemcpyAsync(spin, ws.spin.ptr, bytes_spin, cudaMemcpyDeviceToHost, st));
    }
    if (bytes_res) TX_CUDA(cudaMemcpyAsync(resid, ws.resid.ptr, bytes_res, cudaMemcpyDeviceToHost, st));
    if (t_total) TX_CUDA(cudaMemcpyAsync(t_total, ws.ttot.ptr, n_pts * sizeof(double), cudaMemcpyDeviceToHost, st));
    TX_CUDA(cudaMemcpyAsync(&flag, ws.flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    std::vector<int> conv_host;
    if (check_conv) {
        conv_host.resize(4 * (size_t)nE);
        TX_CUDA(cudaMemcpyAsync(conv_host.data(), ws.nit.ptr, 2 * (size_t)nE * sizeof(int), cudaMemcpyDeviceToHost, st));
        TX_CUDA(cudaMemcpyAsync(conv_host.data() + 2 * (size_t)nE, ws.okb.ptr, 2 * (size_t)nE * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    TX_CUDA(cudaStreamSynchronize(st));
    if (check_conv) {
        double iters = 0.0;
        for (size_t i = 0; i < 2 * (size_t)nE; ++i) iters += conv_host[i];
        ws.last_lead_iters = iters / (2.0 * nE);
        for (size_t i = 0; i < 2 * (size_t)nE; ++i)
            if (!conv_host[2 * (size_t)nE + i])
                return fail(TX_ERR_NOCONV, "Sancho-Rubio decimation did not converge at energy index %zu (%s lead) within %d iterations",
                            i % nE, i < (size_t)nE ? "left" : "right", ws.lead_max_iter);
    }
    if (flag) return fail(TX_ERR_SINGULAR, "a singular block was met during the sweep (results contain inf/nan there)");
    return TX_OK;
}

int tx_transmission_sweep(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                          const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                          const tx_cplx* E, int nE,
                          int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig,
                          int hop_mode,
                          const double* sources, const int* src_idx, int n_src,
                          double* R, double* T, double* resid, double* t_total, double* caroli) {
    return tx_transmission_sweep_spin(nullptr, h, n_h, h1, params, tnn, n_t, n, M, n_ham, E, nE, lead_mode, sigL, sigR, n_sig,
                                      hop_mode, sources, src_idx, n_src, R, T, resid, t_total, caroli, 0, 0, nullptr);
}

}  // extern "C"
