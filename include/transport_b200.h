/*
 * transport_b200 — C ABI of the B200-native T(E) scattering hot path.
 *
 * Drop-in boundary for the wave-function-matching path of cpbunker/transport.  The reference
 * has no FFI of its own (it is pure Python); every entry point below replaces
 * the arithmetic of one reference function and is what a ctypes/cffi binding inside
 * `transport/wfm/__init__.py` would call.  Citations are relative to the
 * reference root.
 *
 * Conventions (same as the reference):
 *   - complex numbers are interleaved (re, im) doubles  == numpy complex128;
 *   - h[M][n][n] on-site blocks, tnn[M-1][n][n] UPPER blocks H_{j,j+1}; the lower block is the
 *     conjugate transpose (wfm/__init__.py:203-207); blocks 0 and M-1 are the first cells of
 *     the left/right lead and receive the self-energies (wfm/__init__.py:246-247);
 *   - all arrays are C-contiguous, row-major;
 *   - functions return TX_OK (0) or a TX_ERR_* code; tx_last_error() gives the message of the
 *     last failure on the calling thread.
 *
 * No function here has a CPU fallback: without a CUDA device they return TX_ERR_CUDA.
 */
#ifndef TRANSPORT_B200_H
#define TRANSPORT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tx_cplx { double re, im; } tx_cplx;

enum {
    TX_OK = 0,
    TX_ERR_ARG = 1,         /* bad argument (shape, mode, null pointer)                       */
    TX_ERR_CUDA = 2,        /* CUDA runtime failure / no device                               */
    TX_ERR_SINGULAR = 3,    /* an exactly singular block was met (results contain inf/nan)    */
    TX_ERR_NOCONV = 4,      /* iterative surface GF did not converge within max_iter          */
    TX_ERR_UNSUPPORTED = 5  /* shape outside what the kernels cover                           */
};

/* How the lead self-energies are obtained. */
enum {
    TX_LEAD_CLOSED = 0,     /* wfm.g_closed fused into the sweep (wfm/__init__.py:335-340)    */
    TX_LEAD_TABLE = 1,      /* Sigma_L, Sigma_R supplied per energy (from tx_surface_gf_* or caller) */
    /* host-pointer sweep entries only: the self-energy tables are built on the device ahead of the sweep and
     * never leave it; parameters (imE, tolerance, iteration cap) from tx_context_set_lead_params              */
    TX_LEAD_SANCHO = 2,     /* Sancho-Rubio decimation of the multi-channel lead (TX_GF_SANCHO)               */
    TX_LEAD_SEPARABLE = 3   /* h01 = t*I leads by the separable closed form (TX_GF_SEPARABLE)                 */
};

/* Structure of the hopping blocks tnn. */
enum {
    TX_HOP_AUTO = 0,        /* host entry points inspect tnn; device entry points: treated as GENERAL */
    TX_HOP_SCALAR = 1,      /* every tnn[j] = t_j * I (all reference drivers: tnn = -tl*I)     */
    TX_HOP_GENERAL = 2      /* arbitrary complex blocks                                        */
};

/* Surface Green's function flavours (tx_surface_gf). */
enum {
    TX_GF_CLOSED = 0,       /* wfm.g_closed      wfm/__init__.py:306-340                       */
    TX_GF_RICEMELE = 1,     /* wfm.g_RiceMele    wfm/__init__.py:352-414                       */
    TX_GF_FIXEDPOINT = 2,   /* wfm.g_iter / g_ith, reference stopping rule  :479-551           */
    TX_GF_SANCHO = 3,       /* Sancho-Rubio decimation of the same fixed point                 */
    TX_GF_SEPARABLE = 4     /* h01 = t*I: g = U diag(g_closed(eps_m)) U^dag, h00 = U eps U^dag  */
};

/* ---- library / device management ------------------------------------------------------- */
const char* tx_version(void);
const char* tx_last_error(void);
int tx_device_count(void);                 /* number of CUDA devices, 0 if none              */
int tx_set_device(int device);
int tx_device_synchronize(void);

/* ---- contexts: reentrancy and stream selection (SURVEY.md §8b "threading / reentrancy") ---------------
 * The reference is synchronous and reentrant (plain Python functions, no global state).  Here everything a
 * sweep call needs besides its arguments lives in a tx_context: the device workspace of the host-pointer
 * entries, the scratch of the device-pointer entries (expanded affine families, site classes), the CUDA
 * stream the work is issued on and the tuning options.  Calls on ONE context are serialised (a host mutex;
 * reuse of its scratch by a later call on another stream is ordered by an event); calls on DISTINCT contexts
 * are independent, so host threads / streams that want concurrency create one context each.
 * Entry points without a context argument (and a NULL context) use a per-device default context, i.e. they
 * are thread-safe but serialised per device.
 *   tx_context_set_stream: use_stream != 0 -> the host-pointer entries issue their copies and kernels on
 *       `stream` (a cudaStream_t; NULL = the legacy default stream) and synchronise that stream before they
 *       return; use_stream == 0 (default) -> a context-owned non-blocking stream.
 *   tx_context_set_option: per-context copy of tx_set_option (options 0, 1, 2, 4).                     */
typedef struct tx_context tx_context;
int tx_context_create(int device, tx_context** out);
int tx_context_destroy(tx_context* ctx);
int tx_context_set_stream(tx_context* ctx, void* stream, int use_stream);
int tx_context_set_option(tx_context* ctx, int option, int value);
/* Parameters of the device-built lead tables (TX_LEAD_SANCHO, TX_LEAD_SEPARABLE): z = E + i*imE, decimation stops
 * at max|alpha|,|beta| < conv_tol or after max_iter iterations (then the sweep returns TX_ERR_NOCONV).  Defaults
 * 0, 1e-14, 200.  ctx = NULL addresses the per-device default context.  tx_context_last_lead_iterations returns the
 * mean number of decimation iterations per (energy, lead) of the last host-pointer sweep of the context. */
int tx_context_set_lead_params(tx_context* ctx, double imE, double conv_tol, int max_iter);
int tx_context_last_lead_iterations(tx_context* ctx, double* mean_iterations);

/* ---- multi-GPU gather fused into the sweep ("NCCL is used only to gather the T(E) arrays") ----
 * One process per GPU.  Each rank allocates its copy of the gathered array with tx_peer_alloc, exports it
 * (tx_peer_export: 64 opaque bytes, CUDA IPC) to the other ranks of the node over the host-side process group, and
 * maps theirs with tx_peer_open.  tx_context_set_gather then makes every sweep issued through `ctx` store the
 * per-point total transmission (the t_total output) ALSO to peers[q][offset + point] for q < n_peers (<= 8; the
 * rank's own buffer is one of them), straight from the K3 epilogue over NVLink, so no collective follows the sweep;
 * n_peers = 0 switches it off.  tx_peer_barrier_dev enqueues a device-side barrier on `stream`: rank `rank` raises
 * flag_peers[q][rank] = epoch on every rank q and waits for flag_peers[rank][q] >= epoch for all q (flag arrays: 8
 * ints each, tx_peer_alloc'ed and exchanged the same way; epoch must grow by 1 per call); after it, every peer's
 * stores of the preceding sweep are visible to later work on `stream`.  *timed_out (device int) is set to 1 if a
 * peer did not arrive within ~2 s (the kernel then returns instead of hanging the GPU). */
int tx_peer_alloc(long long bytes, void** ptr);
int tx_peer_free(void* ptr);
int tx_peer_export(const void* ptr, unsigned char* handle64);
int tx_peer_open(const unsigned char* handle64, void** ptr);
int tx_peer_close(void* ptr);
int tx_context_set_gather(tx_context* ctx, double* const* peers, int n_peers, long long offset);
int tx_peer_barrier_dev(int* const* flag_peers, int n_peers, int rank, int epoch, int* timed_out, void* stream);

/* ---- the hot path: batched T(E) sweep -------------------------------------------------- */
/*
 * Replaces, for every (Hamiltonian p, energy e) point and every source, one call of
 *   wfm.kernel(h, tnn, tnnn, tl, E, converger, Ajsigma, False, False)   wfm/__init__.py:29-137
 * i.e. SelfEnergies (:270-304) + Hmat/Hprime (:168-268) + Green's dense inverse (:139-166) +
 * the R/T channel loop (:108-130), evaluated as one right-to-left block-tridiagonal sweep.
 *
 *  h        [n_h][M][n][n]   n_h = 1 (shared) or n_ham
 *  h1       [M][n][n] or NULL; if given, Hamiltonian p uses h + params[p]*h1  (affine sweeps,
 *           e.g. the exchange coupling J of run_wfm_gate.py:get_hblocks)
 *  params   [n_ham] doubles (only read when h1 != NULL)
 *  tnn      [n_t][M-1][n][n] n_t = 1 (shared) or n_ham
 *  E        [nE] complex energies (the reference passes Im E = 0)
 *  lead_mode TX_LEAD_CLOSED: lead blocks h[0], h[M-1] must be diagonal (checked by the caller,
 *           wfm/__init__.py:327);  TX_LEAD_TABLE: sigL/sigR [n_sig][nE][n][n], n_sig = 1 or n_ham
 *  sources  [n_src][n] real incident amplitudes (Ajsigma) or NULL
 *  src_idx  [n_src] channel indices of one-hot sources or NULL; if both are NULL all n one-hot
 *           sources are evaluated (n_src must then equal n)
 *  R, T     [n_ham][nE][n_src][n] doubles: per-channel reflection / transmission probabilities (both may be NULL when
 *           another output is requested: nothing of that size is then allocated, written or copied)
 *  resid    [n_ham][nE][n_src] doubles or NULL: sum(R)+sum(T)-1 (the reference's unitarity
 *           assert, wfm/__init__.py:135, evaluated on the device)
 *  t_total  [n_ham][nE] doubles or NULL: transmission summed over all evaluated sources and
 *           channels; with all n one-hot sources this is Tr[Gamma_R G Gamma_L G^dag] — the compact
 *           array the multi-GPU path gathers
 *  caroli   (host entry only) [n_ham][nE] doubles or NULL: Tr[Gamma_R G Gamma_L G^dag] with the full
 *           matrices Gamma = i(Sigma - Sigma^dag) — the total transmission when the self-energies
 *           are not diagonal (multi-channel ribbon leads)
 *
 * Block sizes n <= 16 run the register-resident kernel; 16 < n <= 128 (or any n when `caroli` is
 * requested) run the CTA-per-point kernel tx_transmission_sweep_large_dev, which takes table
 * self-energies only (the host entry builds the tables with tx_surface_gf_dev for closed-form leads).
 * n <= 64: 256 threads, two CTAs per SM; 64 < n <= 128 (run_wfm_gate.get_hblocks gives n = 72 at TwoS = 5,
 * 128 at TwoS = 7): 512 threads, one CTA per SM.  n > 128 returns TX_ERR_UNSUPPORTED.
 *
 * tx_transmission_sweep takes HOST pointers and performs the host<->device copies itself;
 * tx_transmission_sweep_dev takes DEVICE pointers and a cudaStream_t (as void*), launches
 * asynchronously and performs no copies.  After one warm-up call with the same sizes (which grows the
 * context's scratch and sets the kernel attributes) the device entries allocate nothing and do not synchronise:
 * they may be captured into a CUDA graph and replayed (tests/test_gpu_compact.py).
 */
int tx_transmission_sweep(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                          const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                          const tx_cplx* E, int nE,
                          int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig,
                          int hop_mode,
                          const double* sources, const int* src_idx, int n_src,
                          double* R, double* T, double* resid, double* t_total, double* caroli);

int tx_transmission_sweep_dev(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                              const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                              const tx_cplx* E, int nE,
                              int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig,
                              int hop_mode,
                              const double* sources, const int* src_idx, int n_src,
                              double* R, double* T, double* resid, double* t_total,
                              int* singular_flag, void* stream);

int tx_transmission_sweep_large_dev(const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                                    const tx_cplx* tnn, int n_t, int n, int M, int n_ham,
                                    const tx_cplx* E, int nE,
                                    const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                                    const double* sources, int n_src,
                                    double* R, double* T, double* resid, double* t_total, double* caroli,
                                    tx_cplx* workspace, int* singular_flag, void* stream);
/*
 * The same sweeps with the spin-resolved channel sums fused into the K3 epilogue — the reduction the reference's
 * drivers do on the host after wfm.kernel (run_wfm_gate.py:404-405,473-477: the n_loc_dof channels split at
 * n_mol_dof into electron spin up / down) — and with every output array optional, so that a sweep that only
 * needs compact results moves 16-128 bytes per point instead of 16 n n_src:
 *   ctx      context or NULL (per-device default context)
 *   n_up     channels [0, n_up) carry electron spin up, [n_up, n) spin down
 *   spin_n   2 or 4
 *   spin     [n_ham][nE][n_src][spin_n] doubles or NULL:
 *              [0] sum of T over the channels of the source's own spin sector     (no spin flip)
 *              [1] sum of T over the other sector                                 (spin flip)
 *              [2], [3] the same for R                                            (spin_n == 4 only)
 *            A source's sector is the one holding the larger share of its incident weight sum_c A_c^2 (a one-hot
 *            source: its own channel's sector).
 *   R, T     may both be NULL (no per-channel arrays are allocated, written or copied); at least one output array
 *            must be given.
 * tx_transmission_sweep / tx_transmission_sweep_dev are these with ctx = NULL and spin = NULL.
 */
int tx_transmission_sweep_spin(tx_context* ctx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                               const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE,
                               int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                               const double* sources, const int* src_idx, int n_src, double* R, double* T,
                               double* resid, double* t_total, double* caroli, int n_up, int spin_n, double* spin);
int tx_transmission_sweep_spin_dev(tx_context* ctx, const tx_cplx* h, int n_h, const tx_cplx* h1, const double* params,
                                   const tx_cplx* tnn, int n_t, int n, int M, int n_ham, const tx_cplx* E, int nE,
                                   int lead_mode, const tx_cplx* sigL, const tx_cplx* sigR, int n_sig, int hop_mode,
                                   const double* sources, const int* src_idx, int n_src, double* R, double* T,
                                   double* resid, double* t_total, int n_up, int spin_n, double* spin,
                                   int* singular_flag, void* stream);

/* Elements (tx_cplx) of `workspace` the large-block sweep needs for n_pts points (0 if it fits in
 * shared memory). */
long long tx_sweep_workspace_elems(int n, long long n_pts);

/* Number of kernel launches tx_transmission_sweep_dev issues for this shape (for accounting; an affine
 * family h0 + J*h1 adds one expansion kernel, short energy axes may fall back to the plain sweep). */
int tx_sweep_launch_count(int n, int M, int hop_mode);
/* Running count of sweep / surface-GF kernel launches issued by this library in the process. */
long long tx_launch_counter(void);

/* ---- lead surface Green's functions and self-energies (K1) -------------------------------- */
/*
 * Batched over energies, one CTA per energy, any n <= 128 (TX_GF_SEPARABLE: n <= 64).  `side` = -1 for the left lead (incoming,
 * wfm/__init__.py:283), +1 for the right lead (:284).  `mode` is one of TX_GF_*:
 *   CLOSED      wfm.g_closed   (:306-340)  h00, h01 diagonal in channel space
 *   RICEMELE    wfm.g_RiceMele (:352-414)  n = 2*n_spin, [A spins..., B spins...] ordering
 *   FIXEDPOINT  wfm.g_iter     (:479-530)  z = Re E + i|imE|, stops when the relative change of the
 *               surface DOS stays below conv_tol for conv_rep consecutive iterations past min_iter
 *   SANCHO      Sancho-Rubio decimation of the same fixed point at z = E + i*imE; stops when
 *               max|alpha|,|beta| < conv_tol (conv_rep/min_iter unused)
 *   SEPARABLE   multi-channel lead whose hopping block is a scalar multiple of the identity (square-lattice
 *               ribbons): g commutes with h00, so g = U diag(g_closed(z, eps_m, t)) U^dag with the reference's
 *               g_closed per transverse mode and h00 = U diag(eps) U^dag (Jacobi eigensolver on the device, once
 *               per call); exact at imE = 0, no iteration.  z = E + i*imE.
 * Outputs (each may be NULL): g [nE][n][n]; sigma [nE][n][n] = h01^dag g h01 (side -1) or
 * h01 g h01^dag (side +1) (wfm.SelfEnergies :301-302); n_iter/conv/converged [nE] (iterative modes).
 * Non-convergence is reported per energy in `converged`, not as an error code.
 */
int tx_surface_gf(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, int mode,
                  double imE, double conv_tol, int conv_rep, int min_iter, int max_iter,
                  tx_cplx* g, tx_cplx* sigma, int* n_iter, double* conv, int* converged);
int tx_surface_gf_dev(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, int mode,
                      double imE, double conv_tol, int conv_rep, int min_iter, int max_iter,
                      tx_cplx* g, tx_cplx* sigma, int* n_iter, double* conv, int* converged,
                      int* singular_flag, void* stream);

/*
 * Both semi-infinite continuations of ONE periodic lead from a single Sancho-Rubio decimation: the right lead (side +1,
 * g_R = (z - h00 - h01 g_R h01^dag)^-1, Sigma_R = h01 g_R h01^dag) and the left lead (side -1,
 * g_L = (z - h00 - h01^dag g_L h01)^-1, Sigma_L = h01^dag g_L h01), z = E + i imE.  The decimation carries the bulk block
 * h00 + sum (alpha G beta + beta G alpha) and one surface block h00 + sum alpha G beta; the other surface block is their
 * difference plus h00, so a junction whose two leads are the same material (wfm.SelfEnergies, wfm/__init__.py:283-287, calls
 * its converger once per lead) costs one decimation instead of two.  Every output may be NULL; n_iter / conv / converged
 * [nE] as in tx_surface_gf.  The host sweep entry uses it by itself when TX_LEAD_SANCHO meets identical lead blocks.
 */
int tx_surface_gf_pair(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, double imE, double conv_tol,
                       int max_iter, tx_cplx* gL, tx_cplx* sigL, tx_cplx* gR, tx_cplx* sigR, int* n_iter, double* conv,
                       int* converged);
int tx_surface_gf_pair_dev(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, double imE, double conv_tol,
                           int max_iter, tx_cplx* gL, tx_cplx* sigL, tx_cplx* gR, tx_cplx* sigR, int* n_iter, double* conv,
                           int* converged, int* singular_flag, void* stream);
/* One fixed-point step, wfm.g_ith (:532-551): ith = 0 -> inv(z - h00); else inv(z - h00 - t g_prev t^dag). */
int tx_g_ith(const tx_cplx* h00, const tx_cplx* h01, int n, const tx_cplx* E, int nE, int side, double imE,
             int ith, const tx_cplx* g_prev, tx_cplx* g_out);

/* ---- Green's function blocks and Hamiltonian assembly (generic n <= 128) ------------------ */
/* First block column G[:,0] -> [nE][M][n][n]  (psi_jsigma of wfm.kernel, :98-99). Host pointers. */
int tx_green_column0(const tx_cplx* h, const tx_cplx* tnn, int n, int M, const tx_cplx* E, int nE,
                     const tx_cplx* sigL, const tx_cplx* sigR, tx_cplx* G);
/* All blocks of inv(E - H') -> [nE][M][M][n][n]  (wfm.Green, :139-166). Host pointers. */
int tx_green_full(const tx_cplx* h, const tx_cplx* tnn, int n, int M, const tx_cplx* E, int nE,
                  const tx_cplx* sigL, const tx_cplx* sigR, tx_cplx* G);
/* Dense block-pentadiagonal H [(M n)][(M n)]  (wfm.Hmat, :168-215). Host pointers. */
int tx_hmat(const tx_cplx* h, const tx_cplx* tnn, const tx_cplx* tnnn, int n, int M, tx_cplx* H);

/* bardeen.Hsysmat (transport/bardeen/__init__.py:690-750): 4-D Hamiltonian [S][S][n][n],
 * S = 2*Ninf + NL + NR + NC, regions inf|L|C|R|inf with block potentials V* and hoppings t* ([n][n]
 * each), HC [NC][NC][n][n] pasted in the centre (NC odd).  Host pointers. */
int tx_hsysmat(const tx_cplx* tinf, const tx_cplx* tL, const tx_cplx* tR, const tx_cplx* VinfL, const tx_cplx* VinfR,
               const tx_cplx* VL, const tx_cplx* VR, int Ninf, int NL, int NR, const tx_cplx* HC, int NC, int n,
               tx_cplx* H);

/* Landauer current from a transmission function (reference: fcdmft/__init__.py
 * landauer:238-270 + np.trapz :426): jE[b][i] = T[i] (n_L - n_R)/(2 pi), I[b] = trapz(jE[b], E), with
 * Fermi functions at kBT (a step function for kBT == 0) and chemical potentials muL[b], muR[b].
 * T, E [nE]; jE [n_bias][nE] or NULL; I [n_bias].  Host pointers / device pointers + stream. */
int tx_landauer_current(const double* T, const double* E, int nE, const double* muL, const double* muR, int n_bias,
                        double kBT, double* jE, double* I);
int tx_landauer_current_dev(const double* T, const double* E, int nE, const double* muL, const double* muR, int n_bias,
                            double kBT, double* jE, double* I, void* stream);

/* ---- Bardeen transfer-Hamiltonian path (BASELINE config 5; transport/bardeen/__init__.py) ---------------
 * The reference diagonalises dense well Hamiltonians with np.linalg.eigh (:459, :483, get_mstates :755-761)
 * and keeps the states below an energy cutoff (:463, :486, :174-182).  For chains every per-channel well
 * Hamiltonian is real symmetric tridiagonal: d[n_mat][S] diagonal, e[n_mat][S-1] off-diagonal.
 *
 * tx_tridiag_eigh_below: eigenpairs with eigenvalue < cut_states[mat], ascending, at most max_states per
 * matrix: evals [n_mat][max_states], evecs [n_mat][max_states][S] (unit norm; may be NULL);
 * n_states[mat] = number of eigenvalues < cut_states[mat] (NOT clamped: > max_states means truncated),
 * n_below[mat] = number < cut_count[mat] (cut_count NULL: same as cut_states; n_below may be NULL).
 * max_states = 0 only counts.  Host pointers.  The _dev forms take device pointers and a stream;
 * scratch is unused (the eigenvector kernel keeps its work rows in shared memory) and may be NULL.
 * Limit: S <= 8533 (shared-memory staging). */
int tx_tridiag_eigh_below(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                          const double* cut_count, int max_states, double* evals, double* evecs, int* n_states,
                          int* n_below);
int tx_tridiag_count_below_dev(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                               const double* cut_count, int* n_states, int* n_below, void* stream);
int tx_tridiag_eigh_below_dev(const double* d, const double* e, int S, int n_mat, const double* cut_states,
                              const int* n_states, int max_states, double* evals, double* evecs, double* scratch,
                              void* stream);
/* tx_bardeen_wells: P independent geometries x n channels.  Left-well states (dL, eL [P][n][S], [P][n][S-1];
 * kept below cutL [P][n]) and right-well states (dR, eR; computed below cutR_states, counted below
 * cutR_count) and the window-averaged Oppenheimer matrix elements (kernel_well :127-151, kernel_well_prime
 * :502-519, matrix_element :835-857)
 *   Mavg[p][beta][m][alpha] = mean over k with |EL[p][alpha][m] - ER[p][beta][k]| < interval of
 *                             <psiR[beta][k] | Hdiff[:, :, beta, alpha] | psiL[alpha][m]>
 * with each element made >= 0 first when sign_fix (:141), the nearest k when interval is NaN, 0.0 when no
 * state qualifies.  Hdiff = Hsys - HL is block tridiagonal in the site index: hd0 [P][S][n][n],
 * hdU [P][S-1][n][n] = Hdiff[j][j+1], hdL [P][S-1][n][n] = Hdiff[j+1][j] (real).
 * Outputs: EL, ER [P][n][max_states]; nL, nR_states, nR_below [P][n]; Mavg [P][n][max_states][n];
 * psiL, psiR [P][n][max_states][S] or NULL.  max_states = 0: counts only (size the buffers, call again). */
int tx_bardeen_wells(const double* dL, const double* eL, const double* dR, const double* eR, const double* cutL,
                     const double* cutR_states, const double* cutR_count, const double* hd0, const double* hdU,
                     const double* hdL, int P, int n, int S, int max_states, double interval, int sign_fix, double* EL,
                     int* nL, double* ER, int* nR_states, int* nR_below, double* Mavg, double* psiL, double* psiR);
long long tx_bardeen_workspace_bytes(int P, int n, int S, int max_states);
int tx_bardeen_wells_dev(const double* dL, const double* eL, const double* dR, const double* eR,
                         const double* cutL, const double* cutR_states, const double* cutR_count,
                         const double* hd0, const double* hdU, const double* hdL, int P, int n, int S, int max_states,
                         double interval, int sign_fix, double* EL, int* nL, double* ER, int* nR_states,
                         int* nR_below, double* Mavg, void* workspace, void* stream);
/* The same pass driven by the REGION PARAMETERS of bardeen.Hsysmat (:690-750) instead of assembled arrays: the three
 * Hamiltonians of kernel_well (:90-125) / kernel_well_prime (:455-500) — left well (Vinf, VL, HCprime, VRprime, Vinf),
 * right well (Vinf, VLprime, HCprime, VR, Vinf), system (Vinf, VL, HC, VR, Vinf) — are assembled ON THE DEVICE, so a sweep
 * over step heights / barrier parameters ships a few doubles per geometry (rbstep.py:43-104 loops kernel_well over VR).
 *   tinf, tL, tR, Vinf [n]: per-channel scalar hoppings (positive t: the blocks hold -t, :718-744) and outer potential
 *   VL, VLprime, VR, VRprime [n_pot][n], n_pot = 1 (shared) or P
 *   hc0 [n_hc][NC][n][n], hcU / hcL [n_hc][NC-1][n][n]: the chain HC (diagonal blocks, HC[j][j+1], HC[j+1][j]; real);
 *   hp0, hpU, hpL: the same for HCprime (NULL = HC);  n_hc = 1 or P
 *   cutL, cutR_states, cutR_count [P][n] as in tx_bardeen_wells;  row_cut != 0 zeroes |M|^2 rows m >= nR_below (:181)
 *   max_states: capacity of EL, ER [P][n][max_states], Mavg [P][n][max_states][n].  *max_needed receives the largest
 *   count; when it exceeds max_states (or max_states = 0) only the counts nL, nR_states, nR_below are returned and the
 *   caller calls again with larger buffers.
 *   Vb [nVb], muR, kBT, I [P][n][n][nVb] (or NULL): bardeen.current (:537-575) of |Mavg|^2 on the device.
 * Host pointers; the device workspace is kept between calls (calls are serialised per process). */
int tx_bardeen_region_sweep(const double* tinf, const double* tL, const double* tR, const double* Vinf,
                            const double* VL, const double* VLprime, const double* VR, const double* VRprime, int n_pot,
                            const double* hc0, const double* hcU, const double* hcL,
                            const double* hp0, const double* hpU, const double* hpL, int n_hc,
                            int Ninf, int NL, int NR, int NC, int n, int P,
                            const double* cutL, const double* cutR_states, const double* cutR_count,
                            double interval, int sign_fix, int row_cut, int max_states, int* max_needed,
                            double* EL, int* nL, double* ER, int* nR_states, int* nR_below, double* Mavg,
                            const double* Vb, int nVb, double muR, double kBT, double* I);
/* Small real matrix product for the basis changes of kernel_well (:156-166): C[M][N] = opA(A) diag(w) opB(B), contraction
 * length K; transA: A is stored [K][M] (else [M][K]); transB: B is stored [N][K] (else [K][N]); w [K] or NULL.  Host pointers. */
int tx_real_gemm(const double* A, const double* B, const double* w, double* C, int M, int N, int K, int transA, int transB);
/* bardeen.current (:537-575, nFD :1014-1020): I[p][alpha][beta][i] = 2 pi sum_m stat(E[p][alpha][m], Vb[i])
 * M2[p][alpha][m][beta], stat = fL (1 - fR) - fR (1 - fL), muL = muR + Vb[i].  E [P][n][nb], M2 [P][n][nb][n]. */
int tx_bardeen_current(const double* E, const double* M2, int P, int n, int nb, const double* Vb, int nVb, double muR,
                       double kBT, double* I);
int tx_bardeen_current_dev(const double* E, const double* M2, int P, int n, int nb, const double* Vb, int nVb, double muR,
                           double kBT, double* I, void* stream);
/* bardeen.Ts_bardeen (:577-632): T[alpha][m][beta] = NL/(k tL[alpha]) NR/tR[beta] dos M2[alpha][m][beta];
 * per-channel lead parameters tL, VL, tR, VR [n].  Outside the band k and dos take the real parts of numpy's
 * complex branches (k = 0 or pi, dos = 0, on BOTH sides of the band).  flags [2 + n*n]: flags[0] = number of wavenumbers
 * outside the band, flags[1] = 0 (reserved), flags[2 + alpha*n + beta] = number of m with a real density of states;
 * the reference (:611, :624, max of the imaginary parts) raises ValueError when flags[0] > 0 or one of the per-(alpha,
 * beta) counts is 0. */
int tx_bardeen_ts(const double* E, const double* M2, int n, int nb, const double* tL, const double* VL, const double* tR,
                  const double* VR, int NL, int NR, double* T, int* flags);

/* Process-wide defaults of the tuning options (contexts created later inherit them; the default contexts follow them).
 *   option 0: value != 0 (default) enables the structure scan of the on-site blocks: diagonal sites take the
 *             column-scaling step of the telescoped sweep (the plain sweep uses it for its diagonal right tail
 *             and scalar runs); value 0 treats every site as a general block (A/B measurements).
 *   option 1: longest chunk of the telescoped sweep in sites, 1..64 (default 32; 1 = one inverse per site).
 *   option 2: log2 of the growth bound that closes a chunk early, 0..500 (default 4: max |D_j| / prod |t| < 32).
 *   option 3: eigenvalue kernel of the Bardeen path: 0 (default) picks by batch size, 1 a warp per eigenvalue
 *             (multisection, shortest critical path), 2 a lane per eigenvalue (bracketed interpolation, least arithmetic; chosen
 *             automatically from 2048 state slots per call).  Process-wide only.
 *   option 5: n = 64 and 96 <= n <= 104 blocks in the global workspace: value != 0 (default) stages the B operand of every product in shared
 *             memory with the TMA engine (cp.async.bulk + mbarrier), 0 reads all fragments straight from L2 (A/B runs).
 *             Process-wide only.
 *   option 6: Sancho-Rubio decimation: value != 0 (default) runs two products per iteration instead of six when the hopping
 *             block is Hermitian (alpha = beta throughout; detected on the device), 0 always the general iteration (A/B runs).
 *             Process-wide only.
 *   option 4: kernel of the 5 <= n <= 8, scalar-hopping sweep (= tx_set_variant):
 *             13  DEFAULT: telescoped sweep (rgf_n8t.cuh): register-resident DMMA recurrence, one inverse per chunk
 *                 of up to kmax sites, diagonal sites as column scalings; eight energies of one Hamiltonian per warp
 *                 (energy axes shorter than 64 that are not a multiple of eight fall back to 6)
 *              6  plain sweep, one inverse per site, tensor-core (DMMA) products, W in shared memory (rgf_n8.cuh)
 *              0  register-resident products (rgf_small.cuh), 4 lanes per point — the kernel general hopping uses
 *             All three compute the same thing (test_all_kernel_variants_agree). */
int tx_set_variant(int variant);
int tx_set_option(int option, int value);

/* Unit-test hook: out[b] = inverse(in[b]) for `count` n x n blocks (n <= 16) using the device
 * routine of the sweep (in-place Gauss-Jordan with implicit partial pivoting). Host pointers. */
int tx_debug_invert(const tx_cplx* in, tx_cplx* out, int count, int n);

/* Same for the CTA-cooperative inverse of the large-block paths (n <= 128; blocked tensor-core Gauss-Jordan with
 * implicit partial pivoting when n is a multiple of 8 and >= 16).  Returns TX_ERR_SINGULAR on a zero pivot. */
int tx_debug_invert_staged(const tx_cplx* in, tx_cplx* out, int count, int n);
/* Measurement hook: device time of `count` CTAs each doing `reps` n x n complex products on global-memory blocks. */
int tx_debug_gemm_ms(int n, int count, int reps, double* ms_out);
/* The same with the B operand staged in shared memory by TMA (cp.async.bulk + mbarrier; n a multiple of 8). */
int tx_debug_gemm_tma_ms(int n, int count, int reps, double* ms_out);
/* Device time in ms of the last tx_debug_invert_staged kernel launch (CUDA events; measurement hook). */
double tx_debug_last_invert_ms(void);

/* ---- measurement helpers --------------------------------------------------------------- */
/*
 * Measures the FP64 FMA-pipe peak of the current device with a register-resident DFMA loop
 * (no memory traffic).  Writes TFLOP/s (2 flops per FMA) to *tflops.  Used only to provide
 * the roofline denominator that MEASURED_PEAKS.json lacks (it has HBM and bf16 only).
 */
int tx_measure_fp64_peak(double* tflops, double* milliseconds);
/* Same for the legacy FP64 tensor path (mma.sync.m8n8k4.f64). */
int tx_measure_dmma_peak(double* tflops, double* milliseconds);
/* Exchange-primitive rates: which = 0 SHFL.b32, 1 LDS.128 group broadcast (1e9 warp-instr/s),
 * 2 same with a 2-way conflict, 3 DFMA:LDS 8:1 mix (TFLOP/s), 4 DFMA+DMMA mix with equal flops (TFLOP/s). */
int tx_microbench(int which, double* value);

#ifdef __cplusplus
}
#endif
#endif /* TRANSPORT_B200_H */
