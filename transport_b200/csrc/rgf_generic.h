// Interface of the generic-size sweep (rgf_generic.cu) for capi.cu.
#pragma once
#include <cuda_runtime.h>

#include "tx_common.cuh"

namespace tx {

struct GenericParams {
    const cd* h;
    long long h_stride;
    const cd* h1;
    const double* params;
    const cd* tnn;
    long long t_stride;
    const cd* E;
    const cd* sigL;
    const cd* sigR;
    long long sig_stride;
    const double* sources;   // [n_src][n] or null (all one-hot)
    double* R;               // [n_pts][n_src][n] or null
    double* T;
    double* resid;           // [n_pts][n_src] or null
    double* ttot;            // [n_pts] or null
    double* caroli;          // [n_pts] or null
    double* spin;            // [n_pts][n_src][spin_n] or null (see emit_spin in rgf_small.cuh)
    int n_up, spin_n;
    PeerGather gather;
    const int* site_class;   // structure scan (rgf_n8.cuh): [n_h][M], >= 1: on-site block (and its affine term) diagonal; or null
    long long class_stride;  // M, or 0 when all Hamiltonians share the blocks
    cd* work;                // global workspace or null (then dynamic shared memory)
    long long work_per_cta;  // elements of `work` per resident CTA (set by the launcher)
    int* singular;
    long long n_pts;
    int n, M, nE, n_src, hop_scalar;
    int np;                  // edge of the work blocks (set by the launcher: n, or n padded to a multiple of 8)
    int kmax, gexp;          // telescoped sweep (scalar hopping): longest chunk, log2 of the growth bound
    int use_tma;             // n = 64: TMA-staged product operand (set by the launcher from tx_set_option(5, .))
};

size_t rgf_generic_workspace_elems(int n, long long n_pts);
int launch_rgf_generic(GenericParams p, cudaStream_t st);

}  // namespace tx
