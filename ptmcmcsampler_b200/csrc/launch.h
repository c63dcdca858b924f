// Host-side launch entry points of the MH kernels.  Each kernel family is compiled in its own translation
// unit (mh_sorted.cu, mh_mma.cu) so that the library builds in parallel; engine.cu sees only these.
#pragma once
#include <cuda_runtime.h>

#include "params.h"

namespace ptm {

// ---- sorted shared-memory kernel (mh_sorted_kernel.cuh), ndim <= 32 -------------------------------------
// static tables of the target, built once per engine on the host and handed to the kernel as a parameter
struct SortedHostTables {
    double P[MAX_REG_DIM * MAX_REG_DIM];  // upper-triangular Gaussian form (-1/2 folded in), row-major d x d
    double mu[MAX_REG_DIM], lo[MAX_REG_DIM], hi[MAX_REG_DIM];  // mean; prior box as an inclusive interval
    int d;
};
struct SortedGeom {
    int cfg;       // index into the instantiated (chains per block, blocks per SM) table
    int nc;        // chains per block actually used (<= the instance's capacity)
    int blocks;    // grid size
    int cap, minb;
};
// pick the launch geometry for `chains` chains of dimension d on a device with `sms` multiprocessors
// (cfg_request < 0: the default; PTMCMC_SORT_CFG / PTMCMC_SORT_NC override for experiments)
SortedGeom sorted_geometry(int d, long long chains, int sms, int cfg_request, int nc_request);
cudaError_t launch_sorted(const DevParams &p, const SortedHostTables &tb, const SortedGeom &g, int device, cudaStream_t stream);
const char *sorted_kernel_name(int d, const SortedGeom &g);

// ---- tensor-core (DMMA) kernel (mh_mma_kernel.cuh), dense Gaussian target, ndim <= 128 -----------------
struct MmaGeom {
    int nt;        // n-tiles (ndim <= 8 nt), 0 = not available
    int nc, ld, smem;
    bool split;    // ndim > 32: mh_mma_split_kernel (two blocks of <= 32 chains per SM); set by mma_geometry
    bool tri;      // the form is held as a Cholesky factor (set before mma_geometry: the split kernels pack its tiles)
};
int mma_pick_nt(int d);
int mma_pf_tiles(const MmaGeom &g);
void mma_geometry(MmaGeom &g, int nc_request);
cudaError_t launch_mma(const DevParams &p, const MmaGeom &g, const double *Uf, const double *Pf, const double *Ut,
                       int device, cudaStream_t stream);
cudaError_t launch_frag_build(const double *src, int d, int nt, int transpose, int packed, double *out, cudaStream_t stream);
cudaError_t launch_transpose(const double *src, int d, double *dst, cudaStream_t stream);

}  // namespace ptm
