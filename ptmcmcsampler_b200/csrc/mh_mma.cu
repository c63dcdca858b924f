// Instantiations and launch geometry of the tensor-core (DMMA) MH kernel (one translation unit).
#include <cstdlib>

#include "launch.h"
#include "mh_mma_split_kernel.cuh"

namespace ptm {
namespace {

constexpr int MMA_SMALL_MINB = 4;  // ndim <= 32: four 256-thread blocks per SM (<= 64 registers per thread)

template <int NT>
cudaError_t launch_nt(const DevParams &p, const MmaGeom &g, const double *Uf, const double *Pf, const double *Ut, int device,
                      cudaStream_t stream)
{
    constexpr bool USMEM = NT <= 4;
    constexpr int MINB = NT <= 4 ? MMA_SMALL_MINB : 1;
    static bool attr_dev[64][2] = {};
    bool &attr_done = attr_dev[device & 63][g.split ? 1 : 0];  // function attributes are per device
    const int pf_tiles = mma_pf_tiles(g);
    MmaArgs a{Uf, Pf, Ut, g.nc, g.ld, g.tri ? 1 : 0, pf_tiles, mma_layout(NT, g.nc, g.ld, USMEM, pf_tiles)};
    const int blocks = (int)(((long long)p.T * p.W + g.nc - 1) / g.nc);
    if constexpr (NT >= 8) {
        if (g.split) {
            if (!attr_done) {
                cudaError_t st = cudaFuncSetAttribute(mh_mma_split_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (st != cudaSuccess) return st;
                attr_done = true;
            }
            mh_mma_split_kernel<NT><<<blocks, MMA_THREADS, g.smem, stream>>>(p, a);
            return cudaGetLastError();
        }
    }
    if (!attr_done) {
        cudaError_t st = cudaFuncSetAttribute(mh_mma_kernel<NT, USMEM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              227 * 1024);
        if (st != cudaSuccess) return st;
        attr_done = true;
    }
    mh_mma_kernel<NT, USMEM, MINB><<<blocks, MMA_THREADS, g.smem, stream>>>(p, a);
    return cudaGetLastError();
}

}  // namespace

// instantiated n-tile counts (ndim <= 8 NT); the smallest one that covers ndim is used
int mma_pick_nt(int d)
{
    const int need = (d + 7) / 8;
    for (int nt : {1, 2, 3, 4, 8, 13, 16})
        if (nt >= need) return nt;
    return 0;
}

// 8x8 tiles of the form's fragment image: the split kernels (NT >= 8) keep only the lower triangle of a Cholesky factor
int mma_pf_tiles(const MmaGeom &g) { return mma_tiles(g.nt, g.split && g.tri); }

// chains per block.  NT <= 4: MMA_SMALL_MINB blocks per SM.  NT >= 8 (split kernels): at most 32 chains (phase L runs two
// warps per 8-chain tile), two blocks per SM when the shared memory allows, else one.
void mma_geometry(MmaGeom &g, int nc_request)
{
    const int NT = g.nt, KP = 8 * NT;
    g.ld = (KP % 16 == 8) ? KP : KP + 8;
    const bool usmem = NT <= 4;
    const char *sv = getenv("PTMCMC_MMA_SPLIT");  // development A/B: 0 = the one-block-per-SM kernel for ndim > 32
    g.split = NT >= 8 && !(sv && atoi(sv) == 0);
    const int pf_tiles = mma_pf_tiles(g);
    auto fits = [&](int nc, int budget) { return mma_layout(NT, nc, g.ld, usmem, pf_tiles).total <= budget; };
    int nc = 0, nc_max = 256;
    if (NT <= 4) {
        for (int c : {128, 96, 64})
            if (!nc && fits(c, (228 / MMA_SMALL_MINB - 1) * 1024)) nc = c;
        for (int c : {256, 192, 128})
            if (!nc && fits(c, 113 * 1024)) nc = c;
    } else if (g.split) {
        nc_max = 32;
        for (int c : {32, 24, 16})
            if (!nc && fits(c, 113 * 1024)) nc = c;
    }
    if (!nc)
        for (int c = nc_max; c >= 8 && !nc; c -= 8)
            if ((c % 64 == 0 || c < 64) && fits(c, 227 * 1024)) nc = c;
    if (const char *v = getenv("PTMCMC_MMA_NC")) nc_request = atoi(v);
    if (nc_request >= 8 && nc_request <= nc_max && nc_request % 8 == 0 && fits(nc_request, 227 * 1024)) nc = nc_request;
    g.nc = nc;
    g.smem = nc ? mma_layout(NT, nc, g.ld, usmem, pf_tiles).total : 0;
    if (!nc) g.nt = 0;
}

cudaError_t launch_mma(const DevParams &p, const MmaGeom &g, const double *Uf, const double *Pf, const double *Ut, int device,
                       cudaStream_t stream)
{
    switch (g.nt) {
    case 1: return launch_nt<1>(p, g, Uf, Pf, Ut, device, stream);
    case 2: return launch_nt<2>(p, g, Uf, Pf, Ut, device, stream);
    case 3: return launch_nt<3>(p, g, Uf, Pf, Ut, device, stream);
    case 4: return launch_nt<4>(p, g, Uf, Pf, Ut, device, stream);
    case 8: return launch_nt<8>(p, g, Uf, Pf, Ut, device, stream);
    case 13: return launch_nt<13>(p, g, Uf, Pf, Ut, device, stream);
    default: return launch_nt<16>(p, g, Uf, Pf, Ut, device, stream);
    }
}

cudaError_t launch_frag_build(const double *src, int d, int nt, int transpose, int packed, double *out, cudaStream_t stream)
{
    const int n = mma_tiles(nt, packed != 0) * 64;
    frag_build_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, d, nt, transpose, packed, out);
    return cudaGetLastError();
}

cudaError_t launch_transpose(const double *src, int d, double *dst, cudaStream_t stream)
{
    transpose_kernel<<<(d * d + 255) / 256, 256, 0, stream>>>(src, d, dst);
    return cudaGetLastError();
}

}  // namespace ptm

#ifdef PTMCMC_MMA_CLOCKS
extern "C" void ptmcmc_debug_mma_clocks(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, ptm::g_mma_clk, sizeof(unsigned long long) * 32);
    if (reset) {
        unsigned long long z[32] = {};
        cudaMemcpyToSymbol(ptm::g_mma_clk, z, sizeof z);
    }
}
#endif
