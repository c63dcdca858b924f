// Instantiations and launch geometry of the sorted shared-memory MH kernel (one translation unit).
#include <cstdio>
#include <cstdlib>

#include "launch.h"
#include "mh_sorted_kernel.cuh"

namespace ptm {
namespace {

// (chains = threads per block, blocks per SM the register allocation must allow)
struct Cfg { int nc, minb; };
constexpr Cfg CFGS[] = {
    {128, 5},  // 0: five 4-warp blocks per SM, 96 registers (measured best on C2: 2.92 ms vs 3.10 ms for 256 x 2)
    {256, 2},  // 1: two 8-warp blocks per SM, 128 registers
    {128, 4},  // 2: four 4-warp blocks per SM, 128 registers
    {128, 6},  // 3: 80 registers
    {160, 4},  // 4: 96 registers, 5-warp blocks
    {96, 6},   // 5: 96 registers, 3-warp blocks
};
constexpr int NCFG = sizeof(CFGS) / sizeof(CFGS[0]);
constexpr int DEFAULT_CFG = 0;

template <int DP, int C, int LK>
cudaError_t launch_lk(const DevParams &p, const SortedHostTables &ht, const SortedGeom &g, int device, cudaStream_t stream)
{
    constexpr Cfg c = CFGS[C];
    using Smem = SortedSmem<DP, c.nc>;
    auto kern = mh_sorted_kernel<DP, c.nc, c.minb, LK>;
    static bool attr_dev[64] = {};
    bool &attr_done = attr_dev[device & 63];  // function attributes are per device
    if (!attr_done) {
        cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (st != cudaSuccess) return st;
        if (const char *v = getenv("PTMCMC_SORT_CARVEOUT"))  // experiment: shared-memory carve-out in percent
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(v));
        attr_done = true;
    }
    SortedTables<DP> tb;
    const int d = ht.d;
    for (int i = 0; i < DP; ++i) {
        for (int j = i; j < DP; ++j) tb.P[tri_index<DP>(i, j)] = (i < d && j < d) ? ht.P[i * d + j] : 0.0;
        tb.mu[i] = i < d ? ht.mu[i] : 0.0;
        tb.lo[i] = i < d ? ht.lo[i] : -__builtin_inf();
        tb.hi[i] = i < d ? ht.hi[i] : __builtin_inf();
    }
    kern<<<g.blocks, c.nc, sizeof(Smem), stream>>>(p, tb, g.nc);
    return cudaGetLastError();
}

template <int DP, int C>
cudaError_t launch_one(const DevParams &p, const SortedHostTables &ht, const SortedGeom &g, int device, cudaStream_t stream)
{
    switch (p.logl_kind) {
    case LOGL_GAUSSIAN: return launch_lk<DP, C, LOGL_GAUSSIAN>(p, ht, g, device, stream);
    case LOGL_CURVED: return launch_lk<DP, C, LOGL_CURVED>(p, ht, g, device, stream);
    default: return launch_lk<DP, C, LOGL_ROSENBROCK>(p, ht, g, device, stream);
    }
}

template <int DP>
cudaError_t launch_dp(const DevParams &p, const SortedHostTables &ht, const SortedGeom &g, int device, cudaStream_t stream)
{
#ifdef PTMCMC_ALL_SORT_CFGS
    if constexpr (DP == 20) {  // the experimental geometries exist for the benchmark width only
        switch (g.cfg) {
        case 1: return launch_one<DP, 1>(p, ht, g, device, stream);
        case 2: return launch_one<DP, 2>(p, ht, g, device, stream);
        case 3: return launch_one<DP, 3>(p, ht, g, device, stream);
        case 4: return launch_one<DP, 4>(p, ht, g, device, stream);
        case 5: return launch_one<DP, 5>(p, ht, g, device, stream);
        default: break;
        }
    }
#endif
    return launch_one<DP, DEFAULT_CFG>(p, ht, g, device, stream);
}

}  // namespace

SortedGeom sorted_geometry(int d, long long chains, int sms, int cfg_request, int nc_request)
{
    SortedGeom g{};
    int cfg = cfg_request;
    if (const char *v = getenv("PTMCMC_SORT_CFG")) cfg = atoi(v);
#ifdef PTMCMC_ALL_SORT_CFGS
    if (d <= 16 || d > 20) cfg = DEFAULT_CFG;
#else
    cfg = DEFAULT_CFG;
#endif
    if (cfg < 0 || cfg >= NCFG) cfg = DEFAULT_CFG;
    const Cfg c = CFGS[cfg];
    g.cfg = cfg; g.minb = c.minb; g.cap = c.nc;
    (void)sms;
    long long nc = c.nc;
    if (const char *v = getenv("PTMCMC_SORT_NC")) nc_request = atoi(v);
    if (nc_request >= 1 && nc_request <= c.nc) nc = nc_request;
    if (nc > chains) nc = chains;
    g.nc = (int)nc;
    g.blocks = (int)((chains + nc - 1) / nc);
    return g;
}

cudaError_t launch_sorted(const DevParams &p, const SortedHostTables &ht, const SortedGeom &g, int device, cudaStream_t stream)
{
    const int d = p.d;
    if (d <= 4) return launch_dp<4>(p, ht, g, device, stream);
    if (d <= 8) return launch_dp<8>(p, ht, g, device, stream);
    if (d <= 12) return launch_dp<12>(p, ht, g, device, stream);
    if (d <= 16) return launch_dp<16>(p, ht, g, device, stream);
    if (d <= 20) return launch_dp<20>(p, ht, g, device, stream);
    if (d <= 24) return launch_dp<24>(p, ht, g, device, stream);
    return launch_dp<32>(p, ht, g, device, stream);
}

const char *sorted_kernel_name(int d, const SortedGeom &g)
{
    static thread_local char buf[96];
    const int dp = d <= 4 ? 4 : d <= 8 ? 8 : d <= 12 ? 12 : d <= 16 ? 16 : d <= 20 ? 20 : d <= 24 ? 24 : 32;
    snprintf(buf, sizeof buf, "mh_sorted_kernel<%d,%d,%d>", dp, g.cap, g.minb);
    return buf;
}

}  // namespace ptm
