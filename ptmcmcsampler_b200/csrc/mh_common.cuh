// Device helpers shared by every MH kernel: jump selection, proposal scales, DE-history addressing, the
// per-chain buffer / record bookkeeping (ref PTMCMCSampler.py _jump :1048-1067, :843-862, updateChains :321-335).
#pragma once
#include "params.h"
#include "rng.cuh"

namespace ptm {

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xFFF0000000000000LL); }
__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7FF0000000000000LL); }

__device__ __forceinline__ int pick_jump(const DevParams &p, Stream &st)
{
    // ind = integers(0, len(propCycle)) over the weight-replicated cycle (ref :1007-1008, :1058)
    const int ind = (int)word_to_int(st.next(), (unsigned long long)p.total_weight);
    int jump = p.cyc_jump[p.ncycle - 1];
    for (int i = 0; i < p.ncycle; ++i)
        if (ind < p.cyc_cum[i]) { jump = p.cyc_jump[i]; break; }
    return jump;
}

__device__ __forceinline__ double cov_jump_scale(double prob, double temp)
{
    // ref :843-862 / :900-920
    double scale = (prob > 0.97) ? 10.0 : (prob > 0.9) ? 0.2 : 1.0;
    if (temp <= 100.0) scale *= sqrt(temp);
    return scale;
}

__device__ __forceinline__ bool in_box(double v, double lo, double hi, int inclusive)
{
    return inclusive ? (lo <= v && hi >= v) : (lo < v && hi > v);
}

__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// 16-byte read-only load that does not allocate in L1: streams (history rows, fragment images larger than L1) must not
// evict what lives there (the kernel's local-memory lines, small tables)
__device__ __forceinline__ double2 ldg_stream(const double2 *ptr)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(ptr));
    return v;
}

// physical row of logical DE-history row r (ring with head slot); 32-bit arithmetic when it fits
__device__ __forceinline__ unsigned long long de_row_offset(unsigned long long r, unsigned long long bufsize, int W,
                                                            long long burn, long long head)
{
    unsigned long long slot, wsel;
    if (bufsize <= 0xFFFFFFFFull) {
        const unsigned r32 = (unsigned)r, s32 = r32 / (unsigned)W;
        slot = s32;
        wsel = r32 - s32 * (unsigned)W;
    } else {
        slot = r / (unsigned long long)W;
        wsel = r - slot * (unsigned long long)W;
    }
    slot += (unsigned long long)head;
    if (slot >= (unsigned long long)burn) slot -= (unsigned long long)burn;
    return slot * (unsigned long long)W + wsel;
}

// updateChains (ref :321-335) for one chain: AM ring slot for the cold rung, thinned record.
template <typename XGet>
__device__ __forceinline__ void bookkeep(const DevParams &p, long long it, int t, int w, XGet xget,
                                         double lnl, double lp, double beta)
{
    const int d = p.d, W = p.W;
    if (t == 0 && p.temp_offset == 0 && p.am) {
        double *dst = p.am + (size_t)(it % p.cov_update) * d * W + w;
        for (int k = 0; k < d; ++k) dst[(size_t)k * W] = xget(k);
    }
    if (t < p.ntr && it % p.thin == 0) {
        const long long row = it / p.thin - p.rec_base;
        if (row >= 0 && row < p.rec_cap) {
            const size_t r = ((size_t)row * p.ntr + t) * W + w;
            double *dst = p.rec_x + r * d;
            for (int k = 0; k < d; ++k) dst[k] = xget(k);
            p.rec_lnl[r] = lnl;
            p.rec_lnp[r] = beta * lnl + lp;
        }
    }
}

}  // namespace ptm
