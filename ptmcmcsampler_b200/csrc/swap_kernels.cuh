// Inter-temperature swap (ref PTMCMCSampler.py PTswap :631-697) for W walkers at once.
//
// The reference gathers (lnL, x) of all rungs to rank 0, sweeps the adjacent pairs hottest first
// with one uniform per pair, and scatters the permuted states back.  Here every walker's ladder is
// swept by one thread (the sweep is sequentially dependent across rungs, independent across
// walkers), then the permutation is applied by one thread per chain into the second state buffer.
#pragma once
#include "params.h"
#include "rng.cuh"

namespace ptm {

// map_out[t][w] = rung whose state moves to rung t.  Because the sweep visits each pair once,
// top-down, only the state currently sitting at position sc+1 (the "carry") is ever displaced.
__global__ void __launch_bounds__(128) swap_decide_kernel(const DevParams p, long long it, int *map_out,
                                                          short *swapmap_trace)
{
    const int W = p.W, T = p.T;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    Stream st(p.seed, PURPOSE_SWAP, (unsigned long long)it, (uint32_t)(p.walker_offset + w), 0u);
    int carry = T - 1;
    double Lcarry = p.lnl[(size_t)(T - 1) * W + w];
    for (int sc = T - 2; sc >= 0; --sc) {
        const double La = p.lnl[(size_t)sc * W + w];  // swap_map[sc] == sc: not visited yet
        const double Lb = Lcarry;
        const double Ta = p.ladder[sc], Tb = p.ladder[sc + 1];
        // ref :673-676, same four-term order
        double lar = -La / Ta;
        lar += -Lb / Tb;
        lar += Lb / Ta;
        lar += La / Tb;
        const double ratio = exp(lar);
        const double u = word_to_unit(st.next());
        if (u <= ratio) {  // ref :679-681
            map_out[(size_t)(sc + 1) * W + w] = sc;
            p.swap_acc[(size_t)sc * W + w] += 1;
        } else {
            map_out[(size_t)(sc + 1) * W + w] = carry;
            carry = sc;
            Lcarry = La;
        }
    }
    map_out[w] = carry;
    if (swapmap_trace)
        for (int j = 0; j < T; ++j) swapmap_trace[(size_t)w * T + j] = (short)map_out[(size_t)j * W + w];
}

// new state of chain (t, w) = old state of chain (map[t][w], w) (ref :684-691); lnprob is
// re-derived as lnlike/temp + logp (ref :695; the log-prior value travels with the state).
// Also performs updateChains for this iteration (ref :627), which the reference runs after the swap.
__global__ void __launch_bounds__(MH_THREADS) swap_apply_kernel(const DevParams p, long long it, const int *map,
                                                                 double *x_new, double *lnl_new, double *lp_new)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const int src = map[c];
    const double *xs = p.x + (size_t)src * d * W + w;
    double *xd = x_new + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) xd[(size_t)k * W] = xs[(size_t)k * W];
    const double lnl = p.lnl[(size_t)src * W + w], lp = p.lp[(size_t)src * W + w];
    lnl_new[c] = lnl;
    lp_new[c] = lp;
    bookkeep(p, it, t, w, [&](int k) { return xs[(size_t)k * W]; }, lnl, lp, 1.0 / p.mh_temp[t]);
}

}  // namespace ptm
