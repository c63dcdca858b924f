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
    Stream st(p, PURPOSE_SWAP, (unsigned long long)it, (uint32_t)(p.walker_offset + w), 0u);
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

// ---------------------------------------------------------------------------------------------
// Ladder sharding: this engine holds rungs [temp_offset, temp_offset + T) of a ladder of Tg rungs.
// The sweep above is cut at the shard boundaries.  A message is (d+3)*W doubles:
// x[d][W], lnl[W], lnprior[W], origin rung[W].  Source codes of a local position: 0..T-1 = local
// rung, T = the carry received from the hotter shard, T+1 = the colder shard's top rung.
// ---------------------------------------------------------------------------------------------

// acceptance of the pair (lower rung: La at Ta, upper rung: Lb at Tb); ref :673-679, same term order.
// One definition for every caller: both shards of a boundary must reach the same bit.
__device__ __forceinline__ bool swap_accept(double La, double Lb, double Ta, double Tb, double u)
{
    double lar = -La / Ta;
    lar += -Lb / Tb;
    lar += Lb / Ta;
    lar += La / Tb;
    return u <= exp(lar);
}

__global__ void __launch_bounds__(256) swap_pack_top_kernel(const DevParams p, double *msg)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long n = (long long)(d + 3) * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx / W), w = (int)(idx % W);
        double v;
        if (k < d) v = p.x[((size_t)(T - 1) * d + k) * W + w];
        else if (k == d) v = p.lnl[(size_t)(T - 1) * W + w];
        else if (k == d + 1) v = p.lp[(size_t)(T - 1) * W + w];
        else v = (double)(p.temp_offset + T - 1);
        msg[idx] = v;
    }
}

// thread per walker: boundary pair with the hotter shard (if any), then the local pairs, top-down.
// The sweep draws one uniform per pair, hottest pair first: global pair sc uses word Tg-2-sc.
__global__ void __launch_bounds__(128) swap_sweep_kernel(const DevParams p, long long it, int Tg, double ladder_above,
                                                         const double *carry_in, double *carry_out, int *map_out,
                                                         int *carry_code, double *carry_L)
{
    const int d = p.d, W = p.W, T = p.T;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    Stream st(p, PURPOSE_SWAP, (unsigned long long)it, (uint32_t)(p.walker_offset + w), 0u);
    int carry = T - 1;
    double Lcarry = p.lnl[(size_t)(T - 1) * W + w];
    if (carry_in) {
        st.seek((uint32_t)(Tg - 2 - (p.temp_offset + T - 1)));
        const double Lb = carry_in[(size_t)d * W + w];
        if (swap_accept(Lcarry, Lb, p.ladder[T - 1], ladder_above, word_to_unit(st.next()))) {
            p.swap_acc[(size_t)(T - 1) * W + w] += 1;
            carry = T;  // the foreign state keeps travelling down
            Lcarry = Lb;
        }
    } else {
        st.seek((uint32_t)(Tg - 2 - (p.temp_offset + T - 2)));
    }
    for (int sc = T - 2; sc >= 0; --sc) {
        const double La = p.lnl[(size_t)sc * W + w];
        if (swap_accept(La, Lcarry, p.ladder[sc], p.ladder[sc + 1], word_to_unit(st.next()))) {
            map_out[(size_t)(sc + 1) * W + w] = sc;
            p.swap_acc[(size_t)sc * W + w] += 1;
        } else {
            map_out[(size_t)(sc + 1) * W + w] = carry;
            carry = sc;
            Lcarry = La;
        }
    }
    carry_code[w] = carry;
    carry_L[w] = Lcarry;
    if (carry_out) {
        if (carry == T) {
            for (int k = 0; k < d + 3; ++k) carry_out[(size_t)k * W + w] = carry_in[(size_t)k * W + w];
        } else {
            for (int k = 0; k < d; ++k) carry_out[(size_t)k * W + w] = p.x[((size_t)carry * d + k) * W + w];
            carry_out[(size_t)d * W + w] = p.lnl[(size_t)carry * W + w];
            carry_out[(size_t)(d + 1) * W + w] = p.lp[(size_t)carry * W + w];
            carry_out[(size_t)(d + 2) * W + w] = (double)(p.temp_offset + carry);
        }
    }
}

// thread per chain: position 0 resolves the boundary pair with the colder shard (the same decision
// that shard took in its sweep), then every position copies its source into the second state buffer
// and does the iteration's updateChains (ref :627).
__global__ void __launch_bounds__(MH_THREADS) swap_finish_kernel(const DevParams p, long long it, int Tg,
                                                                  double ladder_below, const int *map,
                                                                  const int *carry_code, const double *carry_L,
                                                                  const double *carry_in, const double *below_top,
                                                                  double *x_new, double *lnl_new, double *lp_new,
                                                                  short *swapmap_trace)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    int code;
    if (t == 0) {
        code = carry_code[w];
        if (below_top) {
            Stream st(p, PURPOSE_SWAP, (unsigned long long)it, (uint32_t)(p.walker_offset + w), 0u);
            st.seek((uint32_t)(Tg - 2 - (p.temp_offset - 1)));
            if (swap_accept(below_top[(size_t)d * W + w], carry_L[w], ladder_below, p.ladder[0], word_to_unit(st.next())))
                code = T + 1;
        }
    } else {
        code = map[c];
    }
    const double *msg = (code == T) ? carry_in : (code == T + 1) ? below_top : nullptr;
    const double *xs = msg ? msg + w : p.x + (size_t)code * d * W + w;
    const double lnl = msg ? msg[(size_t)d * W + w] : p.lnl[(size_t)code * W + w];
    const double lp = msg ? msg[(size_t)(d + 1) * W + w] : p.lp[(size_t)code * W + w];
    double *xd = x_new + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) xd[(size_t)k * W] = xs[(size_t)k * W];
    lnl_new[c] = lnl;
    lp_new[c] = lp;
    if (swapmap_trace)
        swapmap_trace[(size_t)w * T + t] = (short)(msg ? (int)msg[(size_t)(d + 2) * W + w] : p.temp_offset + code);
    bookkeep(p, it, t, w, [&](int k) { return xs[(size_t)k * W]; }, lnl, lp, 1.0 / p.mh_temp[t]);
}

}  // namespace ptm
