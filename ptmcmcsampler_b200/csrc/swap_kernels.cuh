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

// dst[k W] = src[k W], k < d, with four loads in flight (d is a run-time value: the plain loop waits for each load)
__device__ __forceinline__ void copy_strided(double *dst, const double *src, int d, int W)
{
    int k = 0;
    for (; k + 4 <= d; k += 4) {
        const double v0 = src[(size_t)k * W], v1 = src[(size_t)(k + 1) * W], v2 = src[(size_t)(k + 2) * W],
                     v3 = src[(size_t)(k + 3) * W];
        dst[(size_t)k * W] = v0;
        dst[(size_t)(k + 1) * W] = v1;
        dst[(size_t)(k + 2) * W] = v2;
        dst[(size_t)(k + 3) * W] = v3;
    }
    for (; k < d; ++k) dst[(size_t)k * W] = src[(size_t)k * W];
}

// The sweep is a dependent chain over the rungs of a walker; everything about a pair that does not depend on the
// travelling state is computed for all pairs at once beforehand (thread per pair and walker): the accept uniform and
// the two terms of the log acceptance ratio that hold the lower rung's own lnL.  prep[0][sc][w] = u of pair (sc, sc+1) (and prep[3] = log u)
// (global pair g uses word Tg-2-g of the swap stream: hottest pair first), prep[1] = -La/Ta, prep[2] = La/Tb with La the
// lnL at local rung sc.  Pair index T-1 is the boundary pair with the hotter shard (Tb = ladder_above), present when
// has_above.  The chain that remains per pair is two divisions, three additions, exp and a compare (ref :673-679, same
// four-term order, see swap_accept below).
__global__ void __launch_bounds__(256) swap_prep_kernel(const DevParams p, long long it, int Tg, double ladder_above,
                                                        int has_above, double *prep)
{
    const int W = p.W, T = p.T;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)T * W) return;
    const int sc = (int)(idx / W), w = (int)(idx % W);
    if (sc == T - 1 && !has_above) return;
    const double La = p.lnl[idx];
    const double Ta = p.ladder[sc], Tb = (sc == T - 1) ? ladder_above : p.ladder[sc + 1];
    Stream st(p, PURPOSE_SWAP, (unsigned long long)it, (uint32_t)(p.walker_offset + w), 0u);
    st.seek((uint32_t)(Tg - 2 - (p.temp_offset + sc)));
    const size_t C = (size_t)T * W;
    const double u = word_to_unit(st.next());
    prep[idx] = u;
    prep[C + idx] = -La / Ta;
    prep[2 * C + idx] = La / Tb;
    prep[3 * C + idx] = log(u);
}

// pair sc of walker w against the travelling state's lnL (Lb): u <= exp(-La/Ta - Lb/Tb + Lb/Ta + La/Tb).
// exp is monotone and both exp and log are good to an ulp, so unless lar and log(u) agree to 1e-12 the comparison of the
// logarithms decides exactly as the reference's does; the exponential is evaluated only in that band (and for NaN / inf).
__device__ __forceinline__ bool swap_accept_prep(double u, double logu, double a1, double a4, double Lb, double Ta, double Tb)
{
    double lar = a1;
    lar += -Lb / Tb;
    lar += Lb / Ta;
    lar += a4;
    const double gap = lar - logu, band = 1e-12 * fmax(1.0, fabs(lar));
    if (gap > band) return true;
    if (gap < -band) return false;
    return u <= exp(lar);
}

// The local pairs (T-2 .. 0) of walker w, top-down, starting from the state `carry` (lnL Lcarry) at position T-1.
// map_out[t][w] = rung whose state moves to rung t: because the sweep visits each pair once, only the state currently
// sitting at position sc+1 (the "carry") is ever displaced.  The operands of SWEEP_CHUNK pairs are loaded together (their
// addresses do not depend on the chain), so the chain pays one memory round trip per chunk instead of one per pair.
constexpr int SWEEP_CHUNK = 8;
__device__ __forceinline__ void sweep_local_pairs(const DevParams &p, const double *prep, int w, int &carry, double &Lcarry,
                                                  int *map_out)
{
    const int W = p.W, T = p.T;
    const size_t C = (size_t)T * W;
    for (int sc0 = T - 2; sc0 >= 0; sc0 -= SWEEP_CHUNK) {
        double u[SWEEP_CHUNK], lu[SWEEP_CHUNK], a1[SWEEP_CHUNK], a4[SWEEP_CHUNK], La[SWEEP_CHUNK], Ta[SWEEP_CHUNK], Tb[SWEEP_CHUNK];
#pragma unroll
        for (int j = 0; j < SWEEP_CHUNK; ++j) {
            const int sc = sc0 - j;
            if (sc >= 0) {
                const size_t idx = (size_t)sc * W + w;
                u[j] = prep[idx];
                a1[j] = prep[C + idx];
                a4[j] = prep[2 * C + idx];
                lu[j] = prep[3 * C + idx];
                La[j] = p.lnl[idx];  // swap_map[sc] == sc: not visited yet
                Ta[j] = p.ladder[sc];
                Tb[j] = p.ladder[sc + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < SWEEP_CHUNK; ++j) {
            const int sc = sc0 - j;
            if (sc >= 0) {
                const size_t idx = (size_t)sc * W + w;
                if (swap_accept_prep(u[j], lu[j], a1[j], a4[j], Lcarry, Ta[j], Tb[j])) {  // ref :679-681
                    map_out[idx + W] = sc;
                    atomicAdd(&p.swap_acc[idx], 1ull);  // a reduction: the chain does not wait for the counter's old value
                } else {
                    map_out[idx + W] = carry;
                    carry = sc;
                    Lcarry = La[j];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) swap_decide_kernel(const DevParams p, const double *prep, int *map_out,
                                                          short *swapmap_trace)
{
    const int W = p.W, T = p.T;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    int carry = T - 1;
    double Lcarry = p.lnl[(size_t)(T - 1) * W + w];
    sweep_local_pairs(p, prep, w, carry, Lcarry, map_out);
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
    copy_strided(xd, xs, d, W);
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

// One hop of the neighbour exchange through peer memory (NVLink loads / stores from the kernels themselves instead of a
// send / receive pair): a kernel that consumes a message spins on a flag in its OWN memory until the neighbour has set
// it to this swap's sequence number; a kernel that produces one writes it straight into the neighbour's mailbox and the
// last block to finish publishes the sequence number there.  A wait that times out sets *err and goes on (the run is
// then invalid, but nothing hangs).
struct P2PSync {
    const unsigned long long *wait_flag;  // nullptr: nothing to wait for
    unsigned long long *signal_flag;      // nullptr: nothing to publish
    unsigned *done_ctr;                   // blocks that have finished writing (reset by the last one)
    unsigned long long seq;
    int *err;
};
constexpr long long P2P_TIMEOUT_CLOCKS = 4000000000ll;  // about two seconds

__device__ __forceinline__ void p2p_wait(const P2PSync &s)
{
    if (s.wait_flag) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<const volatile unsigned long long *>(s.wait_flag) < s.seq) {
                if (clock64() - t0 > P2P_TIMEOUT_CLOCKS) {
                    *s.err = 1;
                    break;
                }
                __nanosleep(64);
            }
            __threadfence_system();
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void p2p_signal(const P2PSync &s)
{
    if (s.signal_flag) {
        __threadfence_system();  // this thread's stores to the neighbour's memory
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(s.done_ctr, 1u) == gridDim.x - 1) {
            *s.done_ctr = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(s.signal_flag) = s.seq;
        }
    }
}

// the wait alone, ahead of a kernel of many blocks (they would all spin on the flag and hold the SMs meanwhile)
__global__ void p2p_wait_kernel(const P2PSync sync) { p2p_wait(sync); }

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

__global__ void __launch_bounds__(256) swap_pack_top_kernel(const DevParams p, double *msg, const P2PSync sync)
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
    p2p_signal(sync);
}

// thread per walker: boundary pair with the hotter shard (if any), then the local pairs, top-down, on the terms of
// swap_prep_kernel (launched when the segment ended, long before the carry arrives).
__global__ void __launch_bounds__(128) swap_sweep_kernel(const DevParams p, double ladder_above, const double *prep,
                                                         const double *carry_in, double *carry_out, int *map_out,
                                                         int *carry_code, double *carry_L, const P2PSync sync)
{
    const int d = p.d, W = p.W, T = p.T;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    p2p_wait(sync);
    if (w < W) {
    const size_t C = (size_t)T * W;
    int carry = T - 1;
    double Lcarry = p.lnl[(size_t)(T - 1) * W + w];
    if (carry_in) {
        const size_t idx = (size_t)(T - 1) * W + w;
        const double Lb = carry_in[(size_t)d * W + w];
        if (swap_accept_prep(prep[idx], prep[3 * C + idx], prep[C + idx], prep[2 * C + idx], Lb, p.ladder[T - 1], ladder_above)) {
            atomicAdd(&p.swap_acc[idx], 1ull);  // a reduction: the chain does not wait for the counter's old value
            carry = T;  // the foreign state keeps travelling down
            Lcarry = Lb;
        }
    }
    sweep_local_pairs(p, prep, w, carry, Lcarry, map_out);
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
    p2p_signal(sync);
}

// thread per chain: position 0 resolves the boundary pair with the colder shard (the same decision
// that shard took in its sweep), then every position copies its source into the second state buffer
// and does the iteration's updateChains (ref :627).
__global__ void __launch_bounds__(MH_THREADS) swap_finish_kernel(const DevParams p, long long it, int Tg,
                                                                  double ladder_below, const int *map,
                                                                  const int *carry_code, const double *carry_L,
                                                                  const double *carry_in, const double *below_top,
                                                                  double *x_new, double *lnl_new, double *lp_new,
                                                                  short *swapmap_trace, const P2PSync sync)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    p2p_wait(sync);
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
    copy_strided(xd, xs, d, W);
    lnl_new[c] = lnl;
    lp_new[c] = lp;
    if (swapmap_trace)
        swapmap_trace[(size_t)w * T + t] = (short)(msg ? (int)msg[(size_t)(d + 2) * W + w] : p.temp_offset + code);
    bookkeep(p, it, t, w, [&](int k) { return xs[(size_t)k * W]; }, lnl, lp, 1.0 / p.mh_temp[t]);
}

}  // namespace ptm
