// Divergence-free fused MH kernel (the production path for ndim <= 32, identity parameter group).
//
// The three built-in proposals cost very different amounts (AM: d normals + a d x d mat-vec, SCAM and
// DE: O(d)), and every chain picks its proposal independently each iteration (ref _jump :1048-1067),
// so a thread-per-chain warp executes all three paths every step.  Because the draws are
// counter-based, a chain's jump kind for an iteration is known before any state is touched.  Per
// iteration `it` a block of nc chains therefore runs
//   B(it)     thread r steps the r-th chain of the concatenated lists (AM | SCAM | DE) of iteration it, so that
//             warps are proposal-uniform except at the two list boundaries, and does that chain's
//             buffer / record bookkeeping (ref updateChains :321-335) with the state it just decided;
//   A(it+1)   the same thread then draws the jump kind of iteration it+1 for the chain it OWNS (chain tid) and
//             appends it to that kind's list (warp-aggregated shared-memory atomics); this needs no chain state,
//             so the short SCAM / DE warps do it while the AM warps are still stepping;
// and ONE block barrier.  Lists are double-buffered, their lengths triple-buffered.  Chain state (x, lnL,
// lnprior, counters) lives in shared memory for the whole launch.  The static tables of the target -- the
// Gaussian form (upper triangle, -1/2 folded in), its mean and the prior box -- arrive as a kernel parameter
// and are read as constant-bank operands of the fp64 instructions: the shared-memory data pipe is the
// busiest unit of this kernel (57-65 % in the round-1 capture) and those tables were 40 % of its wavefronts.
// The eigen-factor changes at every covariance update on the device, so it stays in shared memory.
// Results are identical to the thread-per-chain kernel draw for draw.
#pragma once
#include "mh_common.cuh"

namespace ptm {

// static per engine; d < DP is padded with zeros / an unbounded interval
template <int DP>
struct SortedTables {
    double P[DP * (DP + 1) / 2];  // upper triangle of the Gaussian form, packed by rows
    double mu[DP];
    double lo[DP], hi[DP];        // prior box as an inclusive interval (exclusive bounds moved one ulp inwards)
};

template <int DP, int NC>
struct SortedSmem {
    double Us[DP * DP];                        // eigenvectors, row-major
    double sS[DP];
    double xs[DP * NC];                        // [k][chain]
    double lnl[NC], lp[NC];
    double beta[NC], sct[NC], dsc[NC];         // 1/T; sqrt(T) if T <= 100 else 1; 2.4/sqrt(2d) * sqrt(T)
    unsigned long long cnt[3 * NC];            // per jump: proposed (low word) | accepted (high word)
    int ct[NC], cw[NC];                        // rung and walker of each chain of the block
    unsigned short list[2][3 * NC];            // per-kind lists of two consecutive iterations
    int count[3][4];                           // list lengths (AM, SCAM, DE) of three consecutive iterations
};

template <int DP>
__device__ __forceinline__ constexpr int tri_index(int i, int j) { return i * DP - (i * (i - 1)) / 2 + (j - i); }

// log-likelihood of the built-in targets from a full proposal in registers
template <int DP, int LK>
__device__ __forceinline__ double logl_builtin(const DevParams &p, const SortedTables<DP> &tb, const double (&q)[DP])
{
    const int d = p.d;
    if (LK == LOGL_GAUSSIAN) {
        double dv[DP];
#pragma unroll
        for (int j = 0; j < DP; ++j) dv[j] = q[j] - tb.mu[j];
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < DP; ++i) {
            double row = 0.0;
#pragma unroll
            for (int j = i; j < DP; ++j) row = fma(tb.P[tri_index<DP>(i, j)], dv[j], row);
            acc = fma(dv[i], row, acc);
        }
        return acc + p.g_offset;
    }
    if (LK == LOGL_CURVED) {
        double tot = 0.0;
#pragma unroll
        for (int b = 0; b + 1 < DP; b += 2) {
            if (b + 1 < d) {
                const double a = q[b], y = q[b + 1];
                const double t0 = 9.0 + 4.0 * a * a + 9.0 * y;
                tot += log(exp(-a * a - t0 * t0) + 0.5 * exp(-8.0 * a * a - 8.0 * (y - 2.0) * (y - 2.0)));
            }
        }
        return tot;
    }
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i + 1 < DP; ++i) {
        if (i + 1 < d) {
            const double a = q[i + 1] - q[i] * q[i], b = 1.0 - q[i];
            tot -= 100.0 * a * a + b * b;
        }
    }
    return tot / 20.0;
}

// jump kind of iteration `it` for the chain this thread owns (ref :1058), appended to that kind's list
template <int NC>
__device__ __noinline__ void draw_kind(const DevParams &p, long long it, bool have, uint32_t gw, uint32_t gt, unsigned short *list,
                                       int *count)
{
    const int tid = threadIdx.x, lane = tid & 31;
    int kind = 3;  // none
    if (have) {
        Stream st(p, PURPOSE_MH, (unsigned long long)it, gw, gt);
        const int jump = pick_jump(p, st);
        kind = (jump == JUMP_AM) ? 0 : (jump == JUMP_SCAM) ? 1 : 2;
    }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == kk);
        if (m) {
            int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&count[kk], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (kind == kk) list[kk * NC + base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)tid;
        }
    }
}

// LK: the built-in log-likelihood this instance evaluates (one instance per target keeps the code resident in the
// instruction cache: the round-2 capture of a single kernel for all targets showed 130 KB of code and 11 % of the
// stall samples waiting for instructions)
template <int DP, int NC, int MINB, int LK>
__global__ void __launch_bounds__(NC, MINB) mh_sorted_kernel(const __grid_constant__ DevParams p,
                                                             const __grid_constant__ SortedTables<DP> tb, const int nc)
{
    static_assert(NC % 32 == 0, "whole warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = SortedSmem<DP, NC>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int d = p.d, W = p.W, T = p.T;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int idx = tid; idx < DP * DP; idx += NC) {
        const int i = idx / DP, j = idx % DP;
        S.Us[idx] = (i < d && j < d) ? p.U[i * d + j] : 0.0;
    }
    for (int k = tid; k < DP; k += NC) S.sS[k] = (k < d) ? p.sqrtS[k] : 0.0;
    const long long TW = (long long)T * W;
    const long long c0 = (long long)blockIdx.x * nc;       // first chain of this block
    const long long cme = c0 + tid;                        // the chain this thread owns in phase A
    const bool have = tid < nc && cme < TW;
    const int tme = have ? (int)(cme / W) : 0, wme = have ? (int)(cme % W) : 0;
    const double c_am = 2.4 / sqrt(2.0 * d);
    if (have) {
        const double *xg = p.x + (size_t)tme * d * W + wme;
        for (int k = 0; k < DP; ++k) S.xs[k * NC + tid] = (k < d) ? xg[(size_t)k * W] : 0.0;
        S.lnl[tid] = p.lnl[cme];
        S.lp[tid] = p.lp[cme];
        const double temp = p.mh_temp[tme], beta = 1.0 / temp;
        S.beta[tid] = beta;
        S.sct[tid] = (temp <= 100.0) ? sqrt(temp) : 1.0;   // ref :861-862, :919-920
        S.dsc[tid] = c_am * sqrt(1.0 / beta);               // ref :976
        S.ct[tid] = tme;
        S.cw[tid] = wme;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) S.cnt[j * NC + tid] = 0ull;
    if (tid < 12) (&S.count[0][0])[tid] = 0;
    // incremental forms of it % covUpdate, it % thin and it / thin (64-bit div/mod is ~100 instructions)
    // (positions of iteration it0 - 1; the loop advances them first)
    int am_slot = (int)((p.it0 - 1) % p.cov_update), thin_ctr = (int)((p.it0 - 1) % p.thin);
    long long row = (p.it0 - 1) / p.thin - p.rec_base;
    const int cov_update = (int)p.cov_update, thin = (int)p.thin;  // both < 2^31 (checked at create)
    const bool ring = p.temp_offset == 0 && p.am != nullptr;
    const unsigned long long bufsize = (unsigned long long)p.burn * (unsigned long long)W;
    const bool small_buf = bufsize <= 0xFFFFFFFFull;
    const uint32_t npairs = (uint32_t)((d + 1) >> 1);
    __syncthreads();

    const uint32_t gwme = (uint32_t)(p.walker_offset + wme), gtme = (uint32_t)(p.temp_offset + tme);
    draw_kind<NC>(p, p.it0, have, gwme, gtme, S.list[0], S.count[0]);
    __syncthreads();

    int lb = 0, cb = 0;  // list / count buffers of the current iteration
    for (long long it = p.it0; it <= p.it1; ++it) {
        // buffer / record positions of iteration it (ref :327-335)
        if (++am_slot == cov_update) am_slot = 0;
        if (++thin_ctr == thin) { thin_ctr = 0; ++row; }
        const bool keep = it < p.it1 || p.tail;  // the swap kernel does the last iteration's bookkeeping otherwise
        const int cb2 = (cb == 0) ? 2 : cb - 1;  // buffer of iteration it+2 == it-1: everyone has read it
        if (tid < 4) S.count[cb2][tid] = 0;
        const int nA = S.count[cb][0], nS = S.count[cb][1], nD = S.count[cb][2];
        const unsigned short *list = S.list[lb];
        // ---- phase B: thread r steps the r-th chain of (AM | SCAM | DE)
        if (tid < nA + nS + nD) {
            const int kindr = (tid < nA) ? 0 : (tid < nA + nS) ? 1 : 2;
            const int cl = (kindr == 0) ? list[tid] : (kindr == 1) ? list[NC + tid - nA] : list[2 * NC + tid - nA - nS];
            const int t = S.ct[cl], w = S.cw[cl];
            const double beta = S.beta[cl];
            // Words of this chain's stream: 0 = jump index, 1 = group index (both spent), then the proposal's draws
            // and the accept uniform.  The generator is inlined at four places only (code size): block 1 here,
            // block 2 for SCAM / DE, block 3 for DE, and the block loop of AM; rare paths call stream_word.
            const Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + w), (uint32_t)(p.temp_offset + t));
            const uint4 b1 = st.block(1);  // words 2, 3
            uint64_t wacc;                 // word of the accept uniform (ref :616)
            double q[DP];
            if (kindr == 0) {  // AM (ref :879-933): word 2 = prob, word 3 + m = normal pair m, then the accept uniform
                const double prob = word_to_unit(lo_word(b1));
                const double cd = c_am * (((prob > 0.97) ? 10.0 : (prob > 0.9) ? 0.2 : 1.0) * S.sct[cl]);
                // q = x + U delta accumulated column pair by column pair as the normals are drawn (same j order per
                // row as a row-wise dot product, so the same bits).  One trip per Philox block, the next block
                // generated ahead of the normals and FMAs of the current one.
#pragma unroll
                for (int i = 0; i < DP; ++i) q[i] = S.xs[i * NC + cl];
                const uint32_t ja = 3u + npairs, jb = ja >> 1;  // accept word and its block
                uint4 cur = b1;
#pragma unroll 1
                for (uint32_t b = 1; b <= jb; ++b) {
                    uint4 nxt = cur;
                    if (b < jb) nxt = st.block(b + 1);
                    // word 2b -> pair 2b - 3 (word 2 is prob), word 2b + 1 -> pair 2b - 2
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int m = 2 * (int)b - 3 + half;
                        if (m >= 0 && m < (int)npairs) {
                            double z0, z1;
                            word_to_normals(half ? cur.w : cur.y, half ? cur.z : cur.x, z0, z1);
                            const int j = 2 * m;
                            const double d0 = z0 * cd * S.sS[j], d1 = z1 * cd * S.sS[j + 1];  // sS is zero padded
                            const double *Uj = S.Us + j;
#pragma unroll
                            for (int i = 0; i < DP; ++i) {
                                q[i] = fma(Uj[i * DP], d0, q[i]);
                                q[i] = fma(Uj[i * DP + 1], d1, q[i]);
                            }
                        }
                    }
                    cur = nxt;
                }
                wacc = (ja & 1u) ? hi_word(cur) : lo_word(cur);
            } else {
                const uint4 b2 = st.block(2);  // words 4, 5
                if (kindr == 1) {  // SCAM (ref :820-876): prob, k, normal, accept
                    const double prob = word_to_unit(lo_word(b1));
                    const double scale = ((prob > 0.97) ? 10.0 : (prob > 0.9) ? 0.2 : 1.0) * S.sct[cl];
                    const int k = (int)word_to_int32(hi_word(b1), (uint32_t)d);
                    const double cd = 2.4 / sqrt(2.0) * scale;
                    double z0, z1;
                    word_to_normals(b2.y, b2.x, z0, z1);
                    const double coef = z0 * cd * S.sS[k];
#pragma unroll
                    for (int i = 0; i < DP; ++i) q[i] = fma(coef, S.Us[i * DP + k], S.xs[i * NC + cl]);
                    wacc = hi_word(b2);
                } else {  // DE (ref :936-985): mm, nn (redrawn while equal), prob, [scale uniform], accept
                    unsigned long long mm, nn;
                    if (small_buf) {
                        mm = word_to_int32(lo_word(b1), (uint32_t)bufsize);
                        nn = word_to_int32(hi_word(b1), (uint32_t)bufsize);
                    } else {
                        mm = word_to_int(lo_word(b1), bufsize);
                        nn = word_to_int(hi_word(b1), bufsize);
                    }
                    double scale = 1.0;
                    if (mm != nn) {
                        wacc = hi_word(b2);
                        if (!(word_to_unit(lo_word(b2)) > 0.5)) {  // ref :969-976
                            scale = word_to_unit(wacc) * S.dsc[cl];
                            wacc = lo_word(st.block(3));
                        }
                    } else {  // one draw in bufsize: the sequential form through the out-of-line generator
                        uint32_t j = 4;
                        while (mm == nn) nn = word_to_int(stream_word(st, j++), bufsize);
                        if (!(word_to_unit(stream_word(st, j++)) > 0.5)) scale = word_to_unit(stream_word(st, j++)) * S.dsc[cl];
                        wacc = stream_word(st, j);
                    }
                    const double *bm = p.de + de_row_offset(mm, bufsize, W, p.burn, p.de_head) * d;
                    const double *bn = p.de + de_row_offset(nn, bufsize, W, p.burn, p.de_head) * d;
                    if ((d & 1) == 0) {  // rows are 16-byte aligned: half as many load instructions
#pragma unroll
                        for (int i = 0; i < DP; i += 2) {
                            double2 vm = make_double2(0.0, 0.0), vn = vm;
                            if (i < d) {
                                vm = __ldg(reinterpret_cast<const double2 *>(bm + i));
                                vn = __ldg(reinterpret_cast<const double2 *>(bn + i));
                            }
                            q[i] = fma(scale, vm.x - vn.x, S.xs[i * NC + cl]);
                            if (i + 1 < DP) q[i + 1] = fma(scale, vm.y - vn.y, S.xs[(i + 1) * NC + cl]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < DP; ++i) {
                            const double sigma = (i < d) ? (__ldg(bm + i) - __ldg(bn + i)) : 0.0;
                            q[i] = fma(scale, sigma, S.xs[i * NC + cl]);
                        }
                    }
                }
            }
            bool inside = true;
#pragma unroll
            for (int k = 0; k < DP; ++k) inside = inside & (q[k] >= tb.lo[k]) & (q[k] <= tb.hi[k]);
            double lpn = inside ? p.p_inside : neg_inf();
            double lnln = 0.0, lnpn = neg_inf();
            if (inside) {
                lnln = logl_builtin<DP, LK>(p, tb, q);
                lnpn = beta * lnln + lpn;
            }
            const double lnl0 = S.lnl[cl], lp0 = S.lp[cl];
            const bool accept = hastings_accept(lnpn - (beta * lnl0 + lp0), wacc);
            const int jump = (kindr == 0) ? JUMP_AM : (kindr == 1) ? JUMP_SCAM : JUMP_DE;
            S.cnt[jump * NC + cl] += accept ? 0x100000001ull : 1ull;
            if (accept) {
#pragma unroll
                for (int k = 0; k < DP; ++k) S.xs[k * NC + cl] = q[k];
                S.lnl[cl] = lnln;
                S.lp[cl] = lpn;
            } else {
                lnln = lnl0;
                lpn = lp0;
            }
            // ---- updateChains of this iteration for the chain just stepped (ref :627, :321-335)
            if (p.trace && it - 1 < p.trace_cap)
                p.trace[((size_t)(it - 1) * T + t) * W + w] = (unsigned char)(jump | ((int)accept << 7));
            if (keep) {  // (from shared memory, after the accept store, in rolled loops: cold rung / thinned rows only)
                if (ring && t == 0) {
                    double *dst = p.am + (size_t)am_slot * d * W + w;
                    for (int k = 0; k < d; ++k) dst[(size_t)k * W] = S.xs[k * NC + cl];
                }
                if (t < p.ntr && thin_ctr == 0 && row >= 0 && row < p.rec_cap) {
                    const size_t r = ((size_t)row * p.ntr + t) * W + w;
                    double *dst = p.rec_x + r * d;
                    for (int k = 0; k < d; ++k) dst[k] = S.xs[k * NC + cl];
                    p.rec_lnl[r] = lnln;
                    p.rec_lnp[r] = beta * lnln + lpn;
                }
            }
        }
        // ---- phase A of the next iteration, in the shadow of the longer warps of phase B
        const int lb1 = lb ^ 1, cb1 = (cb == 2) ? 0 : cb + 1;
        if (it < p.it1) draw_kind<NC>(p, it + 1, have, gwme, gtme, S.list[lb1], S.count[cb1]);
        lb = lb1;
        cb = cb1;
        __syncthreads();
    }
    if (have) {
        double *xo = p.x + (size_t)tme * d * W + wme;
        for (int k = 0; k < d; ++k) xo[(size_t)k * W] = S.xs[k * NC + tid];
        p.lnl[cme] = S.lnl[tid];
        p.lp[cme] = S.lp[tid];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const unsigned long long c = S.cnt[j * NC + tid];
            p.prop[(size_t)j * TW + cme] += c & 0xFFFFFFFFull;
            p.acc[(size_t)j * TW + cme] += c >> 32;
        }
    }
}

}  // namespace ptm
