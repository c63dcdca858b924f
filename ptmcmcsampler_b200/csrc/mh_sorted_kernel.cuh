// Divergence-free fused MH kernel (the production path for ndim <= 32, identity parameter group).
//
// The three built-in proposals cost very different amounts (AM: d normals + a d x d mat-vec, SCAM and
// DE: O(d)), and every chain picks its proposal independently each iteration (ref _jump :1048-1067),
// so a thread-per-chain warp executes all three paths every step.  Because the draws are
// counter-based, a chain's jump kind for an iteration is known before any state is touched.  Each
// iteration the block therefore
//   A. lets thread i draw the jump kind of "its" chain i and append the chain to that kind's list
//      (warp-aggregated shared-memory atomics), after doing the previous iteration's
//      buffer/record bookkeeping for chain i with coalesced global stores;
//   B. lets thread r process the r-th chain of the concatenated lists (AM | SCAM | DE), so that
//      warps are proposal-uniform except at the two list boundaries.
// Chain state (x, lnL, lnprior, counters) lives in shared memory for the whole launch; two block
// barriers per iteration.  Results are identical to the thread-per-chain kernel draw for draw.
#pragma once
#include "mh_kernels.cuh"

namespace ptm {

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

template <int DP, int NC>
struct SortedSmem {
    double Us[DP * DP];
    double Ps[DP * DP];
    double sS[DP], mus[DP], los[DP], his[DP];
    double xs[DP * NC];      // [k][chain]
    double lnl[NC], lp[NC];
    double temp[NC], beta[NC];
    int ct[NC], cw[NC];      // rung and walker of each chain of the block
    unsigned cnt[6 * NC];    // [prop scam, am, de, acc scam, am, de][chain]
    unsigned short list[3 * NC];
    unsigned char jt[NC];    // jump id | accepted << 7 of the current iteration
    int count[8];            // list lengths (AM, SCAM, DE), double-buffered by iteration parity
};

template <int DP, int NC, int MINB>
__global__ void __launch_bounds__(NC, MINB) mh_sorted_kernel(const DevParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortedSmem<DP, NC> &S = *reinterpret_cast<SortedSmem<DP, NC> *>(smem_raw);
    const int d = p.d, W = p.W, T = p.T;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nc = blockDim.x;  // chains of this block (<= NC, the capacity the arrays are laid out for)
    for (int idx = tid; idx < DP * DP; idx += nc) {
        const int i = idx / DP, j = idx % DP;
        const bool in = (i < d && j < d);
        S.Us[idx] = in ? p.U[i * d + j] : 0.0;
        S.Ps[idx] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_P[i * d + j] : 0.0;
    }
    for (int k = tid; k < DP; k += nc) {
        const bool in = k < d;
        S.sS[k] = in ? p.sqrtS[k] : 0.0;
        S.mus[k] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_mu[k] : 0.0;
        S.los[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_lo[k] : neg_inf();
        S.his[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_hi[k] : pos_inf();
    }
    const long long TW = (long long)T * W;
    const long long c0 = (long long)blockIdx.x * nc;       // first chain of this block
    const long long cme = c0 + tid;                        // the chain this thread owns in phase A
    const bool have = cme < TW;
    const int tme = have ? (int)(cme / W) : 0, wme = have ? (int)(cme % W) : 0;
    if (have) {
        const double *xg = p.x + (size_t)tme * d * W + wme;
#pragma unroll
        for (int k = 0; k < DP; ++k) S.xs[k * NC + tid] = (k < d) ? xg[(size_t)k * W] : 0.0;
        S.lnl[tid] = p.lnl[cme];
        S.lp[tid] = p.lp[cme];
        S.temp[tid] = p.mh_temp[tme];
        S.beta[tid] = 1.0 / p.mh_temp[tme];
        S.ct[tid] = tme;
        S.cw[tid] = wme;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) S.cnt[j * NC + tid] = 0;
    if (tid < 8) S.count[tid] = 0;
    const int inclusive = p.p_inclusive;
    // incremental forms of it % covUpdate, it % thin and it / thin (64-bit div/mod is ~100 instructions)
    long long am_slot = p.it0 % p.cov_update, thin_ctr = p.it0 % p.thin, row = p.it0 / p.thin - p.rec_base;
    const bool cold = have && tme == 0 && p.temp_offset == 0 && p.am != nullptr;
    const bool recorded = have && tme < p.ntr;
    __syncthreads();

    for (long long it = p.it0; it <= p.it1 + 1; ++it) {
        // ---- phase A: bookkeeping of iteration it-1 for the owned chain (ref :627), then the
        //      jump kind of iteration it (ref :1058) and the per-kind lists
        if (have && it > p.it0) {
            const long long ib = it - 1;
            if (p.trace && ib - 1 < p.trace_cap) p.trace[((size_t)(ib - 1) * T + tme) * W + wme] = S.jt[tid];
            if (ib < p.it1 || p.tail) {
                if (cold) {
                    double *dst = p.am + (size_t)am_slot * d * W + wme;
#pragma unroll
                    for (int k = 0; k < DP; ++k)
                        if (k < d) dst[(size_t)k * W] = S.xs[k * NC + tid];
                }
                if (recorded && thin_ctr == 0 && row >= 0 && row < p.rec_cap) {
                    const size_t r = ((size_t)row * p.ntr + tme) * W + wme;
                    double *dst = p.rec_x + r * d;
#pragma unroll
                    for (int k = 0; k < DP; ++k)
                        if (k < d) dst[k] = S.xs[k * NC + tid];
                    p.rec_lnl[r] = S.lnl[tid];
                    p.rec_lnp[r] = S.beta[tid] * S.lnl[tid] + S.lp[tid];
                }
            }
        }
        if (it > p.it0) {  // advance the ring slot / thinning counters from iteration it-1 to it
            if (++am_slot == p.cov_update) am_slot = 0;
            if (++thin_ctr == p.thin) { thin_ctr = 0; ++row; }
        }
        if (it > p.it1) break;
        int *count = S.count + 4 * (int)(it & 1);
        int kind = 3;  // none
        if (have) {
            Stream st(p.seed, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + wme),
                      (uint32_t)(p.temp_offset + tme));
            const int jump = pick_jump(p, st);
            S.jt[tid] = (unsigned char)jump;
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
                if (kind == kk) S.list[kk * NC + base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)tid;
            }
        }
        __syncthreads();
        // ---- phase B: thread r processes the r-th chain of (AM | SCAM | DE)
        const int nA = count[0], nS = count[1], nD = count[2];
        if (tid < 4) S.count[4 * (int)((it + 1) & 1) + tid] = 0;  // next iteration's counters
        if (tid < nA + nS + nD) {
            const int kindr = (tid < nA) ? 0 : (tid < nA + nS) ? 1 : 2;
            const int cl = (kindr == 0) ? S.list[tid] : (kindr == 1) ? S.list[NC + tid - nA] : S.list[2 * NC + tid - nA - nS];
            const int t = S.ct[cl], w = S.cw[cl];
            const double temp = S.temp[cl], beta = S.beta[cl];
            Stream st(p.seed, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + w),
                      (uint32_t)(p.temp_offset + t));
            st.j = 2;  // words 0 (jump index) and 1 (group index of the single group) are spent
            double q[DP];
            if (kindr == 0) {  // AM (ref :879-933)
                const double prob = word_to_unit(st.next());
                const double cd = 2.4 / sqrt(2.0 * d) * cov_jump_scale(prob, temp);
                // q = x + U delta accumulated column pair by column pair as the normals are drawn (same
                // j order per row as a row-wise dot product, so the same bits): no delta array stays live
                // and the draw of pair j+1 overlaps the 2 d FMAs of pair j
#pragma unroll
                for (int i = 0; i < DP; ++i) q[i] = S.xs[i * NC + cl];
#pragma unroll
                for (int j = 0; j < DP; j += 2) {
                    double z0 = 0.0, z1 = 0.0;
                    if (j < d) word_to_normals(st.next(), z0, z1);
                    const double d0 = z0 * cd * S.sS[j];
                    const double d1 = (j + 1 < DP) ? z1 * cd * S.sS[j + 1] : 0.0;
#pragma unroll
                    for (int i = 0; i < DP; ++i) {
                        q[i] = fma(S.Us[i * DP + j], d0, q[i]);
                        if (j + 1 < DP) q[i] = fma(S.Us[i * DP + j + 1], d1, q[i]);
                    }
                }
            } else if (kindr == 1) {  // SCAM (ref :820-876)
                const double prob = word_to_unit(st.next());
                const double scale = cov_jump_scale(prob, temp);
                const int k = (int)word_to_int(st.next(), (unsigned long long)d);
                const double cd = 2.4 / sqrt(2.0) * scale;
                double z0, z1;
                word_to_normals(st.next(), z0, z1);
                const double coef = z0 * cd * S.sS[k];
#pragma unroll
                for (int i = 0; i < DP; ++i) q[i] = fma(coef, S.Us[i * DP + k], S.xs[i * NC + cl]);
            } else {  // DE (ref :936-985)
                const unsigned long long bufsize = (unsigned long long)p.burn * (unsigned long long)W;
                const unsigned long long mm = word_to_int(st.next(), bufsize);
                unsigned long long nn = word_to_int(st.next(), bufsize);
                while (mm == nn) nn = word_to_int(st.next(), bufsize);
                const double prob = word_to_unit(st.next());
                double scale = 1.0;
                if (!(prob > 0.5)) scale = word_to_unit(st.next()) * 2.4 / sqrt(2.0 * d) * sqrt(1.0 / beta);
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
            bool inside = true;
#pragma unroll
            for (int k = 0; k < DP; ++k) inside = inside && in_box(q[k], S.los[k], S.his[k], inclusive);
            const double lpn = inside ? p.p_inside : neg_inf();
            double lnln = 0.0, lnpn = neg_inf();
            if (inside) {
                if (p.logl_kind == LOGL_GAUSSIAN) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < DP; ++i) {
                        double row = 0.0;
#pragma unroll
                        for (int j = i; j < DP; ++j) row = fma(S.Ps[i * DP + j], q[j] - S.mus[j], row);
                        acc = fma(q[i] - S.mus[i], row, acc);
                    }
                    lnln = acc + p.g_offset;
                } else if (p.logl_kind == LOGL_CURVED) {
                    double tot = 0.0;
#pragma unroll
                    for (int b = 0; b + 1 < DP; b += 2) {
                        if (b + 1 < d) {
                            const double a = q[b], y = q[b + 1];
                            const double t0 = 9.0 + 4.0 * a * a + 9.0 * y;
                            tot += log(exp(-a * a - t0 * t0) + 0.5 * exp(-8.0 * a * a - 8.0 * (y - 2.0) * (y - 2.0)));
                        }
                    }
                    lnln = tot;
                } else {
                    double tot = 0.0;
#pragma unroll
                    for (int i = 0; i + 1 < DP; ++i) {
                        if (i + 1 < d) {
                            const double a = q[i + 1] - q[i] * q[i], b = 1.0 - q[i];
                            tot -= 100.0 * a * a + b * b;
                        }
                    }
                    lnln = tot / 20.0;
                }
                lnpn = beta * lnln + lpn;
            }
            const double lnp0 = beta * S.lnl[cl] + S.lp[cl];
            const double diff = lnpn - lnp0;
            const double u = word_to_unit(st.next());
            const bool accept = diff > log(u);
            const int jump = (kindr == 0) ? JUMP_AM : (kindr == 1) ? JUMP_SCAM : JUMP_DE;
            S.cnt[jump * NC + cl] += 1;
            if (accept) {
#pragma unroll
                for (int k = 0; k < DP; ++k) S.xs[k * NC + cl] = q[k];
                S.lnl[cl] = lnln;
                S.lp[cl] = lpn;
                S.cnt[(3 + jump) * NC + cl] += 1;
                S.jt[cl] = (unsigned char)(jump | 0x80);
            }
        }
        __syncthreads();
    }
    if (have) {
        double *xo = p.x + (size_t)tme * d * W + wme;
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < d) xo[(size_t)k * W] = S.xs[k * NC + tid];
        p.lnl[cme] = S.lnl[tid];
        p.lp[cme] = S.lp[tid];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            p.prop[(size_t)j * TW + cme] += S.cnt[j * NC + tid];
            p.acc[(size_t)j * TW + cme] += S.cnt[(3 + j) * NC + tid];
        }
    }
}

}  // namespace ptm
