// Sorted MH kernel with the AM chains' random numbers produced one iteration ahead, in the idle shadow
// of the short warps ("shadow" variant of mh_sorted_kernel.cuh).
//
// In the sorted kernel the AM warps are the critical path of every block iteration: an AM step costs
// ~2300 instructions against ~1000 for SCAM / DE, half of the AM cost being random numbers (7 Philox
// blocks, d normals), and the SCAM / DE warps wait at the barrier meanwhile.  The draws are counter-based
// and independent of the chain state, so the jump kinds and per-kind lists of iteration it+1 are built
// during phase A of iteration it, and during phase B the threads that are NOT stepping an AM chain -- after
// finishing their own chain -- compute the normals, scale and log-uniform of the AM chains of iteration it+1
// into a double-buffered shared-memory queue.  An AM thread then only does the mat-vec, the prior, the
// likelihood and the Hastings test.  Two block barriers per iteration, as before; two blocks per SM are kept by
// bounding the queue to SHADOW_CAP AM chains per buffer (the chains beyond it draw inline, which only
// happens while the cycle has no DE jump yet or for unusual weights).  Identical results, draw for draw.
#pragma once
#include "mh_sorted_kernel.cuh"

namespace ptm {

constexpr int SHADOW_CAP = 128;

template <int DP, int NC>
struct ShadowSmem {
    double Us[DP * DP];
    double Ps[DP * DP];
    double sS[DP], mus[DP], los[DP], his[DP];
    double xs[DP * NC];  // [k][chain]
    double lnl[NC], lp[NC];
    double temp[NC], beta[NC];
    int ct[NC], cw[NC];
    unsigned cnt[6 * NC];
    unsigned short list[2][3 * NC];        // per-kind lists, by iteration parity
    unsigned char jt[NC];                  // jump id | accepted << 7 of the iteration just stepped (trace byte)
    int count[3][4];                       // list lengths (AM, SCAM, DE), by iteration mod 3
    float zf[2][DP * SHADOW_CAP];          // per AM list position < SHADOW_CAP: DP normals (exactly floats)
    double zs[2][2 * SHADOW_CAP];          // ... and cd, log(u)
    int taskctr[2];                        // next draw task of the iteration with that parity
};

template <int DP, int NC, int MINB>
__global__ void __launch_bounds__(NC, MINB) mh_shadow_kernel(const DevParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ShadowSmem<DP, NC> &S = *reinterpret_cast<ShadowSmem<DP, NC> *>(smem_raw);
    const int d = p.d, W = p.W, T = p.T;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int idx = tid; idx < DP * DP; idx += NC) {
        const int i = idx / DP, j = idx % DP;
        const bool in = (i < d && j < d);
        S.Us[idx] = in ? p.U[i * d + j] : 0.0;
        S.Ps[idx] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_P[i * d + j] : 0.0;
    }
    for (int k = tid; k < DP; k += NC) {
        const bool in = k < d;
        S.sS[k] = in ? p.sqrtS[k] : 0.0;
        S.mus[k] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_mu[k] : 0.0;
        S.los[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_lo[k] : neg_inf();
        S.his[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_hi[k] : pos_inf();
    }
    const long long TW = (long long)T * W;
    const long long c0 = (long long)blockIdx.x * NC;
    const long long cme = c0 + tid;
    const bool have = cme < TW;
    const int tme = have ? (int)(cme / W) : 0, wme = have ? (int)(cme % W) : 0;
    {
        const double *xg = p.x + (size_t)tme * d * W + wme;
#pragma unroll
        for (int k = 0; k < DP; ++k) S.xs[k * NC + tid] = (have && k < d) ? xg[(size_t)k * W] : 0.0;
        S.lnl[tid] = have ? p.lnl[cme] : 0.0;
        S.lp[tid] = have ? p.lp[cme] : 0.0;
        const double tp = have ? p.mh_temp[tme] : 1.0;
        S.temp[tid] = tp;
        S.beta[tid] = 1.0 / tp;
        S.ct[tid] = tme;
        S.cw[tid] = wme;
        S.jt[tid] = 0;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) S.cnt[j * NC + tid] = 0;
    if (tid < 12) (&S.count[0][0])[tid] = 0;
    if (tid < 2) S.taskctr[tid] = 0;
    const int inclusive = p.p_inclusive;
    long long am_slot = p.it0 % p.cov_update, thin_ctr = p.it0 % p.thin, row = p.it0 / p.thin - p.rec_base;
    const bool cold = have && tme == 0 && p.temp_offset == 0 && p.am != nullptr;
    const bool recorded = have && tme < p.ntr;
    const int npairs = (d + 1) >> 1, uword = 3 + npairs, am_tasks = ((uword + 2) >> 1) - 1;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    __syncthreads();

    // jump kind of the owned chain for iteration `itx` (ref :1058) appended to that iteration's lists
    auto build_lists = [&](long long itx) {
        int *count = S.count[itx % 3];
        unsigned short *list = S.list[itx & 1];
        int kind = 3;
        if (have) {
            Stream st(p.seed, PURPOSE_MH, (unsigned long long)itx, (uint32_t)(p.walker_offset + wme),
                      (uint32_t)(p.temp_offset + tme));
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
    };
    // the AM chains' draws after word 1 for iteration `itx`: tasks (chain, Philox block >= 1) taken from a
    // shared counter, so whoever is free first draws more (the result does not depend on who draws).
    // AM word order (ref :897-930): 2 = prob, 3 + j = normal pair j, 3 + npairs = accept u.
    auto draw_tasks = [&](long long itx) {
        const int nA = min(S.count[itx % 3][0], SHADOW_CAP), ntask = nA * am_tasks;
        const unsigned short *list = S.list[itx & 1];
        float *zf = S.zf[itx & 1];
        double *zs = S.zs[itx & 1];
        int *ctr = &S.taskctr[itx & 1];
        for (int r = atomicAdd(ctr, 1); r < ntask; r = atomicAdd(ctr, 1)) {
            const int a = r % nA, blk_i = 1 + r / nA;
            const int cl = list[a];
            const uint4 blk = philox4x32_10((uint32_t)itx, (PURPOSE_MH << 24) | (uint32_t)blk_i,
                                            (uint32_t)(p.walker_offset + S.cw[cl]), (uint32_t)(p.temp_offset + S.ct[cl]),
                                            k0, k1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wi = 2 * blk_i + h;
                const uint64_t word = h ? ((uint64_t)blk.z | ((uint64_t)blk.w << 32))
                                        : ((uint64_t)blk.x | ((uint64_t)blk.y << 32));
                if (wi == 2) {
                    zs[a] = 2.4 / sqrt(2.0 * d) * cov_jump_scale(word_to_unit(word), S.temp[cl]);
                } else if (wi < uword) {
                    double z0, z1;
                    word_to_normals(word, z0, z1);
                    zf[(2 * (wi - 3)) * SHADOW_CAP + a] = (float)z0;  // exact: the normals are single precision
                    if (2 * (wi - 3) + 1 < DP) zf[(2 * (wi - 3) + 1) * SHADOW_CAP + a] = (float)z1;
                } else if (wi == uword) {
                    zs[SHADOW_CAP + a] = log(word_to_unit(word));
                }
            }
        }
    };

    // prologue: lists and draws of the first iteration by everybody
    build_lists(p.it0);
    __syncthreads();
    draw_tasks(p.it0);
    __syncthreads();

    for (long long it = p.it0; it <= p.it1 + 1; ++it) {
        // ---- phase A: bookkeeping of iteration it-1 for the owned chain (ref :627); lists of iteration it+1
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
        if (it > p.it0) {
            if (++am_slot == p.cov_update) am_slot = 0;
            if (++thin_ctr == p.thin) { thin_ctr = 0; ++row; }
        }
        if (it > p.it1) break;
        const bool ahead = it + 1 <= p.it1;
        if (ahead) build_lists(it + 1);
        if (tid < 4) S.count[(it + 2) % 3][tid] = 0;  // slot of iteration it+2 == it-1: its readers are done
        if (tid == 0) S.taskctr[(it + 1) & 1] = 0;
        __syncthreads();
        // ---- phase B: thread r steps the r-th chain of (AM | SCAM | DE)
        const int *count = S.count[it % 3];
        const int nA = count[0], nS = count[1], nD = count[2];
        const unsigned short *list = S.list[it & 1];
        const float *zf = S.zf[it & 1];
        const double *zs = S.zs[it & 1];
        if (tid < nA + nS + nD) {
            const int kindr = (tid < nA) ? 0 : (tid < nA + nS) ? 1 : 2;
            const int cl = (kindr == 0) ? list[tid] : (kindr == 1) ? list[NC + tid - nA] : list[2 * NC + tid - nA - nS];
            const int t = S.ct[cl], w = S.cw[cl];
            const double temp = S.temp[cl], beta = S.beta[cl];
            Stream st(p.seed, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + w),
                      (uint32_t)(p.temp_offset + t));
            st.j = 2;  // words 0 (jump index) and 1 (group index of the single group) are spent
            double q[DP];
            double logu = 0.0;
            bool have_logu = false;
            if (kindr == 0) {  // AM (ref :879-933)
#pragma unroll
                for (int i = 0; i < DP; ++i) q[i] = S.xs[i * NC + cl];
                if (tid < SHADOW_CAP) {  // draws from the queue
                    const double cd = zs[tid];
                    logu = zs[SHADOW_CAP + tid];
                    have_logu = true;
#pragma unroll
                    for (int j = 0; j < DP; j += 2) {
                        const double d0 = (j < d) ? (double)zf[j * SHADOW_CAP + tid] * cd * S.sS[j] : 0.0;
                        const double d1 = (j + 1 < d) ? (double)zf[(j + 1) * SHADOW_CAP + tid] * cd * S.sS[j + 1] : 0.0;
#pragma unroll
                        for (int i = 0; i < DP; ++i) {
                            q[i] = fma(S.Us[i * DP + j], d0, q[i]);
                            if (j + 1 < DP) q[i] = fma(S.Us[i * DP + j + 1], d1, q[i]);
                        }
                    }
                } else {  // beyond the queue's capacity: draw inline
                    const double prob = word_to_unit(st.next());
                    const double cd = 2.4 / sqrt(2.0 * d) * cov_jump_scale(prob, temp);
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
#pragma unroll
                for (int i = 0; i < DP; ++i) {
                    const double sigma = (i < d) ? (__ldg(bm + i) - __ldg(bn + i)) : 0.0;
                    q[i] = fma(scale, sigma, S.xs[i * NC + cl]);
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
                        double rowv = 0.0;
#pragma unroll
                        for (int j = i; j < DP; ++j) rowv = fma(S.Ps[i * DP + j], q[j] - S.mus[j], rowv);
                        acc = fma(q[i] - S.mus[i], rowv, acc);
                    }
                    lnln = acc + p.g_offset;
                } else if (p.logl_kind == LOGL_CURVED) {
                    double tot = 0.0;
#pragma unroll
                    for (int bb = 0; bb + 1 < DP; bb += 2) {
                        if (bb + 1 < d) {
                            const double a = q[bb], y = q[bb + 1];
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
                            const double a = q[i + 1] - q[i] * q[i], bb = 1.0 - q[i];
                            tot -= 100.0 * a * a + bb * bb;
                        }
                    }
                    lnln = tot / 20.0;
                }
                lnpn = beta * lnln + lpn;
            }
            const double lnp0 = beta * S.lnl[cl] + S.lp[cl];
            const double diff = lnpn - lnp0;
            if (!have_logu) logu = log(word_to_unit(st.next()));
            const bool accept = diff > logu;
            const int jump = (kindr == 0) ? JUMP_AM : (kindr == 1) ? JUMP_SCAM : JUMP_DE;
            S.cnt[jump * NC + cl] += 1;
            unsigned char tb = (unsigned char)jump;
            if (accept) {
#pragma unroll
                for (int k = 0; k < DP; ++k) S.xs[k * NC + cl] = q[k];
                S.lnl[cl] = lnln;
                S.lp[cl] = lpn;
                S.cnt[(3 + jump) * NC + cl] += 1;
                tb |= 0x80;
            }
            S.jt[cl] = tb;
        }
        // ---- in the shadow of the AM warps: the draws of the AM chains of iteration it+1
        if (ahead) draw_tasks(it + 1);
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
