// Warp-specialised variant of the sorted MH kernel (mh_sorted_kernel.cuh): the block is split into
// DRAW warps and STEP warps that run one iteration apart.
//
// In the sorted kernel an AM chain's thread spends three quarters of its instructions on random numbers
// (d normals by Box-Muller), so the AM warps are ~3.5x longer than SCAM / DE warps and the block waits for
// them at every barrier with most schedulers idle.  The draws are counter-based and do not depend on the
// chain state, so they can be produced ahead of time: NPW draw warps compute, for iteration it+1, every
// chain's jump kind, the per-kind lists and the AM chains' normals / scale / log-uniform into a
// double-buffered shared-memory queue while the 8 step warps (thread r = r-th chain of AM | SCAM | DE, as
// in the sorted kernel) run proposal, prior, likelihood and the Hastings test of iteration it.  The two
// groups hand buffers over with named barriers (bar.arrive / bar.sync): FULL[b] draw -> step, EMPTY[b]
// step -> draw; only the step warps meet at a (256-thread) barrier of their own each iteration.
// Results are identical to the sorted kernel draw for draw.
#pragma once
#include "mh_sorted_kernel.cuh"

namespace ptm {

constexpr int PIPE_NC = 256;  // step threads = chains per block

template <int DP>
struct PipeSmem {
    double Us[DP * DP];
    double Ps[DP * DP];
    double sS[DP], mus[DP], los[DP], his[DP];
    double xs[DP * PIPE_NC];  // [k][chain]
    double lnl[PIPE_NC], lp[PIPE_NC];
    double temp[PIPE_NC], beta[PIPE_NC];
    int ct[PIPE_NC], cw[PIPE_NC];
    unsigned cnt[6 * PIPE_NC];
    unsigned char jt[PIPE_NC];  // jump id | accepted << 7 of the iteration just stepped (trace byte)
    // queue filled by the draw warps, double buffered by iteration parity
    double zb[2][(DP + 2) * PIPE_NC];        // per AM list position: DP normals, cd, log(u)
    unsigned short list[2][3 * PIPE_NC];     // chains of each kind (AM | SCAM | DE)
    unsigned char kind[2][PIPE_NC];          // jump id of every chain
    int count[2][4];
};

enum : int { BAR_DRAW = 1, BAR_STEP = 2, BAR_FULL = 3, BAR_EMPTY = 5 };  // FULL, FULL+1, EMPTY, EMPTY+1

__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int DP, int NPW>
__global__ void __launch_bounds__(PIPE_NC + 32 * NPW, 1) mh_pipe_kernel(const DevParams p)
{
    constexpr int NC = PIPE_NC, NPT = 32 * NPW, NTH = NC + NPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PipeSmem<DP> &S = *reinterpret_cast<PipeSmem<DP> *>(smem_raw);
    const int d = p.d, W = p.W, T = p.T;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int idx = tid; idx < DP * DP; idx += NTH) {
        const int i = idx / DP, j = idx % DP;
        const bool in = (i < d && j < d);
        S.Us[idx] = in ? p.U[i * d + j] : 0.0;
        S.Ps[idx] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_P[i * d + j] : 0.0;
    }
    for (int k = tid; k < DP; k += NTH) {
        const bool in = k < d;
        S.sS[k] = in ? p.sqrtS[k] : 0.0;
        S.mus[k] = (in && p.logl_kind == LOGL_GAUSSIAN) ? p.g_mu[k] : 0.0;
        S.los[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_lo[k] : neg_inf();
        S.his[k] = (in && p.logp_kind == LOGP_UNIFORM) ? p.p_hi[k] : pos_inf();
    }
    const long long TW = (long long)T * W;
    const long long c0 = (long long)blockIdx.x * NC;
    if (tid < NC) {
        const long long cme = c0 + tid;
        const bool have = cme < TW;
        const int tme = have ? (int)(cme / W) : 0, wme = have ? (int)(cme % W) : 0;
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
#pragma unroll
        for (int j = 0; j < 6; ++j) S.cnt[j * NC + tid] = 0;
    }
    __syncthreads();

    if (tid >= NC) {
        // =========================== draw warps: iteration `it` into buffer it & 1 ===========================
        const int ptid = tid - NC;
        const int npairs = (d + 1) >> 1, uword = 3 + npairs, am_tasks = ((uword + 2) >> 1) - 1;
        const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
        for (long long it = p.it0; it <= p.it1; ++it) {
            const int b = (int)(it & 1);
            if (it >= p.it0 + 2) bar_sync(BAR_EMPTY + b, NTH);  // the step warps are done with this buffer
            if (ptid < 4) S.count[b][ptid] = 0;
            bar_sync(BAR_DRAW, NPT);
            // jump kind of every chain (ref :1058) and the per-kind lists
            for (int base = 0; base < NC; base += NPT) {
                const int c = base + ptid;
                int kind = 3;
                if (c < NC && c0 + c < TW) {
                    Stream st(p.seed, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + S.cw[c]),
                              (uint32_t)(p.temp_offset + S.ct[c]));
                    const int jump = pick_jump(p, st);
                    S.kind[b][c] = (unsigned char)jump;
                    kind = (jump == JUMP_AM) ? 0 : (jump == JUMP_SCAM) ? 1 : 2;
                }
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const unsigned m = __ballot_sync(0xffffffffu, kind == kk);
                    if (m) {
                        int at = 0;
                        const int leader = __ffs(m) - 1;
                        if (lane == leader) at = atomicAdd(&S.count[b][kk], __popc(m));
                        at = __shfl_sync(0xffffffffu, at, leader);
                        if (kind == kk) S.list[b][kk * NC + at + __popc(m & ((1u << lane) - 1u))] = (unsigned short)c;
                    }
                }
            }
            bar_sync(BAR_DRAW, NPT);
            // the AM chains' draws after word 1, one task per (chain, Philox block >= 1).
            // AM word order (ref :897-930): 2 = prob, 3 + j = normal pair j, 3 + npairs = accept u.
            const int nA = S.count[b][0], ntask = nA * am_tasks;
            double *zb = S.zb[b];
            for (int r = ptid; r < ntask; r += NPT) {
                const int a = r % nA, blk_i = 1 + r / nA;
                const int cl = S.list[b][a];
                const uint4 blk = philox4x32_10((uint32_t)it, (PURPOSE_MH << 24) | (uint32_t)blk_i,
                                                (uint32_t)(p.walker_offset + S.cw[cl]),
                                                (uint32_t)(p.temp_offset + S.ct[cl]), k0, k1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int wi = 2 * blk_i + h;
                    const uint64_t word = h ? ((uint64_t)blk.z | ((uint64_t)blk.w << 32))
                                            : ((uint64_t)blk.x | ((uint64_t)blk.y << 32));
                    if (wi == 2) {
                        zb[DP * NC + a] = 2.4 / sqrt(2.0 * d) * cov_jump_scale(word_to_unit(word), S.temp[cl]);
                    } else if (wi < uword) {
                        double z0, z1;
                        word_to_normals(word, z0, z1);
                        zb[(2 * (wi - 3)) * NC + a] = z0;
                        if (2 * (wi - 3) + 1 < DP) zb[(2 * (wi - 3) + 1) * NC + a] = z1;
                    } else if (wi == uword) {
                        zb[(DP + 1) * NC + a] = log(word_to_unit(word));
                    }
                }
            }
            __threadfence_block();
            bar_arrive(BAR_FULL + b, NTH);
        }
        return;
    }

    // =============================== step warps: thread tid owns chain tid ===============================
    const long long cme = c0 + tid;
    const bool have = cme < TW;
    const int tme = S.ct[tid], wme = S.cw[tid];
    const int inclusive = p.p_inclusive;
    long long am_slot = p.it0 % p.cov_update, thin_ctr = p.it0 % p.thin, row = p.it0 / p.thin - p.rec_base;
    const bool cold = have && tme == 0 && p.temp_offset == 0 && p.am != nullptr;
    const bool recorded = have && tme < p.ntr;
    for (long long it = p.it0; it <= p.it1 + 1; ++it) {
        // ---- bookkeeping of iteration it-1 for the owned chain (ref :627)
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
        const int b = (int)(it & 1);
        bar_sync(BAR_FULL + b, NTH);  // the draw warps have filled this buffer
        const int nA = S.count[b][0], nS = S.count[b][1], nD = S.count[b][2];
        const unsigned short *list = S.list[b];
        const double *zb = S.zb[b];
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
            if (kindr == 0) {  // AM (ref :879-933), draws from the queue
                const double cd = zb[DP * NC + tid];
                logu = zb[(DP + 1) * NC + tid];
#pragma unroll
                for (int i = 0; i < DP; ++i) q[i] = S.xs[i * NC + cl];
#pragma unroll
                for (int j = 0; j < DP; j += 2) {
                    const double d0 = (j < d) ? zb[j * NC + tid] * cd * S.sS[j] : 0.0;
                    const double d1 = (j + 1 < d) ? zb[(j + 1) * NC + tid] * cd * S.sS[j + 1] : 0.0;
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
            if (kindr != 0) logu = log(word_to_unit(st.next()));
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
        if (it + 2 <= p.it1) bar_arrive(BAR_EMPTY + b, NTH);  // this buffer may be refilled (iteration it+2)
        bar_sync(BAR_STEP, NC);
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
