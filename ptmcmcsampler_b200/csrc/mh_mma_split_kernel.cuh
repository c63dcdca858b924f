// Tensor-core (fp64 DMMA) fused MH kernel for ndim in (32, 128]: the "split" variant of mh_mma_kernel.cuh (same fragment
// layout, same draws, same arithmetic per scalar; read that header first).
//
// What the measurements of the one-block-per-SM kernel said (scripts/mma_clocks.py, scripts/micro/dmma_lat.cu): a DMMA
// issues every 16 clocks per SM sub-partition and two dependent chains in ONE warp saturate the pipe, but the tensor pipe
// was busy 36 % of the time: every phase between two block barriers was a latency chain run by few warps (the jump pick
// and the scalar draws of SCAM / DE steps by the chains' own threads: one or two warps walking both branches in turn;
// the factor's fragments fetched from L2 one k-tile ahead of their use).  Hence:
//   * two blocks of <= 32 chains per SM at <= 128 registers: one block's draw phases run under the other's tensor phases.
//     Shared memory per block: the chains' rows and the form's fragment image (packed lower triangle of the Cholesky
//     factor when the form is positive definite); the factor's image is read through L1 / L2;
//   * everything random about a step except the AM normals is state-free and small, so it is drawn AHEAD by warps that are
//     idle while the others run the Hastings tests of iteration it: the jump kinds and per-kind lists of iteration it+2
//     (warp 7), and for iteration it+1 the DE chains' history rows / scales / accept words, the rows prefetched into L2 a
//     whole iteration before the gather (warp 6), the SCAM chains' component / coefficient / accept word (warp 5) and the
//     AM chains' jump scale (warp 4), each branch-free with its Philox blocks generated side by side.  Lists and jump ids
//     are triple-buffered (iteration mod 3), these scalars double-buffered by the parity of the iteration;
//   * phase R is then evenly spread tasks only: gathers (loads issued first) and (AM chain, two Philox blocks) tasks that
//     store the normals already scaled, delta = z cd sqrt(S);
//   * phase P: every warp takes <= 2 n-tiles of ALL AM tiles, so each fragment of the factor is fetched once per block
//     and feeds up to 8 DMMAs; fragments are prefetched PD k-tiles ahead;
//   * phase L: two warps per 8-chain tile, each a balanced part of the (triangular) quadratic form.
//
// Per iteration, per block (5 block barriers):
//   R   warp 0: buffers / record of iteration it-1; all: gather tasks (DE or SCAM chain, 8 columns) -> the step into the
//       chain's zq row, and AM tasks -> delta into zq
//   P   zq <- U delta for the AM chains by DMMA (results written after a barrier)
//   L   per (tile, half): proposal, box test, part of the quadratic form by DMMA; barrier; per tile: Hastings test, state
//       update, while warps 4..7 make the draws of the next iterations
#pragma once
#include "mh_mma_kernel.cuh"

namespace ptm {

// n-tile at which the quadratic form of a tile is split over two warps: the triangular form costs NT - nt tile products
// for n-tile nt, the symmetric form NT for every n-tile
__host__ __device__ constexpr int mma_split_at(int NT, bool tri)
{
    if (!tri) return (NT + 1) / 2;
    int s = 0, acc = 0;
    while (2 * acc < NT * (NT + 1) / 2) acc += NT - s++;
    return s;  // NT = 8: 3 (21 | 15), 13: 4 (46 | 45), 16: 5 (70 | 66)
}

// n-tiles [N0, N1) of one tile's quadratic form -- y = dv . Pf by DMMA, then sum y^2 (TRI) or sum y dv -- for this lane's
// columns of chain row (xrow, zrow): proposal v = x + mult * z, dv = v - mu.  BOX: also the box test over every column.
// pf points at this lane's element of tile 0 (packed lower-triangular tile order when TRI).
template <int NT, int N0, int N1, bool TRI, bool BOX>
__device__ __forceinline__ double quad_part(const double *xrow, const double *zrow, double mult, const double *mus,
                                            const double *los, const double *his, int inclusive, const double2 *pf, int t,
                                            bool &inside)
{
    constexpr int K0 = (TRI && !BOX) ? N0 : 0;  // first k-tile whose columns this part needs
    double dv[NT][2];
#pragma unroll
    for (int nt = K0; nt < NT; ++nt) {
        const int col = 8 * nt + 2 * t;
        const double2 x = *reinterpret_cast<const double2 *>(xrow + col);
        const double2 z = *reinterpret_cast<const double2 *>(zrow + col);
        const double2 m = *reinterpret_cast<const double2 *>(mus + col);
        const double vx = __dadd_rn(x.x, __dmul_rn(mult, z.x)), vy = __dadd_rn(x.y, __dmul_rn(mult, z.y));
        dv[nt][0] = vx - m.x;
        dv[nt][1] = vy - m.y;
        if (BOX) {
            const double2 lo = *reinterpret_cast<const double2 *>(los + col);
            const double2 hi = *reinterpret_cast<const double2 *>(his + col);
            inside = inside && in_box(vx, lo.x, hi.x, inclusive) && in_box(vy, lo.y, hi.y, inclusive);
        }
    }
    double part = 0.0;
    constexpr int NB = 2;  // independent accumulators in flight: two chains per warp saturate the DMMA pipe
#pragma unroll
    for (int nb = N0; nb < N1; nb += NB) {
        double y[NB][2];
#pragma unroll
        for (int j = 0; j < NB; ++j) y[j][0] = y[j][1] = 0.0;
#pragma unroll
        for (int kk = (TRI ? nb : 0); kk < NT; ++kk) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (nb + j < N1 && (!TRI || kk >= nb + j)) {
                    const double2 b = pf[(TRI ? kk * (kk + 1) / 2 + nb + j : kk * NT + nb + j) * 32];
                    dmma884(y[j][0], y[j][1], dv[kk][0], b.x);
                    dmma884(y[j][0], y[j][1], dv[kk][1], b.y);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            if (nb + j < N1) {
                part = fma(y[j][0], TRI ? y[j][0] : dv[nb + j][0], part);
                part = fma(y[j][1], TRI ? y[j][1] : dv[nb + j][1], part);
            }
        }
    }
    return part;
}

template <int NT>
__global__ void __launch_bounds__(MMA_THREADS, 2) mh_mma_split_kernel(const __grid_constant__ DevParams p,
                                                                      const __grid_constant__ MmaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KP = 8 * NT;
    const int d = p.d, W = p.W, T = p.T;
    const int nc = a.nc, ld = a.ld;  // nc <= 32: one warp owns the chains
    const bool tri = a.tri != 0;
    const MmaLayout &L = a.L;
    double *xs = reinterpret_cast<double *>(smem_raw + L.xs);
    double *zq = reinterpret_cast<double *>(smem_raw + L.zq);
    double *Pf = reinterpret_cast<double *>(smem_raw + L.pf);
    double *sSs = reinterpret_cast<double *>(smem_raw + L.ss);
    double *mus = reinterpret_cast<double *>(smem_raw + L.mu);
    double *los = reinterpret_cast<double *>(smem_raw + L.lo);
    double *his = reinterpret_cast<double *>(smem_raw + L.hi);
    double *s_lnl = reinterpret_cast<double *>(smem_raw + L.lnl);
    double *s_lp = reinterpret_cast<double *>(smem_raw + L.lp);
    double *s_temp = reinterpret_cast<double *>(smem_raw + L.temp);
    double *s_beta = reinterpret_cast<double *>(smem_raw + L.beta);
    // per-chain scalars of a step, [2][nc] by the parity of the iteration
    double *s_sca2 = reinterpret_cast<double *>(smem_raw + L.sca);    // AM: cd; SCAM: coefficient; DE: scale
    unsigned long long *s_rowm2 = reinterpret_cast<unsigned long long *>(smem_raw + L.rowm);  // DE row offsets;
    unsigned long long *s_rown2 = reinterpret_cast<unsigned long long *>(smem_raw + L.rown);  // SCAM: rowm = k d
    unsigned long long *s_uword2 = reinterpret_cast<unsigned long long *>(smem_raw + L.logu);  // accept-uniform word
    double *s_part = reinterpret_cast<double *>(smem_raw + L.part);
    int *s_ct = reinterpret_cast<int *>(smem_raw + L.ct);
    int *s_cw = reinterpret_cast<int *>(smem_raw + L.cw);
    unsigned *s_cnt = reinterpret_cast<unsigned *>(smem_raw + L.cnt);
    unsigned short *s_list2 = reinterpret_cast<unsigned short *>(smem_raw + L.list);  // [3][3][nc], slot = iteration mod 3
    unsigned char *s_jt2 = smem_raw + L.jt;                                             // [3][nc]
    int *s_count = reinterpret_cast<int *>(smem_raw + L.count);                         // [3][4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, t = lane & 3;

    // ---- stage tables and state
    for (int idx = tid; idx < a.pf_tiles * 64; idx += MMA_THREADS) Pf[idx] = a.Pf[idx];
    const bool gauss = p.logl_kind == LOGL_GAUSSIAN, box = p.logp_kind == LOGP_UNIFORM;
    for (int k = tid; k < KP; k += MMA_THREADS) {
        const bool in = k < d;
        sSs[k] = in ? p.sqrtS[k] : 0.0;
        mus[k] = (in && gauss) ? p.g_mu[k] : 0.0;
        los[k] = (in && box) ? p.p_lo[k] : neg_inf();
        his[k] = (in && box) ? p.p_hi[k] : pos_inf();
    }
    const long long TW = (long long)T * W;
    const long long c0 = (long long)blockIdx.x * nc;
    const long long cme = c0 + tid;
    const bool have = tid < nc && cme < TW;
    const int tme = have ? (int)(cme / W) : 0, wme = have ? (int)(cme % W) : 0;
    if (tid < nc) {
        for (int k = 0; k < ld; ++k) {
            xs[tid * ld + k] = (have && k < d) ? p.x[((size_t)tme * d + k) * W + wme] : 0.0;
            zq[tid * ld + k] = 0.0;
        }
        s_lnl[tid] = have ? p.lnl[cme] : 0.0;
        s_lp[tid] = have ? p.lp[cme] : 0.0;
        const double tp = have ? p.mh_temp[tme] : 1.0;
        s_temp[tid] = tp;
        s_beta[tid] = 1.0 / tp;
        s_ct[tid] = tme;
        s_cw[tid] = wme;
        s_jt2[tid] = s_jt2[nc + tid] = s_jt2[2 * nc + tid] = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) s_cnt[j * nc + tid] = 0;
    }
    const int inclusive = p.p_inclusive;
    const int npairs = (d + 1) >> 1, uword = 3 + npairs, am_tasks = ((uword + 2) >> 1) - 1;
    const unsigned long long bufsize = (unsigned long long)p.burn * (unsigned long long)W;
    const int ntiles = nc >> 3;

    // jump kind of iteration `it` (ref :1058) for the chains of the block, by the lanes of ONE warp, and the per-kind lists
    auto pick_kinds = [&](long long it, int buf) {  // buf: list slot = it mod 3
        const bool mine = lane < nc && c0 + lane < TW;
        int kind = 3;
        if (mine) {
            Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + s_cw[lane]),
                      (uint32_t)(p.temp_offset + s_ct[lane]));
            const int jump = pick_jump(p, st);
            s_jt2[buf * nc + lane] = (unsigned char)jump;
            kind = (jump == JUMP_AM) ? 0 : (jump == JUMP_SCAM) ? 1 : 2;
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const unsigned m = __ballot_sync(0xffffffffu, kind == kk);
            if (kind == kk) s_list2[(buf * 3 + kk) * nc + __popc(m & ((1u << lane) - 1u))] = (unsigned short)lane;
            if (lane == 0) s_count[4 * buf + kk] = __popc(m);
        }
    };
    // DE chains of iteration `it` (ref :955-976), one per lane: the two history rows (words 2, 3, redrawn while equal), prob,
    // scale, accept uniform; the rows are prefetched into L2 for the gather tasks of phase R
    auto draw_de = [&](long long it, int slot) {  // slot: list slot = it mod 3; the scalars go by the parity of it
        const int buf = (int)(it & 1), nD = s_count[4 * slot + 2];
        for (int i = lane; i < nD; i += 32) {
            const int cl = s_list2[(slot * 3 + 2) * nc + i];
            const Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + s_cw[cl]),
                            (uint32_t)(p.temp_offset + s_ct[cl]));
            const uint4 b1 = st.block(1), b2 = st.block(2), b3 = st.block(3);
            const unsigned long long mm = word_to_int(lo_word(b1), bufsize);
            unsigned long long nn = word_to_int(hi_word(b1), bufsize);
            uint32_t j = 4;
            auto word = [&](uint32_t jj) {
                return jj == 4 ? lo_word(b2) : jj == 5 ? hi_word(b2) : jj == 6 ? lo_word(b3) : jj == 7 ? hi_word(b3) : stream_word(st, jj);
            };
            while (mm == nn) nn = word_to_int(word(j++), bufsize);
            const unsigned long long om = de_row_offset(mm, bufsize, W, p.burn, p.de_head) * (unsigned long long)d;
            const unsigned long long on = de_row_offset(nn, bufsize, W, p.burn, p.de_head) * (unsigned long long)d;
            s_rowm2[buf * nc + cl] = om;
            s_rown2[buf * nc + cl] = on;
            for (int b = 0; b < 8 * d; b += 128) {
                prefetch_l2(reinterpret_cast<const char *>(p.de + om) + b);
                prefetch_l2(reinterpret_cast<const char *>(p.de + on) + b);
            }
            prefetch_l2(p.de + om + d - 1);
            prefetch_l2(p.de + on + d - 1);
            const double prob = word_to_unit(word(j++));
            double scale = 1.0;
            if (!(prob > 0.5)) scale = word_to_unit(word(j++)) * 2.4 / sqrt(2.0 * d) * sqrt(1.0 / s_beta[cl]);
            s_sca2[buf * nc + cl] = scale;
            s_uword2[buf * nc + cl] = word(j++);
        }
    };
    // SCAM chains of iteration `it` (ref :839-873), one per lane: prob, k, normal, accept uniform = words 2..5; the step is
    // coef * U[:, k] (ref :868-873)
    auto draw_scam = [&](long long it, int slot) {
        const int buf = (int)(it & 1), nS = s_count[4 * slot + 1];
        for (int i = lane; i < nS; i += 32) {
            const int cl = s_list2[(slot * 3 + 1) * nc + i];
            const Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + s_cw[cl]),
                            (uint32_t)(p.temp_offset + s_ct[cl]));
            const uint4 b1 = st.block(1), b2 = st.block(2);
            const double prob = word_to_unit(lo_word(b1));
            const double scale = cov_jump_scale(prob, s_temp[cl]);
            const int k = (int)word_to_int(hi_word(b1), (unsigned long long)d);
            const double cd = 2.4 / sqrt(2.0) * scale;
            double z0, z1;
            word_to_normals(lo_word(b2), z0, z1);
            s_sca2[buf * nc + cl] = z0 * cd * sSs[k];
            s_rowm2[buf * nc + cl] = (unsigned long long)k * (unsigned long long)d;
            s_uword2[buf * nc + cl] = hi_word(b2);
        }
    };
    // buffers / record of iteration ib (ref :627) for the chain of thread tid < nc
    // AM chains of iteration `it` (ref :897-920), one per lane: word 2 -> the jump scale cd (the normals follow in phase R)
    auto draw_am = [&](long long it, int slot) {
        const int buf = (int)(it & 1), nA = s_count[4 * slot];
        for (int i = lane; i < nA; i += 32) {
            const int cl = s_list2[(slot * 3) * nc + i];
            const Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + s_cw[cl]),
                            (uint32_t)(p.temp_offset + s_ct[cl]));
            s_sca2[buf * nc + cl] = 2.4 / sqrt(2.0 * d) * cov_jump_scale(word_to_unit(lo_word(st.block(1))), s_temp[cl]);
        }
    };
    // few blocks have anything to write: the first chain's rung decides for the block (chains are rung-major)
    const bool bk_block = p.trace != nullptr || (int)(c0 / W) < max(p.ntr, (p.temp_offset == 0 && p.am != nullptr) ? 1 : 0);
    auto bookkeeping = [&](long long ib, int slot) {
        // nothing of this rare path stays in registers (or local memory) over the loop: the chain's coordinates come from
        // shared memory, the ring slot and record row from the iteration number
        if (bk_block && tid < nc && c0 + tid < TW) {
            const int tb = s_ct[tid], wb = s_cw[tid];
            if (p.trace && ib - 1 < p.trace_cap) p.trace[((size_t)(ib - 1) * T + tb) * W + wb] = s_jt2[slot * nc + tid];
            if (ib < p.it1 || p.tail) {
                if (tb == 0 && p.temp_offset == 0 && p.am != nullptr) {
                    double *dst = p.am + (size_t)(ib % p.cov_update) * d * W + wb;
                    for (int k = 0; k < d; ++k) dst[(size_t)k * W] = xs[tid * ld + k];
                }
                const long long row = ib / p.thin - p.rec_base;
                if (tb < p.ntr && ib % p.thin == 0 && row >= 0 && row < p.rec_cap) {
                    const size_t rr = ((size_t)row * p.ntr + tb) * W + wb;
                    double *dst = p.rec_x + rr * d;
                    for (int k = 0; k < d; ++k) dst[k] = xs[tid * ld + k];
                    p.rec_lnl[rr] = s_lnl[tid];
                    p.rec_lnp[rr] = s_beta[tid] * s_lnl[tid] + s_lp[tid];
                }
            }
        }
    };

    // the draws run ahead of the steps: jump kinds and lists two iterations, the DE / SCAM scalars one
    int l3 = (int)(p.it0 % 3);  // list slot of the current iteration
    const auto next3 = [](int s) { return s == 2 ? 0 : s + 1; };
    __syncthreads();
    if (warp == MMA_WARPS - 1 && p.it0 <= p.it1) pick_kinds(p.it0, l3);
    __syncthreads();
    if (p.it0 <= p.it1) {
        if (warp == MMA_WARPS - 1 && p.it0 < p.it1) pick_kinds(p.it0 + 1, next3(l3));
        if (warp == MMA_WARPS - 2) draw_de(p.it0, l3);
        if (warp == MMA_WARPS - 3) draw_scam(p.it0, l3);
        if (warp == MMA_WARPS - 4) draw_am(p.it0, l3);
    }
    __syncthreads();
#ifdef PTMCMC_MMA_CLOCKS
    long long clk_prev_ = clock64();
#endif

    for (long long it = p.it0; it <= p.it1; ++it, l3 = next3(l3)) {
        const int buf = (int)(it & 1);
        const unsigned short *s_list = s_list2 + l3 * 3 * nc;
        unsigned char *s_jt = s_jt2 + l3 * nc;
        double *s_sca = s_sca2 + buf * nc;
        const unsigned long long *s_rowm = s_rowm2 + buf * nc, *s_rown = s_rown2 + buf * nc;
        unsigned long long *s_uword = s_uword2 + buf * nc;
        const int nA = s_count[4 * l3], nS = s_count[4 * l3 + 1], nD = s_count[4 * l3 + 2];

        // ================= phase R: evenly spread tasks.  One gather task per (DE or SCAM chain, 8 columns) -- DE: B[mm] -
        // B[nn], SCAM: coef * row k of the transposed factor -- into the chain's zq row, and one task per AM (chain, Philox
        // block) pair.  A thread ISSUES the loads of its r-th gather task, runs its r-th AM task (normals from one Philox
        // block), and only then consumes the loads.
        PTM_WARP0
        PTM_SUB0
        PTM_SUB(24, tid == 0)
        if (warp == 0 && it > p.it0) bookkeeping(it - 1, l3 == 0 ? 2 : l3 - 1);
        PTM_SUB(11, tid == 0)
        {
            // AM tasks take two Philox blocks (generated side by side: the draw arithmetic is a latency chain) and count up
            // from thread 0, gather tasks count down from the last thread
            const int nG = nD + nS, tG = nG * NT, am_tasks2 = (am_tasks + 1) >> 1, tA = nA * am_tasks2;
            const bool vec = (d & 1) == 0;  // rows are 16-byte aligned
            PTM_SUB(5, tid == 0)
            for (int qa = tid, qg = MMA_THREADS - 1 - tid; qg < tG || qa < tA; qa += MMA_THREADS, qg += MMA_THREADS) {
                const bool hg = qg < tG, ha = qa < tA;
                double2 vm[4], vn[4];
                double coef = 1.0;
                double *dst = zq;
                if (hg) {
                    const int ci = qg % nG, seg = qg / nG;
                    const bool de = ci < nD;
                    const int cl = de ? s_list[2 * nc + ci] : s_list[nc + ci - nD];
                    const double *bm = (de ? p.de : a.Ut) + s_rowm[cl], *bn = p.de + (de ? s_rown[cl] : 0ull);
                    coef = de ? 1.0 : s_sca[cl];
                    const int col = 8 * seg;
                    dst = zq + cl * ld + col;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        vm[k] = vn[k] = make_double2(0.0, 0.0);
                        const int c = col + 2 * k;
                        if (vec) {
                            if (c < d) {
                                vm[k] = __ldg(reinterpret_cast<const double2 *>(bm + c));
                                if (de) vn[k] = __ldg(reinterpret_cast<const double2 *>(bn + c));
                            }
                        } else {
                            if (c < d) vm[k].x = __ldg(bm + c);
                            if (c + 1 < d) vm[k].y = __ldg(bm + c + 1);
                            if (de && c < d) vn[k].x = __ldg(bn + c);
                            if (de && c + 1 < d) vn[k].y = __ldg(bn + c + 1);
                        }
                    }
                }
                PTM_SUB(8, tid == 0)
                if (ha) {
                    // AM (ref :897-930): word 2 = prob, 3 + j = normal pair j, 3 + npairs = accept u
                    const int ai = qa % nA, b0 = 1 + 2 * (qa / nA);
                    const int cl = s_list[ai];
                    const double cd = s_sca[cl];
                    const uint32_t cw = (uint32_t)(p.walker_offset + s_cw[cl]), ct = (uint32_t)(p.temp_offset + s_ct[cl]);
                    uint4 blk[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) blk[e] = philox4x32_10(p, (uint32_t)it, (PURPOSE_MH << 24) | (uint32_t)(b0 + e), cw, ct);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int wi = 2 * (b0 + e) + h;
                            const uint64_t word = h ? hi_word(blk[e]) : lo_word(blk[e]);
                            if (wi == 2) {
                                // the jump scale was drawn ahead (draw_am)
                            } else if (wi < uword) {
                                // delta = z cd sqrt(S) (ref :923-926) is what phase P multiplies; padded columns get sqrt(S) = 0
                                double z0, z1;
                                word_to_normals(word, z0, z1);
                                const int j = 2 * (wi - 3);
                                const double2 sv = *reinterpret_cast<const double2 *>(sSs + j);
                                *reinterpret_cast<double2 *>(zq + cl * ld + j) = make_double2(z0 * cd * sv.x, z1 * cd * sv.y);
                            } else if (wi == uword) {
                                s_uword[cl] = word;
                            }
                        }
                    }
                }
                PTM_SUB(9, tid == 0)
                if (hg) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<double2 *>(dst + 2 * k) = make_double2(coef * (vm[k].x - vn[k].x), coef * (vm[k].y - vn[k].y));
                }
                PTM_SUB(10, tid == 0)
            }
        }
        PTM_WARP(16)
        __syncthreads();
        PTM_CLK(0)

        // ================= phase P: the AM chains: zq <- U (z * cd * sqrt(S)) on the tensor cores (q = x + U delta equals
        // the reference's U (U^T x + delta), ref :923-931).  Warp g takes n-tiles [g NT/8, (g+1) NT/8) of ALL AM tiles (<= 4:
        // a block holds <= 32 chains), so a fragment of the factor (from L2, PD k-tiles in flight) is fetched once per block.
        // Every warp reads whole z rows, so the results are written after a barrier.
        {
            constexpr int CN = (NT + MMA_WARPS - 1) / MMA_WARPS, MT = 2, PD = NT < 6 ? NT : 6;
            const int nTA = (nA + 7) >> 3;
            const int n0 = warp * NT / MMA_WARPS, cnt = (warp + 1) * NT / MMA_WARPS - n0;
            const double2 *uf = reinterpret_cast<const double2 *>(a.Uf) + lane + n0 * 32;
            PTM_SUB0
            for (int tp = 0; tp < nTA; tp += MT) {  // MT tiles per pass (a second pass only when > 16 chains drew AM)
                double acc[MT][CN][2];
                const double *zr[MT];
#pragma unroll
                for (int ta = 0; ta < MT; ++ta) {
                    const int ai = (tp + ta) * 8 + r;
                    const int cl = s_list[ai < nA ? ai : nA - 1];
                    zr[ta] = zq + cl * ld + 2 * t;
#pragma unroll
                    for (int j = 0; j < CN; ++j) acc[ta][j][0] = acc[ta][j][1] = 0.0;
                }
                if (cnt > 0) {
                    double2 b[PD][CN];
#pragma unroll
                    for (int s = 0; s < PD; ++s)
#pragma unroll
                        for (int j = 0; j < CN; ++j) b[s][j] = (j < cnt) ? __ldg(uf + (s * NT + j) * 32) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int kk = 0; kk < NT; ++kk) {
#pragma unroll
                        for (int ta = 0; ta < MT; ++ta) {
                            if (tp + ta < nTA) {
                                const double2 dl = *reinterpret_cast<const double2 *>(zr[ta] + 8 * kk);  // delta, scaled in phase R
#pragma unroll
                                for (int j = 0; j < CN; ++j) {
                                    if (j < cnt) {
                                        dmma884(acc[ta][j][0], acc[ta][j][1], dl.x, b[kk % PD][j].x);
                                        dmma884(acc[ta][j][0], acc[ta][j][1], dl.y, b[kk % PD][j].y);
                                    }
                                }
                            }
                        }
                        if (kk + PD < NT) {
#pragma unroll
                            for (int j = 0; j < CN; ++j)
                                if (j < cnt) b[kk % PD][j] = __ldg(uf + ((kk + PD) * NT + j) * 32);
                        }
                    }
                }
                PTM_SUB(12, tid == 0)
                __syncthreads();
#pragma unroll
                for (int ta = 0; ta < MT; ++ta) {
                    if ((tp + ta) * 8 + r < nA) {
#pragma unroll
                        for (int j = 0; j < CN; ++j)
                            if (j < cnt)
                                *reinterpret_cast<double2 *>(const_cast<double *>(zr[ta]) + 8 * (n0 + j)) =
                                    make_double2(acc[ta][j][0], acc[ta][j][1]);
                    }
                }
            }
        }
        __syncthreads();
        PTM_CLK(1)

        // ================= phase L: log-prior, quadratic form on the tensor cores (warp = (tile, half); half 1 leaves its
        // part in s_part), Hastings test
        {
            const bool active = warp < 2 * ntiles;
            const int tile = active ? warp % ntiles : 0, half = warp / ntiles;
            const int cl = tile * 8 + r;
            const bool live = active && c0 + cl < TW;
            const int jump = s_jt[cl] & 0x7F;
            const double mult = (live && jump == JUMP_DE) ? s_sca[cl] : 1.0;
            const double *xrow = xs + cl * ld, *zrow = zq + cl * ld;
            const double2 *pf = reinterpret_cast<const double2 *>(Pf) + lane;
            constexpr int ST = mma_split_at(NT, true), SF = mma_split_at(NT, false);
            bool inside = true;
            double part = 0.0;
            PTM_SUB0
            if (active) {
                if (half == 0)
                    part = tri ? quad_part<NT, 0, ST, true, true>(xrow, zrow, mult, mus, los, his, inclusive, pf, t, inside)
                               : quad_part<NT, 0, SF, false, true>(xrow, zrow, mult, mus, los, his, inclusive, pf, t, inside);
                else
                    part = tri ? quad_part<NT, ST, NT, true, false>(xrow, zrow, mult, mus, los, his, inclusive, pf, t, inside)
                               : quad_part<NT, SF, NT, false, false>(xrow, zrow, mult, mus, los, his, inclusive, pf, t, inside);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                if (half == 1 && t == 0) s_part[cl] = part;
            }
            PTM_SUB(13, tid == 0)
            __syncthreads();
            PTM_CLK(2)
            PTM_SUB_RESTART
            if (active && half == 0) {
                // the four lanes of a quad hold one chain: all must be inside
                const unsigned bal = __ballot_sync(0xffffffffu, inside);
                inside = ((bal >> (4 * r)) & 0xFu) == 0xFu;
                part += s_part[cl];
                const double lnln = tri ? p.g_offset - part : part + p.g_offset;
                const double beta = s_beta[cl];
                const double lpn = inside ? p.p_inside : neg_inf();
                const double lnpn = inside ? beta * lnln + lpn : neg_inf();  // ref :607-612
                const double lnp0 = beta * s_lnl[cl] + s_lp[cl];
                const double diff = lnpn - lnp0;
                const bool accept = live && hastings_accept(diff, s_uword[cl]);  // ref :614-616
                __syncwarp();  // every lane of the quad has read lnl / lp / jt before lane t == 0 updates them
                if (accept) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int col = 8 * nt + 2 * t;
                        const double2 x = *reinterpret_cast<const double2 *>(xrow + col);
                        const double2 z = *reinterpret_cast<const double2 *>(zrow + col);
                        *reinterpret_cast<double2 *>(xs + cl * ld + col) =
                            make_double2(__dadd_rn(x.x, __dmul_rn(mult, z.x)), __dadd_rn(x.y, __dmul_rn(mult, z.y)));
                    }
                }
                if (t == 0 && live) {
                    s_cnt[jump * nc + cl] += 1;
                    if (accept) {
                        s_lnl[cl] = inside ? lnln : 0.0;
                        s_lp[cl] = lpn;
                        s_cnt[(3 + jump) * nc + cl] += 1;
                        s_jt[cl] = (unsigned char)(jump | 0x80);
                    }
                }
                PTM_SUB(14, tid == 0)
            } else if (warp >= MMA_WARPS - 4 && it < p.it1) {
                // state-free draws under the Hastings tests of this iteration: warp 7 the jump kinds and lists of iteration
                // it+2, warp 6 the DE chains, warp 5 the SCAM chains and warp 4 the AM chains' jump scales of iteration it+1
                const int s1 = next3(l3);
                if (warp == MMA_WARPS - 1) {
                    if (it + 2 <= p.it1) pick_kinds(it + 2, next3(s1));
                    PTM_SUB(15, lane == 0)
                } else if (warp == MMA_WARPS - 2) {
                    draw_de(it + 1, s1);
                    PTM_SUB(7, lane == 0)
                } else if (warp == MMA_WARPS - 3) {
                    draw_scam(it + 1, s1);
                    PTM_SUB(6, lane == 0)
                } else {
                    draw_am(it + 1, s1);
                }
            }
        }
        __syncthreads();
        PTM_CLK(3)
    }
    if (warp == 0 && p.it1 >= p.it0) bookkeeping(p.it1, l3 == 0 ? 2 : l3 - 1);
    if (tid < nc && c0 + tid < TW) {
        const int tb = s_ct[tid], wb = s_cw[tid];
        const long long cb = c0 + tid;
        for (int k = 0; k < d; ++k) p.x[((size_t)tb * d + k) * W + wb] = xs[tid * ld + k];
        p.lnl[cb] = s_lnl[tid];
        p.lp[cb] = s_lp[tid];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            p.prop[(size_t)j * TW + cb] += s_cnt[j * nc + tid];
            p.acc[(size_t)j * TW + cb] += s_cnt[(3 + j) * nc + tid];
        }
    }
}

}  // namespace ptm
