// Tensor-core (fp64 DMMA) fused MH kernel for the dense Gaussian target, any ndim <= 128, identity
// parameter group, box / flat prior.
//
// The two dense pieces of a step -- the AM proposal x + U.delta (ref PTMCMCSampler.py :923-931) and
// the Gaussian quadratic form -1/2 (q-mu)^T icov (q-mu) (ref examples/simple.py:34-36) -- are the same
// small matrix applied to many chains' vectors, i.e. GEMMs  [chains x d] . [d x d].  They run on the
// tensor cores as mma.sync m8n8k4 f64 (DMMA; measured 37.1 TFLOP/s on B200, the DFMA peak, at 1/8 of the
// instruction count and with the matrix operand held in fragment order, so the inner loops issue no
// broadcast shared-memory loads).  tcgen05 has no fp64 kind; DMMA is the fp64 tensor path on sm_100a.
//
// Layout.  A tile is 8 chains (the m dimension).  Lane l = 4r + t of a warp holds, for chain r of the
// tile, columns {8nt + 2t, 8nt + 2t + 1} of every n-tile nt: the m8n8 C-fragment layout.  The k index of
// an MMA is only a summation index, so k-step (kk, e) is defined to use columns 8kk + 2t + e: then the
// C layout IS the A layout and a vector produced by one MMA feeds the next without any shuffle, and the
// matrices are pre-arranged in that fragment order (frag_build_kernel).  Chain state, proposals and the
// AM normals live in shared memory as rows [chain][LD], LD = 8 mod 16 doubles (conflict-free 128-bit
// fragment accesses).
//
// Per iteration, per block (NC chains, 4 block barriers):
//   A  thread c: buffers/record of iteration it-1 for chain c, jump kind of iteration it, per-kind lists
//   R  every draw after the jump index, as tasks spread evenly over all threads: (AM chain, Philox
//      block) -> normals into zq; SCAM chain -> step coef * U[:, k] into zq; DE chain -> gathers its two
//      history rows and leaves scale * (B[mm] - B[nn]) in zq (the gathers overlap other tasks' math)
//   P  AM chains gathered 8 at a time into dense tiles: zq <- U (z * cd * sqrt(S)) on the tensor cores
//   L  per tile: q = x + zq, box test, quadratic form by DMMA, quad reduction, Hastings test, state update
// Draw order and arithmetic of every scalar follow the thread-per-chain kernels (mh_kernels.cuh), so the
// jump / accept / swap streams are identical; the quadratic form differs in summation order only.
#pragma once
#include "mh_common.cuh"
#include "mma_f64.cuh"

namespace ptm {

constexpr int MMA_THREADS = 256;
constexpr int MMA_WARPS = MMA_THREADS / 32;

// -DPTMCMC_MMA_CLOCKS: thread 0 of every block adds the clocks between the block barriers to g_mma_clk[phase]
// (development aid, read with ptmcmc_debug_mma_clocks; the barrier release times are the same for every warp)
#ifdef PTMCMC_MMA_CLOCKS
__device__ unsigned long long g_mma_clk[32];
#define PTM_CLK(i)                                                         \
    if (threadIdx.x == 0) {                                                \
        const long long now_ = clock64();                                  \
        atomicAdd(&g_mma_clk[i], (unsigned long long)(now_ - clk_prev_));  \
        clk_prev_ = now_;                                                  \
    }
// finer marks inside a phase, by one chosen thread: PTM_SUB0 starts, PTM_SUB(i) adds the clocks since the last mark
#define PTM_SUB0 long long sub_prev_ = clock64();
#define PTM_SUB(i, cond)                                                   \
    if (cond) {                                                            \
        const long long now_ = clock64();                                  \
        atomicAdd(&g_mma_clk[i], (unsigned long long)(now_ - sub_prev_));  \
        sub_prev_ = now_;                                                  \
    }
#define PTM_SUB_RESTART sub_prev_ = clock64();
// per-warp arrival at the end of a phase: PTM_WARP0 after the barrier that starts it, PTM_WARP(base) before the one ending it
#define PTM_WARP0 const long long warp_t0_ = clock64();
#define PTM_WARP(base) \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_mma_clk[(base) + (threadIdx.x >> 5)], (unsigned long long)(clock64() - warp_t0_));
#else
#define PTM_WARP0
#define PTM_WARP(base)
#define PTM_CLK(i)
#define PTM_SUB0
#define PTM_SUB(i, cond)
#define PTM_SUB_RESTART
#endif

// Fragment-order image of the d x d matrix M used as the B operand of  Y[c][n] = sum_k A[c][k] M(k, n):
//   out[((tile(kk, nt))*32 + lane)*2 + e] = M(8kk + 2(lane&3) + e, 8nt + (lane>>2)),  zero padded,
// tile(kk, nt) = kk*NT + nt, or kk(kk+1)/2 + nt over the tiles kk >= nt only when `packed` (a lower-triangular M).
// transpose: M(k, n) = src[n*d + k] (the AM mat-vec uses U^T), else src[k*d + n].
__host__ __device__ inline int mma_tiles(int NT, bool packed) { return packed ? NT * (NT + 1) / 2 : NT * NT; }

__global__ void frag_build_kernel(const double *src, int d, int NT, int transpose, int packed, double *out)
{
    const int total = mma_tiles(NT, packed != 0) * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int e = idx & 1, lane = (idx >> 1) & 31, tile = idx >> 6;
        int nt = tile % NT, kk = tile / NT;
        if (packed) {
            kk = 0;
            while ((kk + 1) * (kk + 2) / 2 <= tile) ++kk;
            nt = tile - kk * (kk + 1) / 2;
        }
        const int k = 8 * kk + 2 * (lane & 3) + e, n = 8 * nt + (lane >> 2);
        double v = 0.0;
        if (k < d && n < d) v = transpose ? src[n * d + k] : src[k * d + n];
        out[idx] = v;
    }
}

// shared-memory carve-up (all offsets in bytes, 16-byte aligned); see mma_smem_bytes()
struct MmaLayout {
    int xs, zq, pf, uf, ss, mu, lo, hi, lnl, lp, temp, beta, sca, logu, rowm, rown, part, ct, cw, cnt, list, jt, count, total;
    __host__ __device__ int take(int bytes)
    {
        const int at = total;
        total += (bytes + 15) & ~15;
        return at;
    }
};

// pf_tiles: 8x8 tiles of the form's fragment image held in shared memory (mma_tiles); the per-kind lists and jump ids
// are triple-buffered (the split kernel draws the jump kinds two iterations ahead)
__host__ __device__ inline MmaLayout mma_layout(int NT, int nc, int ld, bool usmem, int pf_tiles)
{
    MmaLayout L;
    L.total = 0;
    const int KP = 8 * NT;
    L.xs = L.take(nc * ld * 8);
    L.zq = L.take(nc * ld * 8);
    L.pf = L.take(pf_tiles * 64 * 8);
    L.uf = L.take(usmem ? NT * NT * 64 * 8 : 0);
    L.ss = L.take(KP * 8); L.mu = L.take(KP * 8); L.lo = L.take(KP * 8); L.hi = L.take(KP * 8);
    L.lnl = L.take(nc * 8); L.lp = L.take(nc * 8); L.temp = L.take(nc * 8); L.beta = L.take(nc * 8);
    L.sca = L.take(2 * nc * 8); L.logu = L.take(2 * nc * 8);  // [2]: the split kernel draws them an iteration ahead
    L.rowm = L.take(2 * nc * 8); L.rown = L.take(2 * nc * 8); L.part = L.take(nc * 8);
    L.ct = L.take(nc * 4); L.cw = L.take(nc * 4);
    L.cnt = L.take(6 * nc * 4);
    L.list = L.take(3 * 3 * nc * 2);
    L.jt = L.take(3 * nc);
    L.count = L.take(16 * 4);
    return L;
}

__global__ void transpose_kernel(const double *src, int d, double *dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < d * d) dst[(idx % d) * d + idx / d] = src[idx];
}

struct MmaArgs {
    const double *Uf;     // [NT*NT*64] fragment order of U^T   (global)
    const double *Pf;     // [NT*NT*64] fragment order of -1/2 sym(icov), or of L / sqrt(2) when tri (global)
    const double *Ut;     // [d*d] U transposed: row k = eigenvector k (global; SCAM reads one row)
    int nc;               // chains per block (multiple of 8)
    int ld;               // row stride of the per-chain shared-memory rows (doubles, = 8 mod 16)
    int tri;              // 1: Pf holds the Cholesky factor L of sym(icov) (lower triangular) scaled by 1/sqrt(2):
                          //    lnl = offset - |d^T L / sqrt(2)|^2, and the all-zero tiles kk < nt are skipped
    int pf_tiles;         // tiles of Pf: NT*NT, or NT(NT+1)/2 in packed lower-triangular order (split kernel with tri)
    MmaLayout L;          // computed on the host: the offsets are then plain constant-bank operands
};

// MINB (blocks per SM the register allocation must allow): small ndim runs many small blocks to hide the
// fp64 latency of the draw phase; large ndim needs the registers for the fragment arrays.
template <int NT, bool USMEM, int MINB>
__global__ void __launch_bounds__(MMA_THREADS, MINB) mh_mma_kernel(const __grid_constant__ DevParams p,
                                                             const __grid_constant__ MmaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KP = 8 * NT;
    const int d = p.d, W = p.W, T = p.T;
    const int nc = a.nc, ld = a.ld;
    const bool tri = a.tri != 0;
    const MmaLayout &L = a.L;
    double *xs = reinterpret_cast<double *>(smem_raw + L.xs);
    double *zq = reinterpret_cast<double *>(smem_raw + L.zq);
    double *Pf = reinterpret_cast<double *>(smem_raw + L.pf);
    double *Ufs = reinterpret_cast<double *>(smem_raw + L.uf);
    double *sSs = reinterpret_cast<double *>(smem_raw + L.ss);
    double *mus = reinterpret_cast<double *>(smem_raw + L.mu);
    double *los = reinterpret_cast<double *>(smem_raw + L.lo);
    double *his = reinterpret_cast<double *>(smem_raw + L.hi);
    double *s_lnl = reinterpret_cast<double *>(smem_raw + L.lnl);
    double *s_lp = reinterpret_cast<double *>(smem_raw + L.lp);
    double *s_temp = reinterpret_cast<double *>(smem_raw + L.temp);
    double *s_beta = reinterpret_cast<double *>(smem_raw + L.beta);
    double *s_sca = reinterpret_cast<double *>(smem_raw + L.sca);    // AM: cd; SCAM: coefficient; DE: scale
    unsigned long long *s_rowm = reinterpret_cast<unsigned long long *>(smem_raw + L.rowm);  // DE row offsets;
    unsigned long long *s_rown = reinterpret_cast<unsigned long long *>(smem_raw + L.rown);  // SCAM: rowm = k
    unsigned long long *s_uword = reinterpret_cast<unsigned long long *>(smem_raw + L.logu);  // accept-uniform word
    int *s_ct = reinterpret_cast<int *>(smem_raw + L.ct);
    int *s_cw = reinterpret_cast<int *>(smem_raw + L.cw);
    unsigned *s_cnt = reinterpret_cast<unsigned *>(smem_raw + L.cnt);
    unsigned short *s_list = reinterpret_cast<unsigned short *>(smem_raw + L.list);
    unsigned char *s_jt = smem_raw + L.jt;
    int *s_count = reinterpret_cast<int *>(smem_raw + L.count);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, t = lane & 3;

    // ---- stage tables and state
    for (int idx = tid; idx < NT * NT * 64; idx += MMA_THREADS) {
        Pf[idx] = a.Pf[idx];
        if (USMEM) Ufs[idx] = a.Uf[idx];
    }
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
        s_jt[tid] = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) s_cnt[j * nc + tid] = 0;
    }
    if (tid < 8) s_count[tid] = 0;
    const int inclusive = p.p_inclusive;
    long long am_slot = p.it0 % p.cov_update, thin_ctr = p.it0 % p.thin, row = p.it0 / p.thin - p.rec_base;
    const bool cold = have && tme == 0 && p.temp_offset == 0 && p.am != nullptr;
    const bool recorded = have && tme < p.ntr;
    const int npairs = (d + 1) >> 1, uword = 3 + npairs, am_tasks = ((uword + 2) >> 1) - 1;
    const unsigned long long bufsize = (unsigned long long)p.burn * (unsigned long long)W;
    const int ntiles = nc >> 3;
    __syncthreads();

    for (long long it = p.it0; it <= p.it1 + 1; ++it) {
        // ================= phase A: bookkeeping of it-1 (ref :627), jump kind of it (ref :1058), lists
        if (have && it > p.it0) {
            const long long ib = it - 1;
            if (p.trace && ib - 1 < p.trace_cap) p.trace[((size_t)(ib - 1) * T + tme) * W + wme] = s_jt[tid];
            if (ib < p.it1 || p.tail) {
                if (cold) {
                    double *dst = p.am + (size_t)am_slot * d * W + wme;
                    for (int k = 0; k < d; ++k) dst[(size_t)k * W] = xs[tid * ld + k];
                }
                if (recorded && thin_ctr == 0 && row >= 0 && row < p.rec_cap) {
                    const size_t rr = ((size_t)row * p.ntr + tme) * W + wme;
                    double *dst = p.rec_x + rr * d;
                    for (int k = 0; k < d; ++k) dst[k] = xs[tid * ld + k];
                    p.rec_lnl[rr] = s_lnl[tid];
                    p.rec_lnp[rr] = s_beta[tid] * s_lnl[tid] + s_lp[tid];
                }
            }
        }
        if (it > p.it0) {
            if (++am_slot == p.cov_update) am_slot = 0;
            if (++thin_ctr == p.thin) { thin_ctr = 0; ++row; }
        }
        if (it > p.it1) break;
        int *count = s_count + 4 * (int)(it & 1);
        int kind = 3;
        if (have) {
            Stream st(p, PURPOSE_MH, (unsigned long long)it, (uint32_t)(p.walker_offset + wme),
                      (uint32_t)(p.temp_offset + tme));
            const int jump = pick_jump(p, st);
            s_jt[tid] = (unsigned char)jump;
            kind = (jump == JUMP_AM) ? 0 : (jump == JUMP_SCAM) ? 1 : 2;
            // Every scalar draw of a SCAM or DE step is made here, by the chain's own thread, before any state is touched:
            // phase R is then nothing but evenly spread tasks (gathers and AM normals), none of them a long serial one.
            if (kind == 1) {  // SCAM (ref :839-873): prob, k, normal, accept uniform = words 2..5
                st.j = 2;
                const double prob = word_to_unit(st.next());
                const double scale = cov_jump_scale(prob, s_temp[tid]);
                const int k = (int)word_to_int(st.next(), (unsigned long long)d);
                const double cd = 2.4 / sqrt(2.0) * scale;
                double z0, z1;
                word_to_normals(st.next(), z0, z1);
                s_sca[tid] = z0 * cd * sSs[k];  // the step is coef * U[:, k] (ref :868-873)
                s_rowm[tid] = (unsigned long long)k * (unsigned long long)d;
                s_uword[tid] = st.next();
            } else if (kind == 2) {
                // DE (ref :955-976): the two history rows (words 2, 3, redrawn while equal), prob, scale, accept uniform;
                // the rows are prefetched into L2 for the gather tasks of phase R
                st.j = 2;
                const unsigned long long mm = word_to_int(st.next(), bufsize);
                unsigned long long nn = word_to_int(st.next(), bufsize);
                while (mm == nn) nn = word_to_int(st.next(), bufsize);
                const unsigned long long om = de_row_offset(mm, bufsize, W, p.burn, p.de_head) * (unsigned long long)d;
                const unsigned long long on = de_row_offset(nn, bufsize, W, p.burn, p.de_head) * (unsigned long long)d;
                s_rowm[tid] = om;
                s_rown[tid] = on;
                for (int b = 0; b < 8 * d; b += 128) {
                    prefetch_l2(reinterpret_cast<const char *>(p.de + om) + b);
                    prefetch_l2(reinterpret_cast<const char *>(p.de + on) + b);
                }
                prefetch_l2(p.de + om + d - 1);
                prefetch_l2(p.de + on + d - 1);
                const double prob = word_to_unit(st.next());
                double scale = 1.0;
                if (!(prob > 0.5)) scale = word_to_unit(st.next()) * 2.4 / sqrt(2.0 * d) * sqrt(1.0 / s_beta[tid]);
                s_sca[tid] = scale;
                s_uword[tid] = st.next();
            }
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const unsigned m = __ballot_sync(0xffffffffu, kind == kk);
            if (m) {
                int base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(&count[kk], __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (kind == kk) s_list[kk * nc + base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)tid;
            }
        }
        __syncthreads();
        const int nA = count[0], nS = count[1], nD = count[2];
        if (tid < 4) s_count[4 * (int)((it + 1) & 1) + tid] = 0;

        // ================= phase R: every draw after word 1, as evenly spread tasks
        {
            // One gather task per (DE or SCAM chain, 8 columns) -- DE: B[mm] - B[nn], SCAM: coef * row k of the transposed
            // factor -- into the chain's zq row, and one task per AM (chain, Philox block) pair.  A thread ISSUES the loads of
            // its r-th gather task, runs its r-th AM task (normals from one Philox block), and only then consumes the loads:
            // the gather latency (L2 after the prefetch of phase A) is spent on the draw arithmetic.
            const int nG = nD + nS, tG = nG * NT, tA = nA * am_tasks;
            const bool vec = (d & 1) == 0;  // rows are 16-byte aligned
            for (int q = tid; q < tG || q < tA; q += MMA_THREADS) {
                const bool hg = q < tG, ha = q < tA;
                double2 vm[4], vn[4];
                double coef = 1.0;
                double *dst = zq;
                int col = 0;
                if (hg) {
                    const int ci = q % nG, seg = q / nG;
                    const bool de = ci < nD;
                    const int cl = de ? s_list[2 * nc + ci] : s_list[nc + ci - nD];
                    const double *bm = (de ? p.de : a.Ut) + s_rowm[cl], *bn = p.de + (de ? s_rown[cl] : 0ull);
                    coef = de ? 1.0 : s_sca[cl];
                    col = 8 * seg;
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
                if (ha) {
                    // AM (ref :897-930): word 2 = prob, 3 + j = normal pair j, 3 + npairs = accept u
                    const int ai = q % nA, b = 1 + q / nA;
                    const int cl = s_list[ai];
                    const uint4 blk = philox4x32_10(p, (uint32_t)it, (PURPOSE_MH << 24) | (uint32_t)b,
                                                    (uint32_t)(p.walker_offset + s_cw[cl]), (uint32_t)(p.temp_offset + s_ct[cl]));
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int wi = 2 * b + h;
                        const uint64_t word = h ? ((uint64_t)blk.z | ((uint64_t)blk.w << 32)) : ((uint64_t)blk.x | ((uint64_t)blk.y << 32));
                        if (wi == 2) {
                            s_sca[cl] = 2.4 / sqrt(2.0 * d) * cov_jump_scale(word_to_unit(word), s_temp[cl]);
                        } else if (wi < uword) {
                            double z0, z1;
                            word_to_normals(word, z0, z1);
                            const int j = 2 * (wi - 3);
                            zq[cl * ld + j] = z0;
                            if (j + 1 < KP) zq[cl * ld + j + 1] = z1;
                        } else if (wi == uword) {
                            s_uword[cl] = word;
                        }
                    }
                }
                if (hg) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<double2 *>(dst + 2 * k) = make_double2(coef * (vm[k].x - vn[k].x), coef * (vm[k].y - vn[k].y));
                }
            }
        }
        __syncthreads();

        // ================= phase P: the AM chains, 8 at a time: zq <- U (z * cd * sqrt(S)) on the tensor
        // cores (q = x + U delta equals the reference's U (U^T x + delta), ref :923-931)
        const int nTA = (nA + 7) >> 3;
        // few AM tiles (large ndim: 8 tiles per block, ~3 of them AM): split each tile's n-tiles over NG warps
        // so that all warps work; every item reads the whole z row, so the results are written after a barrier
        const int NG = (nTA > 0 && nTA <= MMA_WARPS / 2) ? min(NT, MMA_WARPS / nTA) : 1;
        if (NG > 1) {
            const int ta = warp / NG, g = warp % NG, NTG = (NT + NG - 1) / NG;
            const int n0 = g * NTG, n1 = min(NT, n0 + NTG);
            const bool item = ta < nTA && n0 < n1;
            const int ai = ta * 8 + r;
            const bool live = item && ai < nA;
            const int cl = s_list[(item && ai < nA) ? ai : nA - 1];
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
            if (item) {
                const double cd = s_sca[cl];
                double dl[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int col = 8 * nt + 2 * t;
                    const double2 z = *reinterpret_cast<const double2 *>(zq + cl * ld + col);
                    const double2 sv = *reinterpret_cast<const double2 *>(sSs + col);
                    dl[nt][0] = (live && col < d) ? z.x * cd * sv.x : 0.0;
                    dl[nt][1] = (live && col + 1 < d) ? z.y * cd * sv.y : 0.0;
                }
                const double2 *uf = reinterpret_cast<const double2 *>(USMEM ? Ufs : a.Uf) + lane;
#pragma unroll
                for (int kk = 0; kk < NT; ++kk) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        if (nt >= n0 && nt < n1) {
                            const double2 b = USMEM ? uf[(kk * NT + nt) * 32] : __ldg(uf + (kk * NT + nt) * 32);
                            dmma884(acc[nt][0], acc[nt][1], dl[kk][0], b.x);
                            dmma884(acc[nt][0], acc[nt][1], dl[kk][1], b.y);
                        }
                    }
                }
            }
            __syncthreads();
            if (live) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    if (nt >= n0 && nt < n1)
                        *reinterpret_cast<double2 *>(zq + cl * ld + 8 * nt + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);
            }
        } else
        for (int ta = warp; ta < nTA; ta += MMA_WARPS) {
            const int ai = ta * 8 + r;
            const bool live = ai < nA;
            const int cl = s_list[live ? ai : nA - 1];
            const double cd = s_sca[cl];
            double dl[NT][2], acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = 8 * nt + 2 * t;
                const double2 z = *reinterpret_cast<const double2 *>(zq + cl * ld + col);
                const double2 s = *reinterpret_cast<const double2 *>(sSs + col);
                dl[nt][0] = (live && col < d) ? z.x * cd * s.x : 0.0;
                dl[nt][1] = (live && col + 1 < d) ? z.y * cd * s.y : 0.0;
                acc[nt][0] = 0.0;
                acc[nt][1] = 0.0;
            }
            const double2 *uf = reinterpret_cast<const double2 *>(USMEM ? Ufs : a.Uf) + lane;
#pragma unroll
            for (int kk = 0; kk < NT; ++kk) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 b = USMEM ? uf[(kk * NT + nt) * 32] : __ldg(uf + (kk * NT + nt) * 32);
                    dmma884(acc[nt][0], acc[nt][1], dl[kk][0], b.x);
                    dmma884(acc[nt][0], acc[nt][1], dl[kk][1], b.y);
                }
            }
            if (live) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2 *>(zq + cl * ld + 8 * nt + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);
            }
        }
        __syncthreads();

        // ================= phase L: log-prior, quadratic form on the tensor cores, Hastings test
        for (int tile = warp; tile < ntiles; tile += MMA_WARPS) {
            const int cl = tile * 8 + r;
            const bool live = c0 + cl < TW;
            const int jump = s_jt[cl] & 0x7F;
            // the step q - x of this lane's columns, from the chain's zq row for every kind: AM = U delta (phase P), SCAM =
            // coef * U[:, k] and DE = B[mm] - B[nn] (phase R; DE times its scale, ref :978-983)
            double stp[NT][2];
            {
                const double mult = (live && jump == JUMP_DE) ? s_sca[cl] : 1.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 z = *reinterpret_cast<const double2 *>(zq + cl * ld + 8 * nt + 2 * t);
                    stp[nt][0] = mult * z.x;
                    stp[nt][1] = mult * z.y;
                }
            }
            double q[NT][2], dv[NT][2];
            bool inside = true;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = 8 * nt + 2 * t;
                const double2 st = make_double2(stp[nt][0], stp[nt][1]);
                const double2 x = *reinterpret_cast<const double2 *>(xs + cl * ld + col);
                const double2 v = make_double2(x.x + st.x, x.y + st.y);
                const double2 m = *reinterpret_cast<const double2 *>(mus + col);
                const double2 lo = *reinterpret_cast<const double2 *>(los + col);
                const double2 hi = *reinterpret_cast<const double2 *>(his + col);
                q[nt][0] = v.x; q[nt][1] = v.y;
                dv[nt][0] = v.x - m.x; dv[nt][1] = v.y - m.y;
                inside = inside && in_box(v.x, lo.x, hi.x, inclusive) && in_box(v.y, lo.y, hi.y, inclusive);
            }
            // the four lanes of a quad hold one chain: all must be inside
            const unsigned bal = __ballot_sync(0xffffffffu, inside);
            inside = ((bal >> (4 * r)) & 0xFu) == 0xFu;
            double part = 0.0;
            const double2 *pf = reinterpret_cast<const double2 *>(Pf) + lane;
            constexpr int NB = NT < 4 ? NT : 4;  // independent accumulators in flight (DMMA latency)
#pragma unroll
            for (int nb = 0; nb < NT; nb += NB) {
                double y[NB][2];
#pragma unroll
                for (int j = 0; j < NB; ++j) y[j][0] = y[j][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < NT; ++kk) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (nb + j < NT && (!tri || kk >= nb + j)) {
                            const double2 b = pf[(kk * NT + nb + j) * 32];
                            dmma884(y[j][0], y[j][1], dv[kk][0], b.x);
                            dmma884(y[j][0], y[j][1], dv[kk][1], b.y);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (nb + j < NT) {
                        part = fma(y[j][0], tri ? y[j][0] : dv[nb + j][0], part);
                        part = fma(y[j][1], tri ? y[j][1] : dv[nb + j][1], part);
                    }
                }
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
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
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2 *>(xs + cl * ld + 8 * nt + 2 * t) = make_double2(q[nt][0], q[nt][1]);
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
        }
        __syncthreads();
    }
    if (have) {
        for (int k = 0; k < d; ++k) p.x[((size_t)tme * d + k) * W + wme] = xs[tid * ld + k];
        p.lnl[cme] = s_lnl[tid];
        p.lp[cme] = s_lp[tid];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            p.prop[(size_t)j * TW + cme] += s_cnt[j * nc + tid];
            p.acc[(size_t)j * TW + cme] += s_cnt[(3 + j) * nc + tid];
        }
    }
}

}  // namespace ptm
