// Pooled adaptive state: running mean / second moment of the cold walkers' samples, covariance,
// per-group eigen-factor (stand-in for np.linalg.svd of a symmetric PSD block), DE history.
//
// Follows ref PTMCMCSampler.py _updateRecursive :769-803 (Welford over the covUpdate buffered
// rows; here the batch of covUpdate x W pooled samples is reduced in one tensor-core pass and merged into
// the running (n, mu, M2) with Chan's formula, which equals the sequential recursion up to rounding),
// _updateDEbuffer :806-817 and shift_array :27-37.
#pragma once
#include "mma_f64.cuh"
#include "params.h"

namespace ptm {

// ---------------------------------------------------------------------------------------------
// Single-pass pooled moments on the tensor cores.  The batch of covUpdate x W cold samples is a tall
// matrix X [N][d]; with a column of ones appended, the Gram matrix G = [X-c, 1]^T [X-c, 1] holds the
// shifted second moments, the shifted sums (last column) and N (corner) at once.  G is accumulated by
// DMMA m8n8k4 with the sample index as k: for a k-step of 4 samples the fragment
//     f[t] = tile[8t + (lane>>2)][s0 + (lane&3)]
// is at the same time the A operand of row-tile t and the B operand of column-tile t, so each 8x8
// output tile costs one DMMA and the AM ring is read exactly once (HBM bound: 1.3 GB at C2).
// c = the running mean before this batch (zeros for the first), for conditioning only.
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_TW = 64;        // walkers per staged tile
constexpr int GRAM_LDT = 68;       // tile row stride (doubles), = 4 mod 16: conflict-free fragment loads
constexpr int GRAM_MAXT = 20;      // upper 8x8 tiles per warp: covers KP <= 136 (ndim <= 135)

// part[block][KP*KP]: upper tiles (mt <= nt) of this block's partial Gram matrix.
// MAXT = upper tiles a warp may own (register arrays): 1 covers KP <= 24, 4 covers KP <= 56.
template <int MAXT>
__global__ void __launch_bounds__(GRAM_THREADS) moments_gram_kernel(const double *am, int d, int W, long long nslots,
                                                                    const double *shift, int KP, double *part)
{
    extern __shared__ double tile[];  // [KP][GRAM_LDT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NTg = KP >> 3, nup = NTg * (NTg + 1) / 2;
    int mtv[MAXT], ntv[MAXT];
    double acc[MAXT][2];
    int mine = 0;
    {
        int idx = 0;
        for (int mt = 0; mt < NTg; ++mt)
            for (int nt = mt; nt < NTg; ++nt, ++idx)
                if (idx % (GRAM_THREADS / 32) == warp) {
#pragma unroll
                    for (int i = 0; i < MAXT; ++i)
                        if (i == mine) { mtv[i] = mt; ntv[i] = nt; }
                    ++mine;
                }
        (void)nup;
    }
#pragma unroll
    for (int i = 0; i < MAXT; ++i) acc[i][0] = acc[i][1] = 0.0;
    const long long ntw = (W + GRAM_TW - 1) / GRAM_TW;
    const double *base = tile + (lane >> 2) * GRAM_LDT + (lane & 3);
    if constexpr (MAXT <= 4) {
        // small KP: a thread's share of a tile fits in registers, so the loads of the NEXT item are in flight while the
        // current one is multiplied (one memory round trip per item instead of two plus the compute)
        constexpr int NV = (MAXT == 1 ? 24 : 56) * GRAM_TW / GRAM_THREADS;
        // a thread's elements keep their parameter index k from item to item: shift, validity and the tile address are
        // fixed, and the loads carry no dependent arithmetic (the shift is subtracted when the values are stored)
        double v[NV], sh[NV];
        int kk[NV];
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int idx = tid + u * GRAM_THREADS, k = idx / GRAM_TW;
            kk[u] = idx < KP * GRAM_TW ? k : KP;
            sh[u] = k < d ? shift[k] : 0.0;
        }
        auto fetch = [&](long long item) {
            const long long slot = item / ntw;
            const int w0 = (int)(item % ntw) * GRAM_TW;
            const double *src = am + (size_t)slot * d * W + w0;
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int ww = (tid + u * GRAM_THREADS) % GRAM_TW;
                const bool in = w0 + ww < W;
                // (a walker beyond W contributes nothing: its value is the shift itself)
                v[u] = (in && kk[u] < d) ? __ldg(src + (size_t)kk[u] * W + ww) : ((in && kk[u] == d) ? 1.0 : sh[u]);
            }
        };
        const long long total = nslots * ntw;
        if (blockIdx.x < total) fetch(blockIdx.x);
        for (long long item = blockIdx.x; item < total; item += gridDim.x) {
            __syncthreads();  // the previous tile has been multiplied
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int idx = tid + u * GRAM_THREADS;
                if (kk[u] < KP) tile[(idx / GRAM_TW) * GRAM_LDT + idx % GRAM_TW] = v[u] - sh[u];
            }
            __syncthreads();
            if (item + gridDim.x < total) fetch(item + gridDim.x);
#pragma unroll 4
            for (int s0 = 0; s0 < GRAM_TW; s0 += 4) {
#pragma unroll
                for (int i = 0; i < MAXT; ++i) {
                    if (i < mine) {
                        const double a = base[mtv[i] * 8 * GRAM_LDT + s0], b = base[ntv[i] * 8 * GRAM_LDT + s0];
                        dmma884(acc[i][0], acc[i][1], a, b);
                    }
                }
            }
        }
    } else
    for (long long item = blockIdx.x; item < nslots * ntw; item += gridDim.x) {
        const long long slot = item / ntw;
        const int w0 = (int)(item % ntw) * GRAM_TW;
        __syncthreads();
        // stage [KP][64 walkers]; four independent loads in flight per thread
        for (int idx0 = tid; idx0 < KP * GRAM_TW; idx0 += 4 * GRAM_THREADS) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * GRAM_THREADS, k = idx / GRAM_TW, ww = idx % GRAM_TW;
                const bool in = w0 + ww < W && idx < KP * GRAM_TW;
                v[u] = 0.0;
                if (in && k < d) v[u] = __ldg(am + ((size_t)slot * d + k) * W + w0 + ww) - __ldg(shift + k);
                else if (in && k == d) v[u] = 1.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * GRAM_THREADS;
                if (idx < KP * GRAM_TW) tile[(idx / GRAM_TW) * GRAM_LDT + idx % GRAM_TW] = v[u];
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int s0 = 0; s0 < GRAM_TW; s0 += 4) {
#pragma unroll
            for (int i = 0; i < MAXT; ++i) {
                if (i < mine) {
                    const double a = base[mtv[i] * 8 * GRAM_LDT + s0], b = base[ntv[i] * 8 * GRAM_LDT + s0];
                    dmma884(acc[i][0], acc[i][1], a, b);
                }
            }
        }
    }
    double *out = part + (size_t)blockIdx.x * KP * KP;
#pragma unroll
    for (int i = 0; i < MAXT; ++i) {
        if (i < mine) {
            const int row = 8 * mtv[i] + (lane >> 2), col = 8 * ntv[i] + 2 * (lane & 3);
            out[row * KP + col] = acc[i][0];
            out[row * KP + col + 1] = acc[i][1];
        }
    }
}

// G[idx] = sum over blocks of part[block][idx], one warp per entry, lanes stride over the blocks and the
// lane sums are combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) moments_gram_sum_kernel(const double *part, int nblocks, int KP, double *G)
{
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= KP * KP) return;
    double s = 0.0;
    for (int k = lane; k < nblocks; k += 32) s += part[(size_t)k * KP * KP + idx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) G[idx] = s;
}

// batch = {n, mean[d], M2c[d*d]} from the Gram matrix G (upper tiles valid): mean = c + S1/n,
// M2c = S2 - S1 S1^T / n
__global__ void moments_gram_batch_kernel(const double *G, int KP, int d, const double *shift, double *batch)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d * d) return;
    const int i = idx / d, j = idx % d;
    const int a = i <= j ? i : j, b = i <= j ? j : i;
    const double n = G[d * KP + d], s1a = G[a * KP + d], s1b = G[b * KP + d];
    batch[1 + d + idx] = G[a * KP + b] - s1a * s1b / n;
    if (i == j) batch[1 + i] = shift[i] + s1a / n;
    if (idx == 0) batch[0] = n;
}

// Jacobi eigen-decomposition of the n x n symmetric matrix a (destroyed); v receives the eigenvectors.
// Parallel (round-robin) ordering: a sweep is m-1 steps (m = n rounded up to even), step s rotating the m/2
// disjoint index pairs of round s of the circle tournament at once -- the rotation of a pair depends only on
// its own 2x2 block, which the other rotations of the step do not touch.  A step is  A <- J^T A J  done as a
// column phase (A J and V J) and a row phase (J^T A), each entry touched by exactly one rotation per phase, so
// the result does not depend on the order in which a phase is executed: the CPU oracle (orc_sym_factor) runs
// the same phases sequentially with the same explicitly fused operations and gets the same bits.
constexpr int JAC_THREADS = 1024;  // one block, alone on its SM: the rotation phases are latency bound (256 threads: 14 ms at n = 100)
constexpr int JAC_MAXPAIRS = 64;  // n <= 128 (a power of two: the phase loops index pairs by tid & 63)

__device__ inline void jacobi_block(int n, double *a, double *v, double *scratch)
{
    __shared__ int s_p[JAC_MAXPAIRS], s_q[JAC_MAXPAIRS];
    __shared__ double s_s[JAC_MAXPAIRS], s_tau[JAC_MAXPAIRS];
    __shared__ int s_rot[JAC_MAXPAIRS];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int m = (n + 1) & ~1, npair = m >> 1;
    for (int idx = tid; idx < n * n; idx += nth) v[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < 60; ++sweep) {
        // converged when every off-diagonal element is exactly zero (the oracle tests sum |a_pq| == 0)
        int nonzero = 0;
        for (int idx = tid; idx < n * n; idx += nth)
            if (idx / n < idx % n && a[idx] != 0.0) nonzero = 1;
        if (!__syncthreads_or(nonzero)) break;
        for (int step = 0; step < m - 1; ++step) {
            // ---- rotation of every pair of this round, from its own 2x2 block
            if (tid < npair) {
                const int k = tid;
                const int i = (k == 0) ? m - 1 : (step + k) % (m - 1);
                const int j = (k == 0) ? step : (step + m - 1 - k) % (m - 1);
                const int p = i < j ? i : j, q = i < j ? j : i;
                int rot = 0;
                double sn = 0.0, tau = 0.0;
                if (q < n) {
                    const double apq = a[p * n + q], app = a[p * n + p], aqq = a[q * n + q];
                    const double g = 100.0 * fabs(apq);
                    if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
                        rot = 2;  // negligible: just clear it
                    } else if (apq != 0.0) {
                        const double h = aqq - app;
                        double t;
                        if (fabs(h) + g == fabs(h)) {
                            t = apq / h;
                        } else {
                            const double theta = 0.5 * h / apq;
                            t = 1.0 / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                            if (theta < 0.0) t = -t;
                        }
                        const double c = 1.0 / sqrt(fma(t, t, 1.0));
                        sn = t * c;
                        tau = sn / (1.0 + c);
                        rot = 1;
                    }
                }
                s_p[k] = p; s_q[k] = q; s_s[k] = sn; s_tau[k] = tau; s_rot[k] = rot;
            }
            __syncthreads();
            // thread (k, r0): pair k = tid mod 64, rows / columns r0, r0 + nth/64, ... (no division in the loops)
            const int k = tid & (JAC_MAXPAIRS - 1), r0 = tid / JAC_MAXPAIRS, rstep = nth / JAC_MAXPAIRS;
            const int rot = k < npair ? s_rot[k] : 0;
            const int p = k < npair ? s_p[k] : 0, q = k < npair ? s_q[k] : 0;
            const double sn = k < npair ? s_s[k] : 0.0, tau = k < npair ? s_tau[k] : 0.0;
            // ---- column phase: A <- A J, V <- V J
            if (rot == 1) {
                for (int r = r0; r < n; r += rstep) {
                    double g = a[r * n + p], h = a[r * n + q];
                    a[r * n + p] = fma(-sn, fma(g, tau, h), g);
                    a[r * n + q] = fma(sn, fma(-h, tau, g), h);
                    g = v[r * n + p]; h = v[r * n + q];
                    v[r * n + p] = fma(-sn, fma(g, tau, h), g);
                    v[r * n + q] = fma(sn, fma(-h, tau, g), h);
                }
            }
            __syncthreads();
            // ---- row phase: A <- J^T A; the rotated (or negligible) pair's off-diagonal is zero by construction
            if (rot == 1) {
                for (int c = r0; c < n; c += rstep) {
                    const double g = a[p * n + c], h = a[q * n + c];
                    a[p * n + c] = (c == q) ? 0.0 : fma(-sn, fma(g, tau, h), g);
                    a[q * n + c] = (c == p) ? 0.0 : fma(sn, fma(-h, tau, g), h);
                }
            } else if (rot == 2 && r0 == 0) {
                a[p * n + q] = 0.0;
                a[q * n + p] = 0.0;
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

struct FactorArgs {
    int d, ngroups;
    const int *goff, *gidx, *uoff, *soff;
    double *cov, *mu, *m2;      // running state (updated in place)
    const double *batch;        // {n_b, mean_b[d], M2c_b[d*d]} or NULL to factor cov as is
    double n_prev;              // samples already in (mu, m2)
    int reset;                  // it == 0 in ref :781-783: forget the running state first
    double *U, *S, *sqrtS;      // outputs, concatenated per group
    double *work_a, *work_v;    // [dmax*dmax] scratch each (global; used when a block does not fit shared memory)
    int smem_doubles;           // dynamic shared memory of the launch, in doubles
    int *ord;                   // [dmax] scratch
};

// One block: merge the batch into the running moments (ref :785-794), then factor every group
// (ref :797-803): eigenvalues by descending magnitude (ties keep index order), S = |lambda|, each
// eigenvector's largest-magnitude component made positive.
__global__ void __launch_bounds__(JAC_THREADS) adapt_finalize_kernel(const FactorArgs f)
{
    __shared__ double scratch[2];
    const int d = f.d;
    if (f.batch) {
        const double nb = f.batch[0];
        const double *mb = f.batch + 1, *m2b = f.batch + 1 + d;
        const double na = f.reset ? 0.0 : f.n_prev;
        const double ntot = na + nb;
        for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
            const int i = idx / d, j = idx % d;
            const double mai = f.reset ? 0.0 : f.mu[i], maj = f.reset ? 0.0 : f.mu[j];
            const double di = mb[i] - mai, dj = mb[j] - maj;
            const double m2a = f.reset ? 0.0 : f.m2[idx];
            const double m2 = m2a + m2b[idx] + di * dj * (na * nb / ntot);
            f.m2[idx] = m2;
            f.cov[idx] = m2 / (ntot - 1.0);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
            const double ma = f.reset ? 0.0 : f.mu[k];
            f.mu[k] = ma + (mb[k] - ma) * (nb / ntot);
        }
        __syncthreads();
    }
    for (int g = 0; g < f.ngroups; ++g) {
        const int g0 = f.goff[g], n = f.goff[g + 1] - g0;
        const int *gi = f.gidx + g0;
        // the block and its eigenvectors live in shared memory when they fit (n <= 119): the rotation phases are
        // latency bound, and an L2 round trip per element made the d=100 factorisation 18.7 ms
        extern __shared__ double jac_smem[];
        const bool in_smem = 2 * n * n <= f.smem_doubles;
        double *a = in_smem ? jac_smem : f.work_a, *v = in_smem ? jac_smem + n * n : f.work_v;
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x)
            a[idx] = f.cov[gi[idx / n] * d + gi[idx % n]];
        __syncthreads();
        jacobi_block(n, a, v, scratch);
        if (threadIdx.x == 0) {
            int *ord = f.ord;
            for (int i = 0; i < n; ++i) ord[i] = i;
            for (int i = 0; i < n; ++i) {
                int best = i;
                for (int j = i + 1; j < n; ++j)
                    if (fabs(a[ord[j] * n + ord[j]]) > fabs(a[ord[best] * n + ord[best]])) best = j;
                const int tmp = ord[best];
                for (int j = best; j > i; --j) ord[j] = ord[j - 1];
                ord[i] = tmp;
            }
        }
        __syncthreads();
        double *U = f.U + f.uoff[g], *S = f.S + f.soff[g], *sS = f.sqrtS + f.soff[g];
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int src = f.ord[k];
            const double lam = fabs(a[src * n + src]);
            S[k] = lam;
            sS[k] = sqrt(lam);
            int big = 0;
            for (int r = 1; r < n; ++r)
                if (fabs(v[r * n + src]) > fabs(v[big * n + src])) big = r;
            const double sg = (v[big * n + src] < 0.0) ? -1.0 : 1.0;
            for (int r = 0; r < n; ++r) U[r * n + k] = sg * v[r * n + src];
        }
        __syncthreads();
    }
}

// Chan merge of nparts batches {n, mean[d], M2c[d*d]} in part order (the same recurrence, in the same order, as
// distributed.merge_batches on the host), result into out.  One block; thread per matrix entry.
__global__ void __launch_bounds__(256) merge_batches_kernel(const double *parts, int nparts, int d, double *out)
{
    const int len = 1 + d + d * d;
    for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
        const int i = idx / d, j = idx % d;
        double n = 0.0, mi = 0.0, mj = 0.0, m2 = 0.0;
        for (int b = 0; b < nparts; ++b) {
            const double *pb = parts + (size_t)b * len;
            const double nb = pb[0];
            if (nb == 0.0) continue;
            const double di = pb[1 + i] - mi, dj = pb[1 + j] - mj, tot = n + nb;
            m2 = m2 + pb[1 + d + idx] + (di * dj) * (n * nb / tot);
            mi = mi + di * (nb / tot);
            mj = mj + dj * (nb / tot);
            n = tot;
        }
        out[1 + d + idx] = m2;
        if (j == 0) out[1 + i] = mi;
        if (idx == 0) out[0] = n;
    }
}

// sqrtS = sqrt(S) after a host-supplied factor (ptmcmc_set_factor)
__global__ void sqrt_kernel(const double *S, double *sqrtS, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sqrtS[i] = sqrt(S[i]);
}

// DE history append (ref :806-817): the reference shifts the buffer left by covUpdate rows and
// copies the AM buffer into the freed tail; here the history is a ring (head advances by
// covUpdate slots) and the AM ring am[slot][k][w] is transposed into de[slot'][w][k].
// A block moves TWD walkers of one slot: d rows of TWD contiguous doubles in, one run of TWD*d contiguous doubles out
// (consecutive walkers' history rows are adjacent), through a [d][TWD+1] shared-memory tile read conflict-free along k.
// grid: (ceil(W/TWD), min(cu, 65535)); a block strides over the slots (gridDim.y is capped at 65535).
__global__ void __launch_bounds__(256) de_append_kernel(const double *am, double *de, int d, int W, int TWD, long long cu,
                                                        long long burn, long long new_head)
{
    extern __shared__ double de_tile[];  // [d][TWD + 1]
    const int w0 = blockIdx.x * TWD, nw = min(TWD, W - w0), ldt = TWD + 1, tid = threadIdx.x;
    for (long long slot = blockIdx.y; slot < cu; slot += gridDim.y) {
        const long long dst_slot = (new_head + (burn - cu) + slot) % burn;
        const double *src = am + (size_t)slot * d * W + w0;
        // four loads in flight per thread (the element count is a run-time value)
        for (int e0 = tid; e0 < d * TWD; e0 += 4 * 256) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256, k = e / TWD, ww = e % TWD;
                v[u] = (e < d * TWD && ww < nw) ? src[(size_t)k * W + ww] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                if (e < d * TWD) de_tile[(e / TWD) * ldt + e % TWD] = v[u];
            }
        }
        __syncthreads();
        double *dst = de + ((size_t)dst_slot * W + w0) * d;
        for (int e = tid; e < nw * d; e += 256) dst[e] = de_tile[(e % d) * ldt + e / d];
        __syncthreads();
    }
}

}  // namespace ptm
