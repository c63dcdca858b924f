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
        const double *base = tile + (lane >> 2) * GRAM_LDT + (lane & 3);
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

// Cyclic Jacobi eigen-decomposition of the n x n symmetric matrix a (destroyed); v receives the
// eigenvectors.  One thread block; rotations are applied by n threads in parallel, in the same
// (p, q) order and with the same formulas as the CPU oracle's orc_sym_factor.
__device__ inline void jacobi_block(int n, double *a, double *v, double *scratch)
{
    const int r = threadIdx.x;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) v[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (threadIdx.x == 0) {
            double off = 0.0;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) off += fabs(a[p * n + q]);
            scratch[0] = off;
        }
        __syncthreads();
        if (scratch[0] == 0.0) break;
        __syncthreads();
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q], app = a[p * n + p], aqq = a[q * n + q];
                const double g = 100.0 * fabs(apq);
                __syncthreads();  // everyone has read the pivot entries
                if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
                    if (threadIdx.x == 0) { a[p * n + q] = 0.0; a[q * n + p] = 0.0; }
                    __syncthreads();
                    continue;
                }
                if (apq == 0.0) continue;
                const double h = aqq - app;
                double t;
                if (fabs(h) + g == fabs(h)) {
                    t = apq / h;
                } else {
                    const double theta = 0.5 * h / apq;
                    t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double sn = t * c;
                const double tau = sn / (1.0 + c);
                if (r < n) {
                    if (r != p && r != q) {
                        const double arp = a[r * n + p], arq = a[r * n + q];
                        const double nrp = arp - sn * (arq + tau * arp);
                        const double nrq = arq + sn * (arp - tau * arq);
                        a[r * n + p] = nrp; a[p * n + r] = nrp;
                        a[r * n + q] = nrq; a[q * n + r] = nrq;
                    }
                    const double vrp = v[r * n + p], vrq = v[r * n + q];
                    v[r * n + p] = vrp - sn * (vrq + tau * vrp);
                    v[r * n + q] = vrq + sn * (vrp - tau * vrq);
                }
                if (threadIdx.x == 0) {
                    a[p * n + p] = app - t * apq;
                    a[q * n + q] = aqq + t * apq;
                    a[p * n + q] = 0.0;
                    a[q * n + p] = 0.0;
                }
                __syncthreads();
            }
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
    double *work_a, *work_v;    // [dmax*dmax] scratch each
    int *ord;                   // [dmax] scratch
};

// One block: merge the batch into the running moments (ref :785-794), then factor every group
// (ref :797-803): eigenvalues by descending magnitude (ties keep index order), S = |lambda|, each
// eigenvector's largest-magnitude component made positive.
__global__ void __launch_bounds__(128) adapt_finalize_kernel(const FactorArgs f)
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
        double *a = f.work_a, *v = f.work_v;
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

// sqrtS = sqrt(S) after a host-supplied factor (ptmcmc_set_factor)
__global__ void sqrt_kernel(const double *S, double *sqrtS, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sqrtS[i] = sqrt(S[i]);
}

// DE history append (ref :806-817): the reference shifts the buffer left by covUpdate rows and
// copies the AM buffer into the freed tail; here the history is a ring (head advances by
// covUpdate slots) and the AM ring am[slot][k][w] is transposed into de[slot'][w][k].
__global__ void __launch_bounds__(256) de_append_kernel(const double *am, double *de, int d, int W, long long cu,
                                                        long long burn, long long new_head)
{
    __shared__ double tile[32][33];
    // grid: (ceil(W/32), ceil(d/32), cu)
    const long long slot = blockIdx.z;
    const long long dst_slot = (new_head + (burn - cu) + slot) % burn;
    const int w0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int kk = ty; kk < 32; kk += 8) {
        const int k = k0 + kk, w = w0 + tx;
        tile[kk][tx] = (k < d && w < W) ? am[((size_t)slot * d + k) * W + w] : 0.0;
    }
    __syncthreads();
    for (int ww = ty; ww < 32; ww += 8) {
        const int w = w0 + ww, k = k0 + tx;
        if (w < W && k < d) de[((size_t)dst_slot * W + w) * d + k] = tile[tx][ww];
    }
}

}  // namespace ptm
