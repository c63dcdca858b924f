// Host side of the B200 PT-MCMC engine and its C ABI (include/ptmcmc_b200.h).
//
// Owns the device-resident state of W walkers x T temperatures and sequences the kernels of one
// reference iteration (ref PTMCMCSampler.py PTMCMCOneStep :530-629): covariance update -> DE
// history update -> fused MH segment -> swap -> buffers/record.  Everything is enqueued on one
// CUDA stream; nothing in ptmcmc_run synchronises with the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ptmcmc_b200.h"
#include "adapt_kernels.cuh"
#include "launch.h"
#include "mh_kernels.cuh"
#include "params.h"
#include "swap_kernels.cuh"
#include "user_target.h"

using namespace ptm;

static_assert(PTMCMC_MAX_CYCLE == MAX_CYCLE, "cycle capacity mismatch");

namespace {

// Mailboxes and the mappings of the neighbours' mailboxes outlive the engines: a new engine of the same geometry (the
// next PTSampler.sample() call) takes the mailbox of the last one together with its sequence number -- every shard has
// made the same number of swaps, so the shards agree on it without talking -- and the IPC mappings are opened once.
struct Mailbox {
    int device;
    size_t bytes;
    double *box;
    unsigned *ctr;
    int *err;
    unsigned long long seq;
    bool in_use;
};
std::mutex g_p2p_mu;
std::vector<Mailbox> g_mailboxes;
std::map<std::string, void *> g_ipc_maps;  // 64-byte handle -> mapping in this process

thread_local std::string g_create_error;

// host layout [T][W][d] <-> device layout [T][d][W]; one block per (walker tile, rung)
__global__ void __launch_bounds__(256) to_device_layout_kernel(const double *src, double *dst, int d, int W)
{
    const int t = blockIdx.y;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)W * d;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(idx / d), k = (int)(idx % d);
        dst[((size_t)t * d + k) * W + w] = src[((size_t)t * W + w) * d + k];
    }
}
__global__ void __launch_bounds__(256) to_host_layout_kernel(const double *src, double *dst, int d, int W)
{
    const int t = blockIdx.y;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)W * d;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(idx / d), k = (int)(idx % d);
        dst[((size_t)t * W + w) * d + k] = src[((size_t)t * d + k) * W + w];
    }
}

__global__ void __launch_bounds__(256) normals_kernel(const unsigned long long *w, long long n, double *z0, double *z1)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        word_to_normals((uint64_t)w[i], z0[i], z1[i]);
}

// fp64 roofline probe: 8 independent DFMA chains per thread, 512 threads per SM (the denominators bench.py
// reports against are measured on the device it runs on, not copied from a data sheet)
__global__ void __launch_bounds__(512) fp64_peak_kernel(double *out, int iters, double a0, double b0)
{
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = a0 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b0);
    }
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

struct Engine {
    ptmcmc_config cfg{};
    int d = 0, W = 0, T = 0, ngroups = 0, njumps = 3, ntr = 1;
    bool identity_group = true, has_state = false, de_in_cycle = false;
    std::vector<double> ladder, mh_temp;
    std::vector<int> goff, gidx, uoff, soff;
    std::vector<int> cyc_jump, cyc_w;
    long long iter = 0, rows = 0, rec_base = 0, de_head = 0, swap_proposed = 0, swap_events = 0;
    long long adapt_done_iter = -1;  // boundary iteration whose covariance update already ran
    long long de_done_iter = -1;     // boundary iteration whose DE-history update already ran
    // ladder sharding
    int Tg = 0;
    bool sharded = false, pending_swap = false, swept = false;
    int *d_carry_code = nullptr;
    double *d_carry_L = nullptr;
    const double *carry_in = nullptr;
    long long nsamp = 0;
    bool pending_propose = false;
    std::string err;
    cudaStream_t stream = nullptr;
    // device memory
    double *x[2] = {nullptr, nullptr}, *lnl[2] = {nullptr, nullptr}, *lp[2] = {nullptr, nullptr};
    int cur = 0;
    double *d_ladder = nullptr, *d_mh_temp = nullptr;
    double *d_cov = nullptr, *d_mu = nullptr, *d_m2 = nullptr, *d_U = nullptr, *d_S = nullptr, *d_sqrtS = nullptr;
    int *d_goff = nullptr, *d_gidx = nullptr, *d_uoff = nullptr, *d_soff = nullptr, *d_ord = nullptr;
    double *d_work_a = nullptr, *d_work_v = nullptr;
    double *d_am = nullptr, *d_de = nullptr;
    double *d_gmu = nullptr, *d_gP = nullptr, *d_plo = nullptr, *d_phi = nullptr;
    double g_offset = 0.0, p_inside = 0.0;
    int p_inclusive = 1;
    double *d_rec_x = nullptr, *d_rec_lnl = nullptr, *d_rec_lnp = nullptr;
    unsigned long long *d_prop = nullptr, *d_acc = nullptr, *d_swap_acc = nullptr;
    unsigned char *d_trace = nullptr;
    short *d_swapmaps = nullptr;
    int *d_map = nullptr;
    // neighbour exchange of the sharded swap through peer memory (ptmcmc_p2p_*): my mailbox holds, per parity of the
    // swap's sequence number, the carry from the hotter shard and the top rung of the colder one, then 8 flags
    double *p2p_box = nullptr, *p2p_above = nullptr, *p2p_below = nullptr;
    bool p2p_above_ipc = false, p2p_below_ipc = false, p2p_on = false;  // (the IPC mappings are process-wide, never closed)
    unsigned long long p2p_seq = 0;
    unsigned *d_p2p_ctr = nullptr;
    int *d_p2p_err = nullptr;
    double *d_swap_prep = nullptr;  // [4][T][W]: state-independent terms of a swap sweep (swap_prep_kernel)
    double *d_part2 = nullptr, *d_batch = nullptr, *d_gram = nullptr;
    double *d_stage = nullptr;  // [T][W][d] staging in the host layout
    int mom_blocks = 0, gram_kp = 0, jac_smem_doubles = 0;
    // host-callback path staging; h_* are engine-owned page-locked host buffers (ptmcmc_callback_buffers)
    double *h_q = nullptr, *h_qxy = nullptr, *h_lnl = nullptr, *h_lp = nullptr, *h_x = nullptr;
    int *h_jump = nullptr;
    double *d_q = nullptr, *d_qxy = nullptr, *d_lnl_new = nullptr, *d_lp_new = nullptr;
    int *d_jump = nullptr;
    unsigned *d_wordpos = nullptr;
    // record sink: rows stream to caller-owned page-locked arrays on a copy stream as segments complete
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rows = nullptr, ev_copied = nullptr, ev_snap[2] = {nullptr, nullptr};
    double *sink_x = nullptr, *sink_lnl = nullptr, *sink_lnp = nullptr;
    long long sink_cap = 0, copied = 0;
    long long *d_snap[2] = {nullptr, nullptr};  // device staging of ptmcmc_snapshot, one per slot
    // timing
    ptmcmc_timing tm{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    int mh_variant = 0;  // 0: default choice, 2: generic kernel, 3: tensor-core kernel, 6: sorted kernel for any ndim <= 32
    uint32_t rk[20] = {};  // Philox round keys of the seed
    UserModule *user = nullptr;  // run-time compiled kernels of a user target (NVRTC), owned by the module cache
    double *d_user_par = nullptr;
    SortedGeom sorted{};   // launch geometry of the sorted shared-memory kernel
    SortedHostTables sorted_tb{};  // its static tables (Gaussian form, mean, prior box)
    // tensor-core (DMMA) kernel: fragment-order matrices and launch geometry
    bool mma_ok = false;
    MmaGeom mma{};
    double *d_Uf = nullptr, *d_Pf = nullptr, *d_gPfull = nullptr, *d_Ut = nullptr;
};

int fail(Engine *e, int code, const char *fmt, ...)
{
    char buf[8192];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    else g_create_error = buf;
    return code;
}

#define CUDA_TRY(e, call)                                                                          \
    do {                                                                                           \
        cudaError_t _st = (call);                                                                  \
        if (_st != cudaSuccess)                                                                    \
            return fail(e, PTMCMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// Device memory comes from the stream-ordered allocator with an unbounded release threshold: building
// and destroying engines repeatedly (one per PTSampler.sample call) then reuses the pooled blocks
// instead of paying cudaMalloc / cudaFree for gigabytes each time.
thread_local cudaStream_t g_alloc_stream = nullptr;

void keep_pool_cached(int device)
{
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

template <typename T>
cudaError_t dalloc(T **p, size_t n)
{
    cudaError_t st = cudaMallocAsync((void **)p, sizeof(T) * (n ? n : 1), g_alloc_stream);
    if (st == cudaSuccess) st = cudaMemsetAsync(*p, 0, sizeof(T) * (n ? n : 1), g_alloc_stream);
    return st;
}

DevParams make_params(const Engine *e)
{
    DevParams p{};
    p.d = e->d; p.W = e->W; p.T = e->T;
    p.walker_offset = e->cfg.walker_offset; p.temp_offset = e->cfg.temp_offset;
    p.ngroups = e->ngroups; p.identity_group = e->identity_group; p.njumps = e->njumps;
    p.seed = e->cfg.seed;
    memcpy(p.rk, e->rk, sizeof p.rk);
    p.x = e->x[e->cur]; p.lnl = e->lnl[e->cur]; p.lp = e->lp[e->cur];
    p.mh_temp = e->d_mh_temp; p.ladder = e->d_ladder;
    p.U = e->d_U; p.sqrtS = e->d_sqrtS;
    p.goff = e->d_goff; p.gidx = e->d_gidx; p.uoff = e->d_uoff; p.soff = e->d_soff;
    p.ncycle = (int)e->cyc_jump.size();
    int cum = 0;
    for (int i = 0; i < p.ncycle; ++i) {
        cum += e->cyc_w[i];
        p.cyc_jump[i] = e->cyc_jump[i];
        p.cyc_cum[i] = cum;
    }
    p.total_weight = cum;
    p.am = e->d_am; p.de = e->d_de;
    p.cov_update = e->cfg.cov_update; p.burn = e->cfg.burn; p.de_head = e->de_head;
    p.logl_kind = e->cfg.logl_kind; p.logp_kind = e->cfg.logp_kind; p.p_inclusive = e->p_inclusive;
    p.g_mu = e->d_gmu; p.g_P = e->d_gP; p.g_offset = e->g_offset; p.p_inside = e->p_inside;
    p.p_lo = e->d_plo; p.p_hi = e->d_phi;
    p.user_par = e->d_user_par; p.n_logl_par = e->cfg.n_logl_user_params; p.n_logp_par = e->cfg.n_logp_user_params;
    p.rec_x = e->d_rec_x; p.rec_lnl = e->d_rec_lnl; p.rec_lnp = e->d_rec_lnp;
    p.rec_base = e->rec_base; p.rec_cap = e->cfg.record_rows; p.thin = e->cfg.thin; p.ntr = e->ntr;
    p.prop = e->d_prop; p.acc = e->d_acc; p.swap_acc = e->d_swap_acc;
    p.trace = e->d_trace; p.trace_cap = e->cfg.trace ? e->cfg.trace_iters : 0;
    return p;
}

struct LaunchTimer {
    Engine *e;
    int cls;
    LaunchTimer(Engine *e_, int cls_, int nlaunch = 1) : e(e_), cls(cls_)
    {
        e->tm.launches[cls] += nlaunch;
        if (e->cfg.timing) cudaEventRecord(e->ev0, e->stream);
    }
    ~LaunchTimer()
    {
        if (e->cfg.timing) {
            cudaEventRecord(e->ev1, e->stream);
            cudaEventSynchronize(e->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e->ev0, e->ev1);
            e->tm.ms[cls] += ms;
        }
    }
};

// every entry point makes the engine's device current (one host thread may drive engines on several devices)
Engine *engine_of(const ptmcmc_engine *h)
{
    Engine *e = (Engine *)h;
    if (e) cudaSetDevice(e->cfg.device);
    return e;
}

int chain_blocks(const Engine *e) { return (int)(((long long)e->T * e->W + MH_THREADS - 1) / MH_THREADS); }
int chain_blocks256(const Engine *e) { return (int)(((long long)e->T * e->W + 255) / 256); }

// the specialised kernels know the three reference proposals; a cycle with the prior-draw jump runs in the generic one
bool fast_reg_path(const Engine *e) { return e->identity_group && e->d <= MAX_REG_DIM && e->njumps == 3 && !e->user; }

cudaError_t build_u_frags(Engine *e)
{
    if (!e->mma_ok) return cudaSuccess;
    cudaError_t st = launch_frag_build(e->d_U, e->d, e->mma.nt, 1, 0, e->d_Uf, e->stream);
    if (st == cudaSuccess) st = launch_transpose(e->d_U, e->d, e->d_Ut, e->stream);
    e->tm.launches[PTMCMC_K_ADAPT] += 2;
    return st;
}

bool use_mma(const Engine *e)
{
    if (!e->mma_ok || e->njumps != 3) return false;
    if (e->mh_variant == 3) return true;
    // measured on B200 (8192 x 16 chains): ndim 24: sorted 3.9e9 vs tensor-core 3.1e9; ndim 32: 1.9e9 vs 2.2e9;
    // beyond 32 the alternative is the local-memory kernel
    return e->mh_variant == 0 && e->d > 24;
}

cudaError_t launch_mh(Engine *e, long long it0, long long it1, bool tail)
{
    DevParams p = make_params(e);
    p.it0 = it0; p.it1 = it1; p.tail = tail ? 1 : 0;
    LaunchTimer lt(e, PTMCMC_K_MH);
    e->tm.chain_steps += (it1 - it0 + 1) * (long long)e->T * e->W;
    if (use_mma(e)) return launch_mma(p, e->mma, e->d_Uf, e->d_Pf, e->d_Ut, e->cfg.device, e->stream);
    if (fast_reg_path(e) && e->mh_variant != 2) return launch_sorted(p, e->sorted_tb, e->sorted, e->cfg.device, e->stream);
    if (e->user) {
        void *args[] = {&p};
        return user_launch(e->user->mh, chain_blocks(e), MH_THREADS, e->stream, args);
    }
    mh_generic_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p);
    return cudaGetLastError();
}

// ref :631-697 + :627 for the iteration `it` that has just been stepped
cudaError_t launch_swap(Engine *e, long long it)
{
    DevParams p = make_params(e);
    LaunchTimer lt(e, PTMCMC_K_SWAP, 3);
    short *tr = nullptr;
    if (e->d_swapmaps && e->swap_events < e->cfg.trace_iters) tr = e->d_swapmaps + (size_t)e->swap_events * e->W * e->T;
    swap_prep_kernel<<<chain_blocks256(e), 256, 0, e->stream>>>(p, it, e->Tg, 0.0, 0, e->d_swap_prep);
    swap_decide_kernel<<<(e->W + 127) / 128, 128, 0, e->stream>>>(p, e->d_swap_prep, e->d_map, tr);
    const int nxt = e->cur ^ 1;
    swap_apply_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, it, e->d_map, e->x[nxt], e->lnl[nxt], e->lp[nxt]);
    e->cur = nxt;
    e->swap_proposed++;
    e->swap_events++;
    return cudaGetLastError();
}

cudaError_t launch_factor(Engine *e, const double *batch, double n_prev, int reset)
{
    FactorArgs f{};
    f.d = e->d; f.ngroups = e->ngroups;
    f.goff = e->d_goff; f.gidx = e->d_gidx; f.uoff = e->d_uoff; f.soff = e->d_soff;
    f.cov = e->d_cov; f.mu = e->d_mu; f.m2 = e->d_m2;
    f.batch = batch; f.n_prev = n_prev; f.reset = reset;
    f.U = e->d_U; f.S = e->d_S; f.sqrtS = e->d_sqrtS;
    f.work_a = e->d_work_a; f.work_v = e->d_work_v; f.ord = e->d_ord;
    f.smem_doubles = e->jac_smem_doubles;
    adapt_finalize_kernel<<<1, JAC_THREADS, sizeof(double) * e->jac_smem_doubles, e->stream>>>(f);
    cudaError_t st = cudaGetLastError();
    return st != cudaSuccess ? st : build_u_frags(e);
}

// local batch moments of the AM ring into d_batch = {n, mean, M2c}
cudaError_t launch_batch_moments(Engine *e)
{
    const int d = e->d, W = e->W;
    const long long cu = e->cfg.cov_update;
    const double n = (double)cu * (double)W;
    (void)n;
    const int KP = e->gram_kp;
    const size_t smem = sizeof(double) * KP * GRAM_LDT;
    if (KP <= 24)
        moments_gram_kernel<1><<<e->mom_blocks, GRAM_THREADS, smem, e->stream>>>(e->d_am, d, W, cu, e->d_mu, KP, e->d_part2);
    else if (KP <= 56)
        moments_gram_kernel<4><<<e->mom_blocks, GRAM_THREADS, smem, e->stream>>>(e->d_am, d, W, cu, e->d_mu, KP, e->d_part2);
    else if (KP <= 104)
        moments_gram_kernel<12><<<e->mom_blocks, GRAM_THREADS, smem, e->stream>>>(e->d_am, d, W, cu, e->d_mu, KP, e->d_part2);
    else
        moments_gram_kernel<GRAM_MAXT><<<e->mom_blocks, GRAM_THREADS, smem, e->stream>>>(e->d_am, d, W, cu, e->d_mu, KP,
                                                                                         e->d_part2);
    moments_gram_sum_kernel<<<(KP * KP + 7) / 8, 256, 0, e->stream>>>(e->d_part2, e->mom_blocks, KP, e->d_gram);
    moments_gram_batch_kernel<<<(d * d + 127) / 128, 128, 0, e->stream>>>(e->d_gram, KP, d, e->d_mu, e->d_batch);
    return cudaGetLastError();
}

// ref :545-560 at the start of iteration it0 (boundary = it0-1)
cudaError_t cov_update(Engine *e, long long boundary)
{
    LaunchTimer lt(e, PTMCMC_K_ADAPT, 4);
    cudaError_t st = launch_batch_moments(e);
    if (st != cudaSuccess) return st;
    const long long it = boundary - e->cfg.cov_update;  // ref :778
    st = launch_factor(e, e->d_batch, (double)it * (double)e->W, it == 0);
    e->nsamp = boundary * (long long)e->W;
    e->adapt_done_iter = boundary;
    return st;
}

// ref :563-585 at the start of iteration it0
int de_update(Engine *e)
{
    const long long cu = e->cfg.cov_update, burn = e->cfg.burn;
    if (cu > burn)
        return fail(e, PTMCMC_ERR_DE_SHAPE,
                    "could not broadcast AM buffer of %lld rows into a DE buffer of %lld rows (covUpdate > burn)",
                    cu, burn);
    {
        LaunchTimer lt(e, PTMCMC_K_DE);
        const long long new_head = (e->de_head + cu) % burn;
        const int twd = e->d <= 40 ? 128 : 32;  // walkers per block: the [d][twd + 1] tile stays under 48 KB
        dim3 grid((e->W + twd - 1) / twd, (unsigned)std::min<long long>(cu, 65535));
        de_append_kernel<<<grid, 256, sizeof(double) * e->d * (twd + 1), e->stream>>>(e->d_am, e->d_de, e->d, e->W, twd, cu, burn,
                                                                                  new_head);
        e->de_head = new_head;
    }
    if (!e->de_in_cycle && e->cfg.de_weight > 0) {
        e->cyc_jump.push_back(PTMCMC_JUMP_DE);
        e->cyc_w.push_back(e->cfg.de_weight);
        e->de_in_cycle = true;
    }
    cudaError_t st = cudaGetLastError();
    if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "de_append_kernel: %s", cudaGetErrorString(st));
    return 0;
}

// maintenance at the start of iteration it0 (ref :545-585)
int maintenance(Engine *e, long long it0)
{
    const long long b = it0 - 1;
    if (b != 0 && b % e->cfg.cov_update == 0 && e->cfg.temp_offset == 0 && e->adapt_done_iter != b) {
        cudaError_t st = cov_update(e, b);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "covariance update: %s", cudaGetErrorString(st));
    }
    if (b != 0 && b % e->cfg.burn == 0 && e->de_done_iter != b) {
        int rc = de_update(e);
        if (rc) return rc;
        e->de_done_iter = b;
    }
    return 0;
}

long long next_multiple(long long it, long long m) { return ((it + m - 1) / m) * m; }

cudaError_t sink_recycle(Engine *e);

int check_rows(Engine *e, long long end_iter)
{
    const long long last_row = end_iter / e->cfg.thin;
    if (last_row - e->rec_base >= e->cfg.record_rows) {
        if (e->sink_x) {  // rows stream to the host: reuse the window once they have landed
            cudaError_t st = sink_recycle(e);
            if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "record sink: %s", cudaGetErrorString(st));
            if (last_row - e->rec_base < e->cfg.record_rows) return 0;
        }
        return fail(e, PTMCMC_ERR_CAPACITY, "record window holds rows [%lld, %lld); iteration %lld needs row %lld",
                    e->rec_base, e->rec_base + e->cfg.record_rows, end_iter, last_row);
    }
    return 0;
}

// counter summary for the host's write cadence (ref jumpDict :602, :622, naccepted :620, nswap_accepted :691):
// out = prop_sum[nj][T] | acc_sum[nj][T] | prop_w0[nj][T] | acc_w0[nj][T] | swap_sum[T] | swap_w0[T]; one block per (j, t)
__global__ void __launch_bounds__(256) counter_summary_kernel(const unsigned long long *prop, const unsigned long long *acc,
                                                              const unsigned long long *swap_acc, int nj, int T, int W,
                                                              long long *out)
{
    __shared__ unsigned long long sh[3][8];
    const int jt = blockIdx.x, j = jt / T, t = jt % T;
    const size_t base = (size_t)jt * W;
    unsigned long long a = 0, b = 0, c = 0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        a += prop[base + w];
        b += acc[base + w];
        if (j == 0) c += swap_acc[(size_t)t * W + w];
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; sh[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[0][i]; b += sh[1][i]; c += sh[2][i]; }
        const int n = nj * T;
        out[jt] = (long long)a;
        out[n + jt] = (long long)b;
        out[2 * n + jt] = (long long)prop[base];
        out[3 * n + jt] = (long long)acc[base];
        if (j == 0) {
            out[4 * n + t] = (long long)c;
            out[4 * n + T + t] = (long long)swap_acc[(size_t)t * W];
        }
    }
}

// rows whose record is complete in stream order -> the sink's host arrays (absolute row positions)
cudaError_t sink_flush(Engine *e)
{
    if (!e->sink_x || e->rows <= e->copied) return cudaSuccess;
    const long long r0 = std::max(e->copied, e->rec_base), r1 = std::min(e->rows, e->sink_cap);
    e->copied = e->rows;
    if (r1 <= r0) return cudaSuccess;
    cudaError_t st = cudaEventRecord(e->ev_rows, e->stream);
    if (st == cudaSuccess) st = cudaStreamWaitEvent(e->copy_stream, e->ev_rows, 0);
    const size_t per_row = (size_t)e->ntr * e->W, off = (size_t)(r0 - e->rec_base) * per_row, dst = (size_t)r0 * per_row;
    const size_t n = (size_t)(r1 - r0) * per_row;
    if (st == cudaSuccess)
        st = cudaMemcpyAsync(e->sink_x + dst * e->d, e->d_rec_x + off * e->d, sizeof(double) * n * e->d, cudaMemcpyDeviceToHost, e->copy_stream);
    if (st == cudaSuccess) st = cudaMemcpyAsync(e->sink_lnl + dst, e->d_rec_lnl + off, sizeof(double) * n, cudaMemcpyDeviceToHost, e->copy_stream);
    if (st == cudaSuccess) st = cudaMemcpyAsync(e->sink_lnp + dst, e->d_rec_lnp + off, sizeof(double) * n, cudaMemcpyDeviceToHost, e->copy_stream);
    if (st == cudaSuccess) st = cudaEventRecord(e->ev_copied, e->copy_stream);
    return st;
}

// with a sink the window is reused once its rows are on the host: no compaction, no host synchronisation
cudaError_t sink_recycle(Engine *e)
{
    cudaError_t st = sink_flush(e);
    if (st == cudaSuccess) st = cudaStreamWaitEvent(e->stream, e->ev_copied, 0);
    e->rec_base = e->rows;
    return st;
}

}  // namespace

extern "C" {

int32_t ptmcmc_abi_version(void) { return PTMCMC_ABI_VERSION; }

int32_t ptmcmc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *ptmcmc_create_error(void) { return g_create_error.c_str(); }
const char *ptmcmc_last_error(const ptmcmc_engine *h) { return h ? ((const Engine *)h)->err.c_str() : ""; }

static int create_impl(Engine *e, const ptmcmc_config *cfg)
{
    const bool tlog = getenv("PTMCMC_CREATE_TIMING") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now();
    e->cfg = *cfg;
    const int d = e->d = cfg->ndim, W = e->W = cfg->nwalkers, T = e->T = cfg->ntemps;
    if (cfg->abi_version != PTMCMC_ABI_VERSION) return fail(nullptr, PTMCMC_ERR_ARG, "ABI version mismatch");
    if (d < 1 || d > MAX_GENERIC_DIM) return fail(nullptr, PTMCMC_ERR_ARG, "ndim must be in [1, %d]", MAX_GENERIC_DIM);
    if (W < 1 || T < 1) return fail(nullptr, PTMCMC_ERR_ARG, "nwalkers and ntemps must be >= 1");
    if (T > 32767) return fail(nullptr, PTMCMC_ERR_ARG, "ntemps must be < 32768");
    if (cfg->cov_update < 1 || cfg->burn < 1 || cfg->tskip < 1 || cfg->thin < 1)
        return fail(nullptr, PTMCMC_ERR_ARG, "covUpdate, burn, Tskip and thin must be >= 1");
    if (cfg->cov_update > 0x7FFFFFFF || cfg->thin > 0x7FFFFFFF)
        return fail(nullptr, PTMCMC_ERR_ARG, "covUpdate and thin must be < 2^31");
    if (!cfg->ladder || !cfg->cov) return fail(nullptr, PTMCMC_ERR_ARG, "ladder and cov are required");
    if (cfg->ncycle < 1 || cfg->ncycle >= PTMCMC_MAX_CYCLE) return fail(nullptr, PTMCMC_ERR_ARG, "No jump proposals specified!");
    if (cfg->record_rows < 1) return fail(nullptr, PTMCMC_ERR_ARG, "record_rows must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, PTMCMC_ERR_CUDA, "no CUDA device: the PT-MCMC engine has no CPU path");
    CUDA_TRY(nullptr, cudaSetDevice(cfg->device));
    // (cudaGetDeviceProperties costs tens of milliseconds per call; one attribute is all that is needed)
    CUDA_TRY(nullptr, cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    keep_pool_cached(cfg->device);
    g_alloc_stream = e->stream;  // allocations, fills and uploads below are ordered on the engine's stream
    CUDA_TRY(nullptr, cudaEventCreate(&e->ev0));
    CUDA_TRY(nullptr, cudaEventCreate(&e->ev1));
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *ev : {&e->ev_rows, &e->ev_copied, &e->ev_snap[0], &e->ev_snap[1]})
        CUDA_TRY(nullptr, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    e->ladder.assign(cfg->ladder, cfg->ladder + T);
    e->mh_temp.assign(cfg->mh_temp ? cfg->mh_temp : cfg->ladder, (cfg->mh_temp ? cfg->mh_temp : cfg->ladder) + T);
    // groups (ref :129-131)
    if (cfg->ngroups <= 0 || !cfg->group_offsets) {
        e->ngroups = 1;
        e->goff = {0, d};
        e->gidx.resize(d);
        for (int i = 0; i < d; ++i) e->gidx[i] = i;
    } else {
        e->ngroups = cfg->ngroups;
        e->goff.assign(cfg->group_offsets, cfg->group_offsets + cfg->ngroups + 1);
        e->gidx.assign(cfg->group_indices, cfg->group_indices + e->goff[cfg->ngroups]);
        for (int v : e->gidx)
            if (v < 0 || v >= d) return fail(nullptr, PTMCMC_ERR_ARG, "group index %d out of range", v);
    }
    e->identity_group = (e->ngroups == 1 && e->goff[1] == d);
    if (e->identity_group)
        for (int i = 0; i < d; ++i) e->identity_group = e->identity_group && e->gidx[i] == i;
    e->uoff.assign(e->ngroups + 1, 0);
    e->soff.assign(e->ngroups + 1, 0);
    int dmax = 0;
    for (int g = 0; g < e->ngroups; ++g) {
        const int dg = e->goff[g + 1] - e->goff[g];
        if (dg < 1) return fail(nullptr, PTMCMC_ERR_ARG, "empty parameter group");
        e->uoff[g + 1] = e->uoff[g] + dg * dg;
        e->soff[g + 1] = e->soff[g] + dg;
        dmax = dg > dmax ? dg : dmax;
    }
    // cycle (ref :1007-1008); zero weights are silently dropped (ref :1001-1004)
    e->njumps = 3;
    for (int i = 0; i < cfg->ncycle; ++i) {
        if (cfg->cycle_weight[i] <= 0) continue;
        e->cyc_jump.push_back(cfg->cycle_jump[i]);
        e->cyc_w.push_back(cfg->cycle_weight[i]);
        if (cfg->cycle_jump[i] + 1 > e->njumps) e->njumps = cfg->cycle_jump[i] + 1;
        if (cfg->cycle_jump[i] == PTMCMC_JUMP_DE) e->de_in_cycle = true;
        if (cfg->cycle_jump[i] == PTMCMC_JUMP_PRIOR && cfg->logp_kind != PTMCMC_LOGP_UNIFORM)
            return fail(nullptr, PTMCMC_ERR_ARG, "the prior-draw jump needs the uniform box prior");
    }
    if (e->cyc_jump.empty()) return fail(nullptr, PTMCMC_ERR_ARG, "No jump proposals specified!");
    e->ntr = cfg->record_hot ? T : 1;
    e->Tg = cfg->ntemps_global > 0 ? cfg->ntemps_global : T;
    if (cfg->temp_offset < 0 || cfg->temp_offset + T > e->Tg)
        return fail(nullptr, PTMCMC_ERR_ARG, "rungs [%d, %d) do not fit a ladder of %d", cfg->temp_offset,
                    cfg->temp_offset + T, e->Tg);
    e->sharded = e->Tg > T;
    if (e->Tg > 32767) return fail(nullptr, PTMCMC_ERR_ARG, "ntemps_global must be < 32768");
    if (const char *v = getenv("PTMCMC_MH_VARIANT")) e->mh_variant = atoi(v);
    philox_round_keys(cfg->seed, e->rk);
    e->sorted = sorted_geometry(d, (long long)T * W, e->sm_count, -1, 0);
    const size_t C = (size_t)T * W;
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(nullptr, dalloc(&e->x[b], C * d));
        CUDA_TRY(nullptr, dalloc(&e->lnl[b], C));
        CUDA_TRY(nullptr, dalloc(&e->lp[b], C));
    }
    CUDA_TRY(nullptr, dalloc(&e->d_stage, C * d));
    CUDA_TRY(nullptr, dalloc(&e->d_ladder, T));
    CUDA_TRY(nullptr, dalloc(&e->d_mh_temp, T));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_ladder, e->ladder.data(), sizeof(double) * T, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_mh_temp, e->mh_temp.data(), sizeof(double) * T, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, dalloc(&e->d_cov, (size_t)d * d));
    CUDA_TRY(nullptr, dalloc(&e->d_mu, d));
    CUDA_TRY(nullptr, dalloc(&e->d_m2, (size_t)d * d));
    CUDA_TRY(nullptr, dalloc(&e->d_U, e->uoff[e->ngroups]));
    CUDA_TRY(nullptr, dalloc(&e->d_S, e->soff[e->ngroups]));
    CUDA_TRY(nullptr, dalloc(&e->d_sqrtS, e->soff[e->ngroups]));
    CUDA_TRY(nullptr, dalloc(&e->d_goff, e->goff.size()));
    CUDA_TRY(nullptr, dalloc(&e->d_gidx, e->gidx.size()));
    CUDA_TRY(nullptr, dalloc(&e->d_uoff, e->uoff.size()));
    CUDA_TRY(nullptr, dalloc(&e->d_soff, e->soff.size()));
    CUDA_TRY(nullptr, dalloc(&e->d_ord, dmax));
    CUDA_TRY(nullptr, dalloc(&e->d_work_a, (size_t)dmax * dmax));
    CUDA_TRY(nullptr, dalloc(&e->d_work_v, (size_t)dmax * dmax));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_goff, e->goff.data(), sizeof(int) * e->goff.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_gidx, e->gidx.data(), sizeof(int) * e->gidx.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_uoff, e->uoff.data(), sizeof(int) * e->uoff.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_soff, e->soff.data(), sizeof(int) * e->soff.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_cov, cfg->cov, sizeof(double) * d * d, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(nullptr, dalloc(&e->d_am, (size_t)cfg->cov_update * d * W));  // ref :220
    CUDA_TRY(nullptr, dalloc(&e->d_de, (size_t)cfg->burn * W * d));        // ref :221
    // targets
    if (cfg->logl_kind == PTMCMC_LOGL_GAUSSIAN) {
        if (!cfg->logl_params) return fail(nullptr, PTMCMC_ERR_ARG, "Gaussian log-likelihood needs parameters");
        const double *mu = cfg->logl_params, *A = mu + d;
        e->g_offset = A[(size_t)d * d];
        // -0.5 * x^T A x folded to the upper triangle: P_ii = -A_ii/2, P_ij = -(A_ij + A_ji)/2 (j > i)
        std::vector<double> P((size_t)d * d, 0.0);
        for (int i = 0; i < d; ++i) {
            P[(size_t)i * d + i] = -0.5 * A[(size_t)i * d + i];
            for (int j = i + 1; j < d; ++j) P[(size_t)i * d + j] = -0.5 * (A[(size_t)i * d + j] + A[(size_t)j * d + i]);
        }
        CUDA_TRY(nullptr, dalloc(&e->d_gmu, d));
        CUDA_TRY(nullptr, dalloc(&e->d_gP, (size_t)d * d));
        CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_gmu, mu, sizeof(double) * d, cudaMemcpyHostToDevice, e->stream));
        CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_gP, P.data(), sizeof(double) * d * d, cudaMemcpyHostToDevice, e->stream));
    } else if (cfg->logl_kind == PTMCMC_LOGL_CURVED) {
        if (d % 2) return fail(nullptr, PTMCMC_ERR_ARG, "curved log-likelihood needs an even ndim");
    } else if (cfg->logl_kind != PTMCMC_LOGL_ROSENBROCK && cfg->logl_kind != PTMCMC_LOGL_EXTERNAL &&
               cfg->logl_kind != PTMCMC_LOGL_USER) {
        return fail(nullptr, PTMCMC_ERR_ARG, "unknown logl_kind %d", cfg->logl_kind);
    }
    if (cfg->logl_kind == PTMCMC_LOGL_USER || cfg->logp_kind == PTMCMC_LOGP_USER) {
        if (cfg->logl_kind == PTMCMC_LOGL_USER && !(cfg->logl_source && *cfg->logl_source))
            return fail(nullptr, PTMCMC_ERR_ARG, "PTMCMC_LOGL_USER needs logl_source");
        if (cfg->logp_kind == PTMCMC_LOGP_USER && !(cfg->logp_source && *cfg->logp_source))
            return fail(nullptr, PTMCMC_ERR_ARG, "PTMCMC_LOGP_USER needs logp_source");
        if (cfg->n_logl_user_params < 0 || cfg->n_logp_user_params < 0 ||
            (cfg->n_logl_user_params + cfg->n_logp_user_params > 0 && !cfg->user_params))
            return fail(nullptr, PTMCMC_ERR_ARG, "user_params does not match n_logl_user_params + n_logp_user_params");
        std::string err;
        e->user = user_module(cfg->logl_kind == PTMCMC_LOGL_USER ? cfg->logl_source : nullptr,
                              cfg->logp_kind == PTMCMC_LOGP_USER ? cfg->logp_source : nullptr, cfg->device, err);
        if (!e->user) return fail(nullptr, PTMCMC_ERR_ARG, "%s", err.c_str());
        const size_t np = (size_t)cfg->n_logl_user_params + cfg->n_logp_user_params;
        CUDA_TRY(nullptr, dalloc(&e->d_user_par, np));
        if (np) CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_user_par, cfg->user_params, sizeof(double) * np, cudaMemcpyHostToDevice, e->stream));
    }
    if (cfg->logp_kind == PTMCMC_LOGP_UNIFORM) {
        if (!cfg->logp_params) return fail(nullptr, PTMCMC_ERR_ARG, "uniform log-prior needs parameters");
        CUDA_TRY(nullptr, dalloc(&e->d_plo, d));
        CUDA_TRY(nullptr, dalloc(&e->d_phi, d));
        CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_plo, cfg->logp_params, sizeof(double) * d, cudaMemcpyHostToDevice, e->stream));
        CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_phi, cfg->logp_params + d, sizeof(double) * d, cudaMemcpyHostToDevice, e->stream));
        e->p_inside = cfg->logp_params[2 * d];
        e->p_inclusive = cfg->logp_params[2 * d + 1] != 0.0;
    } else if (cfg->logp_kind != PTMCMC_LOGP_FLAT && cfg->logp_kind != PTMCMC_LOGP_EXTERNAL && cfg->logp_kind != PTMCMC_LOGP_USER) {
        return fail(nullptr, PTMCMC_ERR_ARG, "unknown logp_kind %d", cfg->logp_kind);
    }
    if (d <= MAX_REG_DIM) {  // static tables of the sorted kernel
        SortedHostTables &tb = e->sorted_tb;
        tb.d = d;
        for (int i = 0; i < d; ++i) {
            tb.mu[i] = cfg->logl_kind == PTMCMC_LOGL_GAUSSIAN ? cfg->logl_params[i] : 0.0;
            tb.lo[i] = -std::numeric_limits<double>::infinity();
            tb.hi[i] = std::numeric_limits<double>::infinity();
            if (cfg->logp_kind == PTMCMC_LOGP_UNIFORM) {
                // the open interval of an exclusive prior is the closed interval of the neighbouring doubles
                const double lo = cfg->logp_params[i], hi = cfg->logp_params[d + i];
                tb.lo[i] = e->p_inclusive ? lo : std::nextafter(lo, std::numeric_limits<double>::infinity());
                tb.hi[i] = e->p_inclusive ? hi : std::nextafter(hi, -std::numeric_limits<double>::infinity());
            }
            for (int j = 0; j < d; ++j) tb.P[i * d + j] = 0.0;
        }
        if (cfg->logl_kind == PTMCMC_LOGL_GAUSSIAN) {
            const double *A = cfg->logl_params + d;
            for (int i = 0; i < d; ++i) {
                tb.P[i * d + i] = -0.5 * A[(size_t)i * d + i];
                for (int j = i + 1; j < d; ++j) tb.P[i * d + j] = -0.5 * (A[(size_t)i * d + j] + A[(size_t)j * d + i]);
            }
        }
    }
    // record window, counters, trace
    const size_t rrows = (size_t)cfg->record_rows * e->ntr * W;
    CUDA_TRY(nullptr, dalloc(&e->d_rec_x, rrows * d));
    CUDA_TRY(nullptr, dalloc(&e->d_rec_lnl, rrows));
    CUDA_TRY(nullptr, dalloc(&e->d_rec_lnp, rrows));
    CUDA_TRY(nullptr, dalloc(&e->d_prop, C * e->njumps));
    CUDA_TRY(nullptr, dalloc(&e->d_acc, C * e->njumps));
    CUDA_TRY(nullptr, dalloc(&e->d_swap_acc, C));
    CUDA_TRY(nullptr, dalloc(&e->d_map, C));
    CUDA_TRY(nullptr, dalloc(&e->d_swap_prep, 4 * C));
    if (e->sharded) {
        CUDA_TRY(nullptr, dalloc(&e->d_carry_code, W));
        CUDA_TRY(nullptr, dalloc(&e->d_carry_L, W));
    }
    if (cfg->trace) {
        CUDA_TRY(nullptr, dalloc(&e->d_trace, (size_t)cfg->trace_iters * C));
        CUDA_TRY(nullptr, dalloc(&e->d_swapmaps, (size_t)cfg->trace_iters * C));
    }
    e->gram_kp = ((d + 1) + 7) / 8 * 8;  // ndim + the column of ones, padded to the 8x8 tile
    e->mom_blocks = (e->gram_kp <= 56 ? 8 : e->gram_kp <= 104 ? 3 : 2) * e->sm_count;  // small tiles: many blocks per SM hide the staging latency
    if (e->gram_kp <= 56) {
        // persistent blocks: exactly the resident number (a partial second wave would run at a fraction of the occupancy)
        int nb = 0;
        const size_t smem = sizeof(double) * e->gram_kp * GRAM_LDT;
        cudaError_t st = e->gram_kp <= 24 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, moments_gram_kernel<1>, GRAM_THREADS, smem)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, moments_gram_kernel<4>, GRAM_THREADS, smem);
        if (st == cudaSuccess && nb > 0) e->mom_blocks = nb * e->sm_count;
    }
    CUDA_TRY(nullptr, dalloc(&e->d_part2, (size_t)e->mom_blocks * e->gram_kp * e->gram_kp));
    CUDA_TRY(nullptr, dalloc(&e->d_gram, (size_t)e->gram_kp * e->gram_kp));
    CUDA_TRY(nullptr, dalloc(&e->d_batch, (size_t)1 + d + (size_t)d * d));
    {
        const size_t smem = sizeof(double) * e->gram_kp * GRAM_LDT;
        if (smem > 48 * 1024)
        {
            CUDA_TRY(nullptr, cudaFuncSetAttribute(moments_gram_kernel<GRAM_MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(nullptr, cudaFuncSetAttribute(moments_gram_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    // the factorisation keeps the largest group's block and eigenvectors in shared memory when they fit
    {
        const size_t want = sizeof(double) * 2 * (size_t)dmax * dmax;
        if (want + 4096 <= 227 * 1024) {  // the kernel's static shared memory (rotation tables) comes on top
            e->jac_smem_doubles = 2 * dmax * dmax;
            if (want > 48 * 1024) {
                static size_t jac_attr[64] = {};
                if (jac_attr[cfg->device & 63] < want) {
                    CUDA_TRY(nullptr, cudaFuncSetAttribute(adapt_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
                    jac_attr[cfg->device & 63] = want;
                }
            }
        }
    }
    // tensor-core path: dense Gaussian target, one identity group, box or flat prior
    e->mma.nt = mma_pick_nt(d);
    e->mma_ok = cfg->logl_kind == PTMCMC_LOGL_GAUSSIAN && e->identity_group && e->mma.nt > 0 &&
                (cfg->logp_kind == PTMCMC_LOGP_UNIFORM || cfg->logp_kind == PTMCMC_LOGP_FLAT);
    if (e->mma_ok) {
        const double *A = cfg->logl_params + d;
        std::vector<double> Pn((size_t)d * d), Lc((size_t)d * d, 0.0);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) Pn[(size_t)i * d + j] = -0.25 * (A[(size_t)i * d + j] + A[(size_t)j * d + i]);
        // Cholesky of sym(icov) / 2 = Lc Lc^T: the quadratic form becomes a sum of squares and the
        // strictly-upper 8x8 tiles of the operand vanish (46 % fewer DMMAs at d=100).  A matrix that is not
        // positive definite keeps the symmetric form.
        bool spd = getenv("PTMCMC_MMA_FULL") == nullptr;
        for (int j = 0; j < d && spd; ++j) {
            double s = -Pn[(size_t)j * d + j];
            for (int k = 0; k < j; ++k) s -= Lc[(size_t)j * d + k] * Lc[(size_t)j * d + k];
            if (!(s > 0.0)) { spd = false; break; }
            const double ljj = std::sqrt(s);
            Lc[(size_t)j * d + j] = ljj;
            for (int i = j + 1; i < d; ++i) {
                double v = -Pn[(size_t)i * d + j];
                for (int k = 0; k < j; ++k) v -= Lc[(size_t)i * d + k] * Lc[(size_t)j * d + k];
                Lc[(size_t)i * d + j] = v / ljj;
            }
        }
        e->mma.tri = spd;
        if (spd) Pn = Lc;
        mma_geometry(e->mma, 0);
        e->mma_ok = e->mma.nt > 0;
        if (e->mma_ok) {
            const size_t nf = (size_t)e->mma.nt * e->mma.nt * 64;
            CUDA_TRY(nullptr, dalloc(&e->d_Uf, nf));
            CUDA_TRY(nullptr, dalloc(&e->d_Pf, nf));
            CUDA_TRY(nullptr, dalloc(&e->d_gPfull, (size_t)d * d));
            CUDA_TRY(nullptr, dalloc(&e->d_Ut, (size_t)d * d));
            // on the engine's stream: the buffers are stream-ordered allocations (a pageable source is staged before the call returns)
            CUDA_TRY(nullptr, cudaMemcpyAsync(e->d_gPfull, Pn.data(), sizeof(double) * d * d, cudaMemcpyHostToDevice, e->stream));
            CUDA_TRY(nullptr, launch_frag_build(e->d_gPfull, d, e->mma.nt, 0, mma_pf_tiles(e->mma) != e->mma.nt * e->mma.nt,
                                                e->d_Pf, e->stream));
        }
    }
    const double t_alloc = now();
    // initial factor (ref :138-145)
    cudaError_t st = launch_factor(e, nullptr, 0.0, 0);
    if (st != cudaSuccess) return fail(nullptr, PTMCMC_ERR_CUDA, "initial factorisation: %s", cudaGetErrorString(st));
    e->tm.launches[PTMCMC_K_ADAPT] += 1;
    const double t_launch = now();
    CUDA_TRY(nullptr, cudaStreamSynchronize(e->stream));
    if (tlog)
        fprintf(stderr, "ptmcmc_create: setup+alloc+upload %.2f ms, factor launch %.2f ms, sync %.2f ms\n", t_alloc - t_begin,
                t_launch - t_alloc, now() - t_launch);
    return 0;
}

ptmcmc_engine *ptmcmc_create(const ptmcmc_config *cfg)
{
    g_create_error.clear();
    if (!cfg) {
        fail(nullptr, PTMCMC_ERR_ARG, "null config");
        return nullptr;
    }
    Engine *e = new Engine();
    if (create_impl(e, cfg) != 0) {
        ptmcmc_destroy((ptmcmc_engine *)e);
        return nullptr;
    }
    return (ptmcmc_engine *)e;
}

void ptmcmc_destroy(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return;
    if (e->stream) cudaStreamSynchronize(e->stream);
    void *ptrs[] = {e->x[0], e->x[1], e->lnl[0], e->lnl[1], e->lp[0], e->lp[1], e->d_ladder, e->d_mh_temp,
                    e->d_cov, e->d_mu, e->d_m2, e->d_U, e->d_S, e->d_sqrtS, e->d_goff, e->d_gidx, e->d_uoff,
                    e->d_soff, e->d_ord, e->d_work_a, e->d_work_v, e->d_am, e->d_de, e->d_gmu, e->d_gP,
                    e->d_plo, e->d_phi, e->d_rec_x, e->d_rec_lnl, e->d_rec_lnp, e->d_prop, e->d_acc,
                    e->d_swap_acc, e->d_trace, e->d_swapmaps, e->d_map, e->d_swap_prep, e->d_part2, e->d_batch, e->d_gram,
                    e->d_q, e->d_qxy, e->d_lnl_new, e->d_lp_new, e->d_jump, e->d_wordpos, e->d_stage,
                    e->d_carry_code, e->d_carry_L, e->d_Uf, e->d_Pf, e->d_gPfull, e->d_Ut, e->d_snap[0], e->d_snap[1], e->d_user_par};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, e->stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    if (e->p2p_box) {  // the mailbox goes back to the process-level cache with its sequence number
        std::lock_guard<std::mutex> lock(g_p2p_mu);
        for (Mailbox &m : g_mailboxes)
            if (m.box == e->p2p_box) {
                m.seq = e->p2p_seq;
                m.in_use = false;
            }
    }
    for (void *hp : {(void *)e->h_q, (void *)e->h_qxy, (void *)e->h_lnl, (void *)e->h_lp, (void *)e->h_x, (void *)e->h_jump})
        if (hp) cudaFreeHost(hp);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (cudaEvent_t ev : {e->ev_rows, e->ev_copied, e->ev_snap[0], e->ev_snap[1]})
        if (ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

// host [T][W][d] -> device [T][d][W]: one DMA into the staging buffer, transposed on the device
static int upload_state(Engine *e, const double *x0)
{
    const int d = e->d, W = e->W, T = e->T;
    const size_t n = (size_t)T * W * d;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_stage, x0, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    dim3 grid((unsigned)std::min<long long>(((long long)W * d + 255) / 256, 4096), (unsigned)T);
    to_device_layout_kernel<<<grid, 256, 0, e->stream>>>(e->d_stage, e->x[e->cur], d, W);
    e->tm.launches[PTMCMC_K_INIT] += 1;
    CUDA_TRY(e, cudaGetLastError());
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));  // x0 may be reused by the caller
    return 0;
}

static int finish_set_state(Engine *e)
{
    e->iter = 0;
    e->rows = 1;
    e->rec_base = 0;
    e->copied = 0;
    e->has_state = true;
    e->pending_swap = e->swept = false;
    DevParams p = make_params(e);
    {
        LaunchTimer lt(e, PTMCMC_K_INIT);
        bookkeep_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, 0);  // ref :491
    }
    CUDA_TRY(e, cudaGetLastError());
    CUDA_TRY(e, sink_flush(e));
    return 0;
}

int32_t ptmcmc_set_state(ptmcmc_engine *h, const double *x0)
{
    Engine *e = engine_of(h);
    if (!e || !x0) return PTMCMC_ERR_ARG;
    if (e->cfg.logl_kind == PTMCMC_LOGL_EXTERNAL || e->cfg.logp_kind == PTMCMC_LOGP_EXTERNAL)
        return fail(e, PTMCMC_ERR_STATE, "external targets: use ptmcmc_set_state_external");
    int rc = upload_state(e, x0);
    if (rc) return rc;
    DevParams p = make_params(e);
    {
        LaunchTimer lt(e, PTMCMC_K_INIT);
        if (e->user) {
            void *args[] = {&p};
            CUDA_TRY(e, user_launch(e->user->init_eval, chain_blocks(e), MH_THREADS, e->stream, args));
        } else {
            init_eval_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p);
        }
    }
    CUDA_TRY(e, cudaGetLastError());
    return finish_set_state(e);
}

int32_t ptmcmc_set_state_external(ptmcmc_engine *h, const double *x0, const double *lnl, const double *lnprior)
{
    Engine *e = engine_of(h);
    if (!e || !x0 || !lnl || !lnprior) return PTMCMC_ERR_ARG;
    int rc = upload_state(e, x0);
    if (rc) return rc;
    const size_t C = (size_t)e->T * e->W;
    CUDA_TRY(e, cudaMemcpyAsync(e->lnl[e->cur], lnl, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->lp[e->cur], lnprior, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));  // lnl / lnprior may be reused by the caller
    {   // targets the device knows are evaluated here; lnlike0 = -inf outside the prior (ref :481-483)
        DevParams p = make_params(e);
        LaunchTimer lt(e, PTMCMC_K_INIT);
        if (e->user) {
            void *args[] = {&p};
            CUDA_TRY(e, user_launch(e->user->init_eval, chain_blocks(e), MH_THREADS, e->stream, args));
        } else {
            init_eval_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p);
        }
        CUDA_TRY(e, cudaGetLastError());
    }
    return finish_set_state(e);
}

int32_t ptmcmc_run(ptmcmc_engine *h, int64_t niter)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->has_state) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_run before ptmcmc_set_state");
    if (e->pending_propose) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_run between propose and accept");
    if (e->cfg.logl_kind == PTMCMC_LOGL_EXTERNAL || e->cfg.logp_kind == PTMCMC_LOGP_EXTERNAL || e->njumps > PTMCMC_JUMP_EXT0)
        return fail(e, PTMCMC_ERR_STATE, "external targets or jumps: drive with ptmcmc_propose / ptmcmc_accept");
    if (niter < 0) return fail(e, PTMCMC_ERR_ARG, "niter < 0");
    if (e->pending_swap) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_run with a sharded swap pending (ptmcmc_swap_*)");
    const long long end = e->iter + niter;
    // the Philox counter carries the iteration in one 32-bit word: beyond that the draws would repeat
    if (end > 0xFFFFFFFFll) return fail(e, PTMCMC_ERR_ARG, "iteration %lld is beyond 2^32 - 1, the range of the random streams", end);
    const long long cu = e->cfg.cov_update, burn = e->cfg.burn, tskip = e->cfg.tskip;
    if (e->sharded && niter > 0 && end > next_multiple(e->iter + 1, tskip))
        return fail(e, PTMCMC_ERR_STATE,
                    "ladder-sharded engine: ptmcmc_run must stop at the swap iteration %lld (asked to reach %lld)",
                    next_multiple(e->iter + 1, tskip), end);
    int rc = e->sink_x ? 0 : check_rows(e, end);
    if (rc) return rc;
    const long long thin = e->cfg.thin, cap = e->cfg.record_rows;
    while (e->iter < end) {
        const long long it0 = e->iter + 1;
        rc = maintenance(e, it0);
        if (rc) return rc;
        long long seg_end = end;
        seg_end = std::min(seg_end, next_multiple(it0, cu));
        seg_end = std::min(seg_end, next_multiple(it0, burn));
        const bool swaps = e->Tg > 1;
        if (swaps) seg_end = std::min(seg_end, next_multiple(it0, tskip));
        if (e->sink_x && seg_end / thin - e->rec_base >= cap) {
            // the window is full: its rows are on their way to the host; reuse it from slot 0 once they have landed
            cudaError_t st = sink_recycle(e);
            if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "record sink: %s", cudaGetErrorString(st));
            seg_end = std::min(seg_end, (e->rec_base + cap) * thin - 1);
        }
        const bool swap_now = swaps && (seg_end % tskip == 0);
        cudaError_t st = launch_mh(e, it0, seg_end, !swap_now);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "MH kernel: %s", cudaGetErrorString(st));
        if (swap_now && e->sharded) {
            e->pending_swap = true;  // the neighbour exchange is driven by the caller
            e->swept = false;
            {
                // everything of the sweep that does not wait for the hotter shard's carry, right behind the MH kernel
                DevParams p = make_params(e);
                LaunchTimer lt(e, PTMCMC_K_SWAP);
                swap_prep_kernel<<<chain_blocks256(e), 256, 0, e->stream>>>(p, seg_end, e->Tg, e->cfg.ladder_above,
                                                                             e->cfg.temp_offset + e->T < e->Tg ? 1 : 0,
                                                                             e->d_swap_prep);
            }
            if (cudaGetLastError() != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "swap_prep_kernel");
        } else if (swap_now) {
            st = launch_swap(e, seg_end);
            if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "swap kernels: %s", cudaGetErrorString(st));
        }
        e->iter = seg_end;
        // a pending sharded swap writes the record of its iteration in ptmcmc_swap_finish
        e->rows = std::max(e->rows, (long long)((e->iter - (e->pending_swap ? 1 : 0)) / e->cfg.thin + 1));
        if (e->sink_x) {
            st = sink_flush(e);
            if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "record sink: %s", cudaGetErrorString(st));
        }
    }
    return 0;
}

// start iteration iter + 1: maintenance, jump and proposal of every chain into the device staging (ref :545-601)
static int propose_core(Engine *e)
{
    if (!e->has_state) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_propose before set_state");
    if (e->sharded) return fail(e, PTMCMC_ERR_STATE, "host callbacks are not available on a ladder-sharded engine");
    if (e->pending_propose) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_propose called twice");
    if (e->iter + 1 > 0xFFFFFFFFll) return fail(e, PTMCMC_ERR_ARG, "iteration beyond 2^32 - 1, the range of the random streams");
    if (!e->d_q) {  // staging for the host round trip, allocated on first use
        g_alloc_stream = e->stream;
        const size_t C0 = (size_t)e->T * e->W;
        CUDA_TRY(e, dalloc(&e->d_q, C0 * e->d));
        CUDA_TRY(e, dalloc(&e->d_qxy, C0));
        CUDA_TRY(e, dalloc(&e->d_lnl_new, C0));
        CUDA_TRY(e, dalloc(&e->d_lp_new, C0));
        CUDA_TRY(e, dalloc(&e->d_jump, C0));
        CUDA_TRY(e, dalloc(&e->d_wordpos, C0));
    }
    const long long it0 = e->iter + 1;
    int rc = check_rows(e, it0);
    if (rc) return rc;
    rc = maintenance(e, it0);
    if (rc) return rc;
    DevParams p = make_params(e);
    p.it0 = p.it1 = it0;
    {
        LaunchTimer lt(e, PTMCMC_K_PROPOSE);
        propose_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, e->d_q, e->d_jump, e->d_wordpos);
    }
    CUDA_TRY(e, cudaGetLastError());
    return 0;
}

// finish the iteration from the device staging: Hastings test, swap, buffers and record (ref :605-627)
static int accept_core(Engine *e)
{
    const size_t C = (size_t)e->T * e->W;
    const long long it = e->iter + 1;
    DevParams p = make_params(e);
    p.it0 = p.it1 = it;
    {
        LaunchTimer lt(e, PTMCMC_K_ACCEPT);
        if (e->user) {
            void *args[] = {&p, &e->d_q, &e->d_qxy, &e->d_lnl_new, &e->d_lp_new, &e->d_jump, &e->d_wordpos};
            CUDA_TRY(e, user_launch(e->user->accept, chain_blocks(e), MH_THREADS, e->stream, args));
        } else {
            accept_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, e->d_q, e->d_qxy, e->d_lnl_new, e->d_lp_new,
                                                                         e->d_jump, e->d_wordpos);
        }
    }
    CUDA_TRY(e, cudaGetLastError());
    e->tm.chain_steps += (long long)C;
    if (e->T > 1 && it % e->cfg.tskip == 0) {
        cudaError_t st = launch_swap(e, it);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "swap kernels: %s", cudaGetErrorString(st));
    } else {
        LaunchTimer lt(e, PTMCMC_K_ACCEPT);
        bookkeep_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, it);
    }
    CUDA_TRY(e, cudaGetLastError());
    e->iter = it;
    e->rows = std::max(e->rows, (long long)(e->iter / e->cfg.thin + 1));
    e->pending_propose = false;
    CUDA_TRY(e, sink_flush(e));
    return 0;
}

int32_t ptmcmc_propose(ptmcmc_engine *h, double *q, int32_t *jump)
{
    Engine *e = engine_of(h);
    if (!e || !q || !jump) return PTMCMC_ERR_ARG;
    int rc = propose_core(e);
    if (rc) return rc;
    const size_t C = (size_t)e->T * e->W;
    CUDA_TRY(e, cudaMemcpyAsync(q, e->d_q, sizeof(double) * C * e->d, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(jump, e->d_jump, sizeof(int) * C, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->pending_propose = true;
    return 0;
}

int32_t ptmcmc_accept(ptmcmc_engine *h, const double *q, const double *qxy, const double *lnl, const double *lnprior)
{
    Engine *e = engine_of(h);
    if (!e || !q || !qxy || !lnl || !lnprior) return PTMCMC_ERR_ARG;
    if (!e->pending_propose) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_accept without ptmcmc_propose");
    const size_t C = (size_t)e->T * e->W;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_q, q, sizeof(double) * C * e->d, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_qxy, qxy, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_lnl_new, lnl, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_lp_new, lnprior, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    int rc = accept_core(e);
    if (rc) return rc;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));  // the host buffers may be reused by the caller
    return 0;
}

int32_t ptmcmc_callback_buffers(ptmcmc_engine *h, double **q, int32_t **jump, double **qxy, double **lnl, double **lnprior,
                                double **x)
{
    Engine *e = engine_of(h);
    if (!e || !q || !jump || !qxy || !lnl || !lnprior || !x) return PTMCMC_ERR_ARG;
    const size_t C = (size_t)e->T * e->W;
    if (!e->h_q) {
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_q, sizeof(double) * C * e->d, cudaHostAllocDefault));
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_x, sizeof(double) * C * e->d, cudaHostAllocDefault));
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_qxy, sizeof(double) * C, cudaHostAllocDefault));
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_lnl, sizeof(double) * C, cudaHostAllocDefault));
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_lp, sizeof(double) * C, cudaHostAllocDefault));
        CUDA_TRY(e, cudaHostAlloc((void **)&e->h_jump, sizeof(int) * C, cudaHostAllocDefault));
        memset(e->h_qxy, 0, sizeof(double) * C);
        memset(e->h_lnl, 0, sizeof(double) * C);
        memset(e->h_lp, 0, sizeof(double) * C);
    }
    *q = e->h_q; *jump = e->h_jump; *qxy = e->h_qxy; *lnl = e->h_lnl; *lnprior = e->h_lp; *x = e->h_x;
    return 0;
}

int32_t ptmcmc_propose_pinned(ptmcmc_engine *h, int32_t want_x)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->h_q) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_propose_pinned before ptmcmc_callback_buffers");
    int rc = propose_core(e);
    if (rc) return rc;
    const size_t C = (size_t)e->T * e->W;
    CUDA_TRY(e, cudaMemcpyAsync(e->h_q, e->d_q, sizeof(double) * C * e->d, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->h_jump, e->d_jump, sizeof(int) * C, cudaMemcpyDeviceToHost, e->stream));
    if (want_x) {  // the current points in the host layout, for custom jumps (ref :1058-1065: func(x, iter, beta))
        dim3 grid((unsigned)std::min<long long>(((long long)e->W * e->d + 255) / 256, 4096), (unsigned)e->T);
        to_host_layout_kernel<<<grid, 256, 0, e->stream>>>(e->x[e->cur], e->d_stage, e->d, e->W);
        e->tm.launches[PTMCMC_K_PROPOSE] += 1;
        CUDA_TRY(e, cudaGetLastError());
        CUDA_TRY(e, cudaMemcpyAsync(e->h_x, e->d_stage, sizeof(double) * C * e->d, cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));  // the one host synchronisation of an iteration
    e->pending_propose = true;
    return 0;
}

int32_t ptmcmc_accept_pinned(ptmcmc_engine *h, int32_t q_modified)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->pending_propose || !e->h_q) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_accept_pinned without ptmcmc_propose_pinned");
    const size_t C = (size_t)e->T * e->W;
    // proposals the host did not touch are still in the device staging: nothing to upload
    if (q_modified) CUDA_TRY(e, cudaMemcpyAsync(e->d_q, e->h_q, sizeof(double) * C * e->d, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_qxy, e->h_qxy, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    if (e->cfg.logl_kind == PTMCMC_LOGL_EXTERNAL)
        CUDA_TRY(e, cudaMemcpyAsync(e->d_lnl_new, e->h_lnl, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    if (e->cfg.logp_kind == PTMCMC_LOGP_EXTERNAL)
        CUDA_TRY(e, cudaMemcpyAsync(e->d_lp_new, e->h_lp, sizeof(double) * C, cudaMemcpyHostToDevice, e->stream));
    // no synchronisation: the copies and kernels are ordered on the engine's stream, and the host does not write the
    // buffers again before the next ptmcmc_propose_pinned has synchronised
    return accept_core(e);
}

int64_t ptmcmc_iteration(const ptmcmc_engine *h) { return h ? ((const Engine *)h)->iter : -1; }

int32_t ptmcmc_sync(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->copy_stream));
    return 0;
}

int32_t ptmcmc_get_state(ptmcmc_engine *h, double *x, double *lnl, double *lnprior, double *lnprob)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    const int d = e->d, W = e->W, T = e->T;
    const size_t C = (size_t)T * W;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (x) {
        dim3 grid((unsigned)std::min<long long>(((long long)W * d + 255) / 256, 4096), (unsigned)T);
        to_host_layout_kernel<<<grid, 256, 0, e->stream>>>(e->x[e->cur], e->d_stage, d, W);
        e->tm.launches[PTMCMC_K_INIT] += 1;
        CUDA_TRY(e, cudaGetLastError());
        CUDA_TRY(e, cudaMemcpyAsync(x, e->d_stage, sizeof(double) * C * d, cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    }
    std::vector<double> l, p_;
    if (lnl || lnprob) {
        l.resize(C);
        CUDA_TRY(e, cudaMemcpy(l.data(), e->lnl[e->cur], sizeof(double) * C, cudaMemcpyDeviceToHost));
        if (lnl) memcpy(lnl, l.data(), sizeof(double) * C);
    }
    if (lnprior || lnprob) {
        p_.resize(C);
        CUDA_TRY(e, cudaMemcpy(p_.data(), e->lp[e->cur], sizeof(double) * C, cudaMemcpyDeviceToHost));
        if (lnprior) memcpy(lnprior, p_.data(), sizeof(double) * C);
    }
    if (lnprob)
        for (int t = 0; t < T; ++t)
            for (int w = 0; w < W; ++w) {
                const size_t c = (size_t)t * W + w;
                lnprob[c] = 1.0 / e->mh_temp[t] * l[c] + p_[c];  // ref :487, :612, :695
            }
    return 0;
}

int64_t ptmcmc_rows(const ptmcmc_engine *h) { return h ? ((const Engine *)h)->rows : -1; }
int64_t ptmcmc_row_base(const ptmcmc_engine *h) { return h ? ((const Engine *)h)->rec_base : -1; }

int32_t ptmcmc_get_chain(ptmcmc_engine *h, int64_t row0, int64_t nrows, double *chain, double *lnl, double *lnprob)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (row0 < e->rec_base || row0 + nrows > e->rows || nrows < 0)
        return fail(e, PTMCMC_ERR_ARG, "rows [%lld, %lld) are not resident (window [%lld, %lld))", (long long)row0,
                    (long long)(row0 + nrows), e->rec_base, e->rows);
    const size_t per_row = (size_t)e->ntr * e->W;
    const size_t off = (size_t)(row0 - e->rec_base) * per_row, n = (size_t)nrows * per_row;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (chain) CUDA_TRY(e, cudaMemcpy(chain, e->d_rec_x + off * e->d, sizeof(double) * n * e->d, cudaMemcpyDeviceToHost));
    if (lnl) CUDA_TRY(e, cudaMemcpy(lnl, e->d_rec_lnl + off, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (lnprob) CUDA_TRY(e, cudaMemcpy(lnprob, e->d_rec_lnp + off, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

int32_t ptmcmc_release_rows(ptmcmc_engine *h, int64_t upto_row)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (upto_row < e->rec_base || upto_row > e->rows) return fail(e, PTMCMC_ERR_ARG, "release_rows out of range");
    const long long keep = e->rows - upto_row;  // rows that stay resident
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (keep > 0) {
        const size_t per_row = (size_t)e->ntr * e->W;
        const size_t src = (size_t)(upto_row - e->rec_base) * per_row, n = (size_t)keep * per_row;
        // staged through a pooled scratch buffer (source and destination overlap), all on the engine's stream
        double *tmp = nullptr;
        g_alloc_stream = e->stream;
        CUDA_TRY(e, cudaMallocAsync((void **)&tmp, sizeof(double) * n * e->d, e->stream));
        struct { double *buf; size_t width; } parts[3] = {{e->d_rec_x, (size_t)e->d}, {e->d_rec_lnl, 1}, {e->d_rec_lnp, 1}};
        for (auto &pt : parts) {
            const size_t bytes = sizeof(double) * n * pt.width;
            CUDA_TRY(e, cudaMemcpyAsync(tmp, pt.buf + src * pt.width, bytes, cudaMemcpyDeviceToDevice, e->stream));
            CUDA_TRY(e, cudaMemcpyAsync(pt.buf, tmp, bytes, cudaMemcpyDeviceToDevice, e->stream));
        }
        CUDA_TRY(e, cudaFreeAsync(tmp, e->stream));
    }
    e->rec_base = upto_row;
    return 0;
}

int32_t ptmcmc_get_adapt(ptmcmc_engine *h, double *cov, double *mu, double *m2, int64_t *nsamp)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    const int d = e->d;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (cov) CUDA_TRY(e, cudaMemcpy(cov, e->d_cov, sizeof(double) * d * d, cudaMemcpyDeviceToHost));
    if (mu) CUDA_TRY(e, cudaMemcpy(mu, e->d_mu, sizeof(double) * d, cudaMemcpyDeviceToHost));
    if (m2) CUDA_TRY(e, cudaMemcpy(m2, e->d_m2, sizeof(double) * d * d, cudaMemcpyDeviceToHost));
    if (nsamp) *nsamp = e->nsamp;
    return 0;
}

int32_t ptmcmc_get_factor(ptmcmc_engine *h, double *U, double *S)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (U) CUDA_TRY(e, cudaMemcpy(U, e->d_U, sizeof(double) * e->uoff[e->ngroups], cudaMemcpyDeviceToHost));
    if (S) CUDA_TRY(e, cudaMemcpy(S, e->d_S, sizeof(double) * e->soff[e->ngroups], cudaMemcpyDeviceToHost));
    return 0;
}

int32_t ptmcmc_set_factor(ptmcmc_engine *h, const double *U, const double *S)
{
    Engine *e = engine_of(h);
    if (!e || !U || !S) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_U, U, sizeof(double) * e->uoff[e->ngroups], cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(e->d_S, S, sizeof(double) * e->soff[e->ngroups], cudaMemcpyHostToDevice, e->stream));
    const int n = e->soff[e->ngroups];
    sqrt_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->d_S, e->d_sqrtS, n);
    CUDA_TRY(e, build_u_frags(e));
    e->tm.launches[PTMCMC_K_ADAPT] += 1;
    CUDA_TRY(e, cudaGetLastError());
    return 0;
}

int32_t ptmcmc_get_buffers(ptmcmc_engine *h, double *am, double *de)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    const int d = e->d, W = e->W;
    const long long cu = e->cfg.cov_update, burn = e->cfg.burn;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (am) {
        std::vector<double> tmp((size_t)cu * d * W);
        CUDA_TRY(e, cudaMemcpy(tmp.data(), e->d_am, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
        for (long long s = 0; s < cu; ++s)
            for (int k = 0; k < d; ++k)
                for (int w = 0; w < W; ++w) am[((size_t)s * W + w) * d + k] = tmp[((size_t)s * d + k) * W + w];
    }
    if (de) {
        // un-rotate the ring so that row 0 is the oldest slot, as in the reference's shifted array
        const size_t row = (size_t)W * d;
        const long long first = burn - e->de_head;
        CUDA_TRY(e, cudaMemcpy(de, e->d_de + (size_t)e->de_head * row, sizeof(double) * (size_t)first * row, cudaMemcpyDeviceToHost));
        if (e->de_head)
            CUDA_TRY(e, cudaMemcpy(de + (size_t)first * row, e->d_de, sizeof(double) * (size_t)e->de_head * row, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int32_t ptmcmc_adapt_begin(ptmcmc_engine *h, double *batch_out)
{
    Engine *e = engine_of(h);
    if (!e || !batch_out) return PTMCMC_ERR_ARG;
    const long long b = e->iter;
    if (b == 0 || b % e->cfg.cov_update != 0 || e->adapt_done_iter == b) return 0;
    {
        LaunchTimer lt(e, PTMCMC_K_ADAPT, 3);
        cudaError_t st = launch_batch_moments(e);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "batch moments: %s", cudaGetErrorString(st));
    }
    const size_t n = (size_t)1 + e->d + (size_t)e->d * e->d;
    CUDA_TRY(e, cudaMemcpyAsync(batch_out, e->d_batch, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return 1;
}

int32_t ptmcmc_adapt_finish(ptmcmc_engine *h, const double *batch_in)
{
    Engine *e = engine_of(h);
    if (!e || !batch_in) return PTMCMC_ERR_ARG;
    const long long b = e->iter;
    if (b == 0 || b % e->cfg.cov_update != 0 || e->adapt_done_iter == b)
        return fail(e, PTMCMC_ERR_STATE, "no covariance update is due at iteration %lld", b);
    const size_t n = (size_t)1 + e->d + (size_t)e->d * e->d;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_batch, batch_in, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    const long long it = b - e->cfg.cov_update;
    const double n_prev = (it == 0) ? 0.0 : (double)e->nsamp;
    {
        LaunchTimer lt(e, PTMCMC_K_ADAPT);
        cudaError_t st = launch_factor(e, e->d_batch, n_prev, it == 0);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "factor: %s", cudaGetErrorString(st));
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->nsamp = (long long)(n_prev + batch_in[0]);
    e->adapt_done_iter = b;
    return 0;
}

int32_t ptmcmc_adapt_begin_dev(ptmcmc_engine *h, void **dev_batch, int64_t *ndoubles)
{
    Engine *e = engine_of(h);
    if (!e || !dev_batch || !ndoubles) return PTMCMC_ERR_ARG;
    *dev_batch = e->d_batch;
    *ndoubles = (int64_t)1 + e->d + (int64_t)e->d * e->d;
    const long long b = e->iter;
    if (b == 0 || b % e->cfg.cov_update != 0 || e->adapt_done_iter == b) return 0;
    LaunchTimer lt(e, PTMCMC_K_ADAPT, 3);
    cudaError_t st = launch_batch_moments(e);
    if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "batch moments: %s", cudaGetErrorString(st));
    return 1;
}

int32_t ptmcmc_adapt_finish_dev(ptmcmc_engine *h, const double *dev_parts, int32_t nparts, int64_t nsamples)
{
    Engine *e = engine_of(h);
    if (!e || !dev_parts || nparts < 1) return PTMCMC_ERR_ARG;
    const long long b = e->iter;
    if (b == 0 || b % e->cfg.cov_update != 0 || e->adapt_done_iter == b)
        return fail(e, PTMCMC_ERR_STATE, "no covariance update is due at iteration %lld", b);
    const long long it = b - e->cfg.cov_update;
    const double n_prev = (it == 0) ? 0.0 : (double)e->nsamp;
    {
        LaunchTimer lt(e, PTMCMC_K_ADAPT, 2);
        merge_batches_kernel<<<1, 256, 0, e->stream>>>(dev_parts, nparts, e->d, e->d_batch);
        cudaError_t st = cudaGetLastError();
        if (st == cudaSuccess) st = launch_factor(e, e->d_batch, n_prev, it == 0);
        if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "pooled factor: %s", cudaGetErrorString(st));
    }
    e->nsamp = (long long)n_prev + nsamples;
    e->adapt_done_iter = b;
    return 0;
}

int32_t ptmcmc_factor_dev(ptmcmc_engine *h, void **dev_U, int64_t *usize, void **dev_S, int64_t *ssize)
{
    Engine *e = engine_of(h);
    if (!e || !dev_U || !dev_S || !usize || !ssize) return PTMCMC_ERR_ARG;
    *dev_U = e->d_U; *usize = e->uoff[e->ngroups];
    *dev_S = e->d_S; *ssize = e->soff[e->ngroups];
    return 0;
}

int32_t ptmcmc_factor_refresh(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    const int n = e->soff[e->ngroups];
    sqrt_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->d_S, e->d_sqrtS, n);
    e->tm.launches[PTMCMC_K_ADAPT] += 1;
    CUDA_TRY(e, cudaGetLastError());
    CUDA_TRY(e, build_u_frags(e));
    return 0;
}

int64_t ptmcmc_swap_msg_doubles(const ptmcmc_engine *h)
{
    const Engine *e = (const Engine *)h;
    return e ? (int64_t)(e->d + 3) * e->W : -1;
}

int32_t ptmcmc_swap_pending(const ptmcmc_engine *h) { return h && ((const Engine *)h)->pending_swap ? 1 : 0; }

int32_t ptmcmc_swap_pack_top(ptmcmc_engine *h, double *dev_msg)
{
    Engine *e = engine_of(h);
    if (!e || !dev_msg) return PTMCMC_ERR_ARG;
    if (!e->sharded || !e->pending_swap) return fail(e, PTMCMC_ERR_STATE, "no sharded swap is pending");
    DevParams p = make_params(e);
    {
        LaunchTimer lt(e, PTMCMC_K_SWAP);
        const long long n = (long long)(e->d + 3) * e->W;
        swap_pack_top_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 2048), 256, 0, e->stream>>>(p, dev_msg, P2PSync{});
    }
    CUDA_TRY(e, cudaGetLastError());
    return 0;
}

int32_t ptmcmc_swap_sweep(ptmcmc_engine *h, const double *dev_carry_in, double *dev_carry_out)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->sharded || !e->pending_swap || e->swept) return fail(e, PTMCMC_ERR_STATE, "no sharded swap sweep is due");
    const bool hottest = e->cfg.temp_offset + e->T == e->Tg, coldest = e->cfg.temp_offset == 0;
    if ((dev_carry_in == nullptr) != hottest || (dev_carry_out == nullptr) != coldest)
        return fail(e, PTMCMC_ERR_ARG, "carry_in is NULL exactly on the hottest shard, carry_out exactly on the coldest");
    DevParams p = make_params(e);
    {
        LaunchTimer lt(e, PTMCMC_K_SWAP);
        swap_sweep_kernel<<<(e->W + 127) / 128, 128, 0, e->stream>>>(p, e->cfg.ladder_above, e->d_swap_prep, dev_carry_in,
                                                                      dev_carry_out, e->d_map, e->d_carry_code,
                                                                      e->d_carry_L, P2PSync{});
    }
    CUDA_TRY(e, cudaGetLastError());
    e->carry_in = dev_carry_in;
    e->swept = true;
    return 0;
}

int32_t ptmcmc_swap_finish(ptmcmc_engine *h, const double *dev_below_top)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->sharded || !e->pending_swap || !e->swept) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_swap_finish before ptmcmc_swap_sweep");
    if ((dev_below_top == nullptr) != (e->cfg.temp_offset == 0))
        return fail(e, PTMCMC_ERR_ARG, "below_top is NULL exactly on the coldest shard");
    DevParams p = make_params(e);
    short *tr = nullptr;
    if (e->d_swapmaps && e->swap_events < e->cfg.trace_iters) tr = e->d_swapmaps + (size_t)e->swap_events * e->W * e->T;
    const int nxt = e->cur ^ 1;
    {
        LaunchTimer lt(e, PTMCMC_K_SWAP);
        swap_finish_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, e->iter, e->Tg, e->cfg.ladder_below, e->d_map,
                                                                          e->d_carry_code, e->d_carry_L, e->carry_in,
                                                                          dev_below_top, e->x[nxt], e->lnl[nxt], e->lp[nxt],
                                                                          tr, P2PSync{});
    }
    CUDA_TRY(e, cudaGetLastError());
    e->cur = nxt;
    e->swap_proposed++;
    e->swap_events++;
    e->pending_swap = false;
    e->swept = false;
    e->carry_in = nullptr;
    e->rows = std::max(e->rows, (long long)(e->iter / e->cfg.thin + 1));
    CUDA_TRY(e, sink_flush(e));
    return 0;
}

// ---- the same three steps with the messages moved by the kernels themselves through peer memory ------------------
namespace {

inline double *p2p_slot(double *box, size_t n, int par, int kind) { return box + (size_t)(par * 2 + kind) * n; }
inline unsigned long long *p2p_flag(double *box, size_t n, int par, int kind)
{
    return reinterpret_cast<unsigned long long *>(box + 4 * n) + (par * 2 + kind);
}
}  // namespace

int32_t ptmcmc_p2p_open(ptmcmc_engine *h, void *ipc_handle_out, void **mailbox_out)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->sharded) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_p2p_open on an engine that is not a ladder shard");
    const size_t n = (size_t)(e->d + 3) * e->W, bytes = 4 * n * sizeof(double) + 64;
    if (!e->p2p_box) {
        std::lock_guard<std::mutex> lock(g_p2p_mu);
        for (Mailbox &m : g_mailboxes) {
            if (!m.in_use && m.device == e->cfg.device && m.bytes == bytes) {
                m.in_use = true;
                e->p2p_box = m.box;
                e->d_p2p_ctr = m.ctr;
                e->d_p2p_err = m.err;
                e->p2p_seq = m.seq;
                cudaMemset(m.err, 0, sizeof(int));
                break;
            }
        }
        if (!e->p2p_box) {
            // cudaMalloc, not the stream-ordered pool: the allocation is exported to the neighbours' processes
            Mailbox m{e->cfg.device, bytes, nullptr, nullptr, nullptr, 0ull, true};
            CUDA_TRY(e, cudaMalloc(&m.box, bytes));
            CUDA_TRY(e, cudaMemset(m.box, 0, bytes));
            CUDA_TRY(e, cudaMalloc(&m.ctr, 4 * sizeof(unsigned)));
            CUDA_TRY(e, cudaMemset(m.ctr, 0, 4 * sizeof(unsigned)));
            CUDA_TRY(e, cudaMalloc(&m.err, sizeof(int)));
            CUDA_TRY(e, cudaMemset(m.err, 0, sizeof(int)));
            g_mailboxes.push_back(m);
            e->p2p_box = m.box;
            e->d_p2p_ctr = m.ctr;
            e->d_p2p_err = m.err;
            e->p2p_seq = 0;
        }
    }
    if (mailbox_out) *mailbox_out = e->p2p_box;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t hd;
        CUDA_TRY(e, cudaIpcGetMemHandle(&hd, e->p2p_box));
        static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(ipc_handle_out, &hd, sizeof hd);
    }
    return 0;
}

int64_t ptmcmc_p2p_seq(const ptmcmc_engine *h) { return h ? (int64_t)((const Engine *)h)->p2p_seq : -1; }

int32_t ptmcmc_p2p_connect(ptmcmc_engine *h, const void *above, const void *below, int32_t ipc, int64_t seq0)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->p2p_box) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_p2p_connect before ptmcmc_p2p_open");
    const bool hottest = e->cfg.temp_offset + e->T == e->Tg, coldest = e->cfg.temp_offset == 0;
    if ((above == nullptr) != hottest || (below == nullptr) != coldest)
        return fail(e, PTMCMC_ERR_ARG, "above is NULL exactly on the hottest shard, below exactly on the coldest");
    auto map = [&](const void *src, double **dst, bool *is_ipc) -> cudaError_t {
        *dst = nullptr;
        *is_ipc = false;
        if (!src) return cudaSuccess;
        if (!ipc) {
            *dst = (double *)src;  // a mailbox of another engine in this process
            return cudaSuccess;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, src, sizeof hd);
        const std::string key((const char *)src, sizeof hd);
        std::lock_guard<std::mutex> lock(g_p2p_mu);
        auto it = g_ipc_maps.find(key);
        if (it != g_ipc_maps.end()) {
            *dst = (double *)it->second;
            return cudaSuccess;
        }
        void *ptr = nullptr;
        cudaError_t st = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (st == cudaSuccess) {
            *dst = (double *)ptr;
            g_ipc_maps[key] = ptr;  // stays mapped for the life of the process
        }
        return st;
    };
    CUDA_TRY(e, map(above, &e->p2p_above, &e->p2p_above_ipc));
    CUDA_TRY(e, map(below, &e->p2p_below, &e->p2p_below_ipc));
    if (seq0 < (int64_t)e->p2p_seq) return fail(e, PTMCMC_ERR_ARG, "ptmcmc_p2p_connect: seq0 is behind this mailbox's sequence number");
    e->p2p_seq = (unsigned long long)seq0;  // the shards' common starting point: flags only ever grow
    e->p2p_on = true;
    return 0;
}

int32_t ptmcmc_swap_p2p(ptmcmc_engine *h, int32_t phase)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->p2p_on) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_swap_p2p before ptmcmc_p2p_connect");
    if (!e->sharded || !e->pending_swap) return fail(e, PTMCMC_ERR_STATE, "no sharded swap is pending");
    const bool hottest = e->cfg.temp_offset + e->T == e->Tg, coldest = e->cfg.temp_offset == 0;
    const size_t n = (size_t)(e->d + 3) * e->W;
    const unsigned long long seq = e->p2p_seq + 1;
    const int par = (int)(seq & 1ull);
    DevParams p = make_params(e);
    if (phase == 0) {  // my top rung into the hotter neighbour's mailbox
        if (hottest) return 0;
        LaunchTimer lt(e, PTMCMC_K_SWAP);
        const P2PSync sync{nullptr, p2p_flag(e->p2p_above, n, par, 1), e->d_p2p_ctr, seq, e->d_p2p_err};
        swap_pack_top_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 2048), 256, 0, e->stream>>>(
            p, p2p_slot(e->p2p_above, n, par, 1), sync);
    } else if (phase == 1) {  // wait for the carry, sweep, carry into the colder neighbour's mailbox
        if (e->swept) return fail(e, PTMCMC_ERR_STATE, "no sharded swap sweep is due");
        const double *cin = hottest ? nullptr : p2p_slot(e->p2p_box, n, par, 0);
        double *cout = coldest ? nullptr : p2p_slot(e->p2p_below, n, par, 0);
        const P2PSync sync{hottest ? nullptr : p2p_flag(e->p2p_box, n, par, 0), coldest ? nullptr : p2p_flag(e->p2p_below, n, par, 0),
                           e->d_p2p_ctr + 1, seq, e->d_p2p_err};
        {
            LaunchTimer lt(e, PTMCMC_K_SWAP);
            swap_sweep_kernel<<<(e->W + 127) / 128, 128, 0, e->stream>>>(p, e->cfg.ladder_above, e->d_swap_prep, cin, cout, e->d_map,
                                                                          e->d_carry_code, e->d_carry_L, sync);
        }
        e->carry_in = cin;
        e->swept = true;
    } else if (phase == 2) {  // wait for the colder neighbour's top rung, resolve position 0, permute
        if (!e->swept) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_swap_p2p(2) before ptmcmc_swap_p2p(1)");
        const double *below_top = coldest ? nullptr : p2p_slot(e->p2p_box, n, par, 1);
        const P2PSync sync{coldest ? nullptr : p2p_flag(e->p2p_box, n, par, 1), nullptr, nullptr, seq, e->d_p2p_err};
        short *tr = nullptr;
        if (e->d_swapmaps && e->swap_events < e->cfg.trace_iters) tr = e->d_swapmaps + (size_t)e->swap_events * e->W * e->T;
        const int nxt = e->cur ^ 1;
        {
            LaunchTimer lt(e, PTMCMC_K_SWAP, coldest ? 1 : 2);
            if (!coldest) p2p_wait_kernel<<<1, 32, 0, e->stream>>>(sync);
            swap_finish_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, e->iter, e->Tg, e->cfg.ladder_below, e->d_map,
                                                                              e->d_carry_code, e->d_carry_L, e->carry_in, below_top,
                                                                              e->x[nxt], e->lnl[nxt], e->lp[nxt], tr, P2PSync{});
        }
        e->cur = nxt;
        e->swap_proposed++;
        e->swap_events++;
        e->pending_swap = false;
        e->swept = false;
        e->carry_in = nullptr;
        e->p2p_seq = seq;
        e->rows = std::max(e->rows, (long long)(e->iter / e->cfg.thin + 1));
        CUDA_TRY(e, sink_flush(e));
    } else {
        return fail(e, PTMCMC_ERR_ARG, "ptmcmc_swap_p2p: phase is 0 (pack), 1 (sweep) or 2 (finish)");
    }
    CUDA_TRY(e, cudaGetLastError());
    return 0;
}

int32_t ptmcmc_p2p_error(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->d_p2p_err) return 0;
    int err = 0;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaMemcpy(&err, e->d_p2p_err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return fail(e, PTMCMC_ERR_STATE, "a neighbour's swap message did not arrive within the time limit");
    return 0;
}

int32_t ptmcmc_am_ring(ptmcmc_engine *h, void **dev_ptr, int64_t *ndoubles)
{
    Engine *e = engine_of(h);
    if (!e || !dev_ptr || !ndoubles) return PTMCMC_ERR_ARG;
    *dev_ptr = e->d_am;
    *ndoubles = (int64_t)e->cfg.cov_update * e->d * e->W;
    return 0;
}

int32_t ptmcmc_maintain(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->has_state) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_maintain before ptmcmc_set_state");
    if (e->pending_swap || e->pending_propose) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_maintain inside an iteration");
    return maintenance(e, e->iter + 1);
}

namespace {

struct StateHeader {
    uint64_t magic;
    int32_t version, d, W, T, Tg, temp_offset, njumps, de_in_cycle;
    int64_t cov_update, burn, iter, de_head, swap_proposed, swap_events, nsamp, adapt_done_iter, de_done_iter;
    int64_t usize, ssize;
    // what keys the draws and the record layout: a checkpoint continues exactly only under the same values
    uint64_t seed, config_hash;
    int64_t thin, tskip;
    int32_t walker_offset, pad;
};

// FNV-1a over the ladder, the MH temperatures and the proposal cycle
uint64_t config_hash_of(const Engine *e)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    mix(e->ladder.data(), sizeof(double) * e->ladder.size());
    mix(e->mh_temp.data(), sizeof(double) * e->mh_temp.size());
    const int nc = e->cfg.ncycle;
    mix(e->cfg.cycle_jump, sizeof(int32_t) * nc);
    mix(e->cfg.cycle_weight, sizeof(int32_t) * nc);
    mix(&e->cfg.de_weight, sizeof(int32_t));
    return h;
}
constexpr uint64_t STATE_MAGIC = 0x3030324250544d43ull;  // "CMTPB200"

struct StateField {
    void *ptr;
    size_t bytes;
};

std::vector<StateField> state_fields(Engine *e)
{
    const size_t C = (size_t)e->T * e->W, d = e->d;
    const size_t us = e->uoff[e->ngroups], ss = e->soff[e->ngroups];
    return {{e->x[e->cur], C * d * 8},       {e->lnl[e->cur], C * 8},       {e->lp[e->cur], C * 8},
            {e->d_cov, d * d * 8},           {e->d_mu, d * 8},              {e->d_m2, d * d * 8},
            {e->d_U, us * 8},                {e->d_S, ss * 8},              {e->d_sqrtS, ss * 8},
            {e->d_am, (size_t)e->cfg.cov_update * d * e->W * 8},            {e->d_de, (size_t)e->cfg.burn * e->W * d * 8},
            {e->d_prop, C * e->njumps * 8},  {e->d_acc, C * e->njumps * 8}, {e->d_swap_acc, C * 8}};
}

}  // namespace

int64_t ptmcmc_state_bytes(const ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return -1;
    size_t n = sizeof(StateHeader);
    for (const StateField &f : state_fields(e)) n += f.bytes;
    return (int64_t)n;
}

int32_t ptmcmc_save_state(ptmcmc_engine *h, void *buf, int64_t nbytes)
{
    Engine *e = engine_of(h);
    if (!e || !buf) return PTMCMC_ERR_ARG;
    if (!e->has_state || e->pending_swap || e->pending_propose)
        return fail(e, PTMCMC_ERR_STATE, "ptmcmc_save_state: no state yet, or inside an iteration");
    if (nbytes < ptmcmc_state_bytes(h)) return fail(e, PTMCMC_ERR_ARG, "checkpoint buffer too small");
    StateHeader hd{};
    hd.magic = STATE_MAGIC; hd.version = PTMCMC_ABI_VERSION;
    hd.d = e->d; hd.W = e->W; hd.T = e->T; hd.Tg = e->Tg; hd.temp_offset = e->cfg.temp_offset; hd.njumps = e->njumps;
    hd.de_in_cycle = e->de_in_cycle ? 1 : 0;
    hd.cov_update = e->cfg.cov_update; hd.burn = e->cfg.burn; hd.iter = e->iter; hd.de_head = e->de_head;
    hd.swap_proposed = e->swap_proposed; hd.swap_events = e->swap_events; hd.nsamp = e->nsamp;
    hd.adapt_done_iter = e->adapt_done_iter; hd.de_done_iter = e->de_done_iter;
    hd.usize = e->uoff[e->ngroups]; hd.ssize = e->soff[e->ngroups];
    hd.seed = e->cfg.seed; hd.config_hash = config_hash_of(e);
    hd.thin = e->cfg.thin; hd.tskip = e->cfg.tskip; hd.walker_offset = e->cfg.walker_offset;
    char *out = (char *)buf;
    memcpy(out, &hd, sizeof hd);
    out += sizeof hd;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (const StateField &f : state_fields(e)) {
        CUDA_TRY(e, cudaMemcpyAsync(out, f.ptr, f.bytes, cudaMemcpyDeviceToHost, e->stream));
        out += f.bytes;
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int32_t ptmcmc_load_state(ptmcmc_engine *h, const void *buf, int64_t nbytes)
{
    Engine *e = engine_of(h);
    if (!e || !buf) return PTMCMC_ERR_ARG;
    if (nbytes < (int64_t)sizeof(StateHeader)) return fail(e, PTMCMC_ERR_ARG, "checkpoint truncated");
    StateHeader hd;
    memcpy(&hd, buf, sizeof hd);
    if (hd.magic != STATE_MAGIC || hd.version != PTMCMC_ABI_VERSION) return fail(e, PTMCMC_ERR_ARG, "not a checkpoint of this ABI version");
    if (hd.d != e->d || hd.W != e->W || hd.T != e->T || hd.Tg != e->Tg || hd.temp_offset != e->cfg.temp_offset ||
        hd.cov_update != e->cfg.cov_update || hd.burn != e->cfg.burn || hd.usize != e->uoff[e->ngroups] ||
        hd.ssize != e->soff[e->ngroups])
        return fail(e, PTMCMC_ERR_ARG, "checkpoint was written by an engine with a different configuration");
    if (hd.njumps != e->njumps) return fail(e, PTMCMC_ERR_ARG, "checkpoint has %d jump kinds, engine %d", hd.njumps, e->njumps);
    if (hd.seed != e->cfg.seed)
        return fail(e, PTMCMC_ERR_ARG, "checkpoint was written with seed %llu, engine has %llu: the draws are keyed by the seed, "
                    "so the run would not continue where it stopped (ptmcmc_state_seed reads it back)",
                    (unsigned long long)hd.seed, (unsigned long long)e->cfg.seed);
    if (hd.thin != e->cfg.thin || hd.tskip != e->cfg.tskip || hd.walker_offset != e->cfg.walker_offset)
        return fail(e, PTMCMC_ERR_ARG, "checkpoint was written with thin=%lld Tskip=%lld walker_offset=%d, engine has %lld / %lld / %d",
                    (long long)hd.thin, (long long)hd.tskip, hd.walker_offset, (long long)e->cfg.thin, (long long)e->cfg.tskip,
                    e->cfg.walker_offset);
    if (hd.config_hash != config_hash_of(e))
        return fail(e, PTMCMC_ERR_ARG, "checkpoint was written with a different ladder or proposal cycle");
    if (nbytes < ptmcmc_state_bytes(h)) return fail(e, PTMCMC_ERR_ARG, "checkpoint truncated");
    const char *in = (const char *)buf + sizeof hd;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (const StateField &f : state_fields(e)) {
        CUDA_TRY(e, cudaMemcpyAsync(f.ptr, in, f.bytes, cudaMemcpyHostToDevice, e->stream));
        in += f.bytes;
    }
    e->iter = hd.iter; e->de_head = hd.de_head; e->swap_proposed = hd.swap_proposed; e->swap_events = hd.swap_events;
    e->nsamp = hd.nsamp; e->adapt_done_iter = hd.adapt_done_iter; e->de_done_iter = hd.de_done_iter;
    if (hd.de_in_cycle && !e->de_in_cycle && e->cfg.de_weight > 0) {  // DE had joined the cycle (ref :563-585)
        e->cyc_jump.push_back(PTMCMC_JUMP_DE);
        e->cyc_w.push_back(e->cfg.de_weight);
        e->de_in_cycle = true;
    }
    e->rows = e->iter / e->cfg.thin + 1;
    e->rec_base = e->rows;  // the record window restarts empty after the checkpointed iteration
    e->copied = e->rows;
    e->has_state = true;
    e->pending_swap = e->swept = e->pending_propose = false;
    cudaError_t st = build_u_frags(e);
    if (st != cudaSuccess) return fail(e, PTMCMC_ERR_CUDA, "fragment rebuild: %s", cudaGetErrorString(st));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int32_t ptmcmc_state_seed(const void *buf, int64_t nbytes, uint64_t *seed)
{
    if (!buf || !seed || nbytes < (int64_t)sizeof(StateHeader)) return PTMCMC_ERR_ARG;
    StateHeader hd;
    memcpy(&hd, buf, sizeof hd);
    if (hd.magic != STATE_MAGIC || hd.version != PTMCMC_ABI_VERSION) return PTMCMC_ERR_ARG;
    *seed = hd.seed;
    return 0;
}

int32_t ptmcmc_replay(ptmcmc_engine *h, int64_t niter, int64_t repeat, int64_t nrows, const double *x, const double *lnl,
                      const double *lnprior)
{
    Engine *e = engine_of(h);
    if (!e || !x || !lnl || !lnprior || repeat < 1 || nrows < 1 || niter < 0) return PTMCMC_ERR_ARG;
    if (!e->has_state) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_replay before ptmcmc_set_state");
    if (e->sharded || e->pending_propose) return fail(e, PTMCMC_ERR_STATE, "ptmcmc_replay on a sharded engine or inside an iteration");
    const long long end = e->iter + niter;
    const long long row_base = (e->iter + 1) / repeat;
    if (end / repeat - row_base >= nrows) return fail(e, PTMCMC_ERR_ARG, "ptmcmc_replay: %lld rows do not cover iterations up to %lld", (long long)nrows, end);
    int rc = check_rows(e, end);
    if (rc) return rc;
    const size_t C = (size_t)e->T * e->W;
    double *dx = nullptr, *dl = nullptr, *dp = nullptr;
    g_alloc_stream = e->stream;
    CUDA_TRY(e, dalloc(&dx, (size_t)nrows * C * e->d));
    CUDA_TRY(e, dalloc(&dl, (size_t)nrows * C));
    CUDA_TRY(e, dalloc(&dp, (size_t)nrows * C));
    CUDA_TRY(e, cudaMemcpyAsync(dx, x, sizeof(double) * nrows * C * e->d, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(dl, lnl, sizeof(double) * nrows * C, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(dp, lnprior, sizeof(double) * nrows * C, cudaMemcpyHostToDevice, e->stream));
    const long long cu = e->cfg.cov_update, burn = e->cfg.burn;
    while (e->iter < end) {
        const long long it0 = e->iter + 1;
        rc = maintenance(e, it0);
        if (rc) break;
        long long seg_end = std::min<long long>(end, std::min(next_multiple(it0, cu), next_multiple(it0, burn)));
        DevParams p = make_params(e);
        p.it0 = it0; p.it1 = seg_end; p.tail = 1;
        {
            LaunchTimer lt(e, PTMCMC_K_INIT);
            replay_kernel<<<chain_blocks(e), MH_THREADS, 0, e->stream>>>(p, dx, dl, dp, repeat, row_base);
        }
        e->iter = seg_end;
    }
    cudaFreeAsync(dx, e->stream);
    cudaFreeAsync(dl, e->stream);
    cudaFreeAsync(dp, e->stream);
    if (rc) return rc;
    CUDA_TRY(e, cudaGetLastError());
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->rows = std::max(e->rows, (long long)(e->iter / e->cfg.thin + 1));
    CUDA_TRY(e, sink_flush(e));
    return 0;
}

int32_t ptmcmc_njumps(const ptmcmc_engine *h) { return h ? ((const Engine *)h)->njumps : -1; }

int32_t ptmcmc_get_counters(ptmcmc_engine *h, int64_t *prop, int64_t *acc, int64_t *swap_acc, int64_t *swap_proposed)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    const size_t C = (size_t)e->T * e->W;
    const int nj = e->njumps;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (prop) CUDA_TRY(e, cudaMemcpy(prop, e->d_prop, sizeof(int64_t) * C * nj, cudaMemcpyDeviceToHost));
    if (acc) CUDA_TRY(e, cudaMemcpy(acc, e->d_acc, sizeof(int64_t) * C * nj, cudaMemcpyDeviceToHost));
    if (swap_acc) CUDA_TRY(e, cudaMemcpy(swap_acc, e->d_swap_acc, sizeof(int64_t) * C, cudaMemcpyDeviceToHost));
    if (swap_proposed) *swap_proposed = e->swap_proposed;
    return 0;
}

int32_t ptmcmc_get_trace(ptmcmc_engine *h, uint8_t *trace, int64_t iters, int16_t *swapmaps, int64_t events)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (!e->cfg.trace) return fail(e, PTMCMC_ERR_STATE, "engine created without trace");
    const size_t C = (size_t)e->T * e->W;
    if (iters > e->cfg.trace_iters || events > e->cfg.trace_iters) return fail(e, PTMCMC_ERR_ARG, "trace request exceeds capacity");
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (trace) CUDA_TRY(e, cudaMemcpy(trace, e->d_trace, (size_t)iters * C, cudaMemcpyDeviceToHost));
    if (swapmaps) CUDA_TRY(e, cudaMemcpy(swapmaps, e->d_swapmaps, sizeof(short) * (size_t)events * C, cudaMemcpyDeviceToHost));
    return 0;
}

int32_t ptmcmc_set_sink(ptmcmc_engine *h, double *chain, double *lnl, double *lnprob, int64_t row_capacity)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    if (e->copy_stream) CUDA_TRY(e, cudaStreamSynchronize(e->copy_stream));
    if (!chain) {
        e->sink_x = e->sink_lnl = e->sink_lnp = nullptr;
        e->sink_cap = 0;
        return 0;
    }
    if (!lnl || !lnprob || row_capacity < 1) return fail(e, PTMCMC_ERR_ARG, "ptmcmc_set_sink: three arrays and a row capacity");
    e->sink_x = chain; e->sink_lnl = lnl; e->sink_lnp = lnprob;
    e->sink_cap = row_capacity;
    e->copied = e->has_state ? e->rows : 0;
    return 0;
}

int64_t ptmcmc_snapshot_bytes(const ptmcmc_engine *h)
{
    const Engine *e = (const Engine *)h;
    if (!e) return -1;
    const size_t ni = 4 + (size_t)4 * e->njumps * e->T + 2 * (size_t)e->T;
    const size_t nd = 2 * (size_t)e->d * e->d + e->d + e->uoff[e->ngroups] + e->soff[e->ngroups];
    return (int64_t)(8 * (ni + nd));
}

int32_t ptmcmc_snapshot(ptmcmc_engine *h, void *host, int64_t nbytes, int32_t slot)
{
    Engine *e = engine_of(h);
    if (!e || !host || slot < 0 || slot > 1) return PTMCMC_ERR_ARG;
    if (nbytes < ptmcmc_snapshot_bytes(h)) return fail(e, PTMCMC_ERR_ARG, "snapshot buffer too small");
    const int nj = e->njumps, T = e->T, d = e->d;
    const size_t ns = (size_t)4 * nj * T + 2 * (size_t)T;
    const size_t nd = 2 * (size_t)d * d + d + e->uoff[e->ngroups] + e->soff[e->ngroups];
    if (!e->d_snap[slot]) {
        g_alloc_stream = e->stream;
        CUDA_TRY(e, dalloc(&e->d_snap[slot], ns + nd));
    }
    // staged on the compute stream (the next covariance update may rewrite cov / factor right away), copied out on the
    // copy stream: the compute stream never waits for the host transfer
    long long *stage = e->d_snap[slot];
    counter_summary_kernel<<<nj * T, 256, 0, e->stream>>>(e->d_prop, e->d_acc, e->d_swap_acc, nj, T, e->W, stage);
    e->tm.launches[PTMCMC_K_INIT] += 1;
    CUDA_TRY(e, cudaGetLastError());
    double *sd = (double *)(stage + ns);
    struct { const double *src; size_t n; } parts[] = {{e->d_cov, (size_t)d * d}, {e->d_mu, (size_t)d}, {e->d_m2, (size_t)d * d},
                                                       {e->d_U, (size_t)e->uoff[e->ngroups]}, {e->d_S, (size_t)e->soff[e->ngroups]}};
    for (auto &pt : parts) {
        CUDA_TRY(e, cudaMemcpyAsync(sd, pt.src, 8 * pt.n, cudaMemcpyDeviceToDevice, e->stream));
        sd += pt.n;
    }
    CUDA_TRY(e, sink_flush(e));
    CUDA_TRY(e, cudaEventRecord(e->ev_rows, e->stream));
    CUDA_TRY(e, cudaStreamWaitEvent(e->copy_stream, e->ev_rows, 0));
    long long *hi = (long long *)host;
    hi[0] = e->iter; hi[1] = e->swap_proposed; hi[2] = e->nsamp; hi[3] = nj;
    CUDA_TRY(e, cudaMemcpyAsync(hi + 4, stage, 8 * (ns + nd), cudaMemcpyDeviceToHost, e->copy_stream));
    CUDA_TRY(e, cudaEventRecord(e->ev_snap[slot], e->copy_stream));
    return 0;
}

int32_t ptmcmc_snapshot_wait(ptmcmc_engine *h, int32_t slot)
{
    Engine *e = engine_of(h);
    if (!e || slot < 0 || slot > 1) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, cudaEventSynchronize(e->ev_snap[slot]));
    return 0;
}

int32_t ptmcmc_sink_wait(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, sink_flush(e));
    CUDA_TRY(e, cudaStreamSynchronize(e->copy_stream));
    return 0;
}

int32_t ptmcmc_get_timing(ptmcmc_engine *h, ptmcmc_timing *out)
{
    Engine *e = engine_of(h);
    if (!e || !out) return PTMCMC_ERR_ARG;
    *out = e->tm;
    return 0;
}

int32_t ptmcmc_reset_timing(ptmcmc_engine *h)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    e->tm = ptmcmc_timing{};
    return 0;
}

void *ptmcmc_stream(ptmcmc_engine *h) { return h ? (void *)((Engine *)h)->stream : nullptr; }

const char *ptmcmc_mh_kernel_name(ptmcmc_engine *h)
{
    Engine *e = (Engine *)h;
    if (!e) return "";
    if (use_mma(e)) {
        static thread_local char buf[64];
        snprintf(buf, sizeof buf, "%s<%d> (%d chains per block)", e->mma.split ? "mh_mma_split_kernel" : "mh_mma_kernel", e->mma.nt,
                 e->mma.nc);
        return buf;
    }
    if (fast_reg_path(e) && e->mh_variant != 2) return sorted_kernel_name(e->d, e->sorted);
    return e->user ? "mh_generic_kernel (NVRTC, user target)" : "mh_generic_kernel";
}

int32_t ptmcmc_user_compile_check(const char *logl_source, const char *logp_source, int32_t cc_major, int32_t cc_minor,
                                  char *log, int64_t log_capacity)
{
    std::vector<char> cubin;
    std::string msg;
    const int rc = user_compile_cubin(logl_source, logp_source, cc_major, cc_minor, cubin, msg);
    if (log && log_capacity > 0) {
        strncpy(log, msg.c_str(), (size_t)log_capacity - 1);
        log[log_capacity - 1] = 0;
    }
    return rc == 0 ? (int32_t)std::min<size_t>(cubin.size(), 0x7FFFFFFF) : PTMCMC_ERR_ARG;
}

int32_t ptmcmc_measure_fp64_peak(int32_t device, double *tflops)
{
    if (!tflops) return PTMCMC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PTMCMC_ERR_CUDA;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return PTMCMC_ERR_CUDA;
    const int blocks = 2 * sms, threads = 512, iters = 40000;
    double *out = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc((void **)&out, sizeof(double) * blocks * threads) != cudaSuccess) return PTMCMC_ERR_CUDA;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 0.999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * blocks * (double)threads * iters * 8 / (ms * 1e-3) / 1e12;
        if (rep && tf > best) best = tf;
    }
    cudaError_t st = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return st == cudaSuccess ? 0 : PTMCMC_ERR_CUDA;
}

int32_t ptmcmc_test_normals(int32_t device, const uint64_t *words, int64_t n, double *z0, double *z1)
{
    if (!words || !z0 || !z1 || n < 0) return PTMCMC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PTMCMC_ERR_CUDA;
    unsigned long long *dw = nullptr;
    double *d0 = nullptr, *d1 = nullptr;
    const size_t m = (size_t)(n ? n : 1);
    cudaError_t st = cudaMalloc((void **)&dw, 8 * m);
    if (st == cudaSuccess) st = cudaMalloc((void **)&d0, 8 * m);
    if (st == cudaSuccess) st = cudaMalloc((void **)&d1, 8 * m);
    if (st == cudaSuccess) st = cudaMemcpy(dw, words, 8 * (size_t)n, cudaMemcpyHostToDevice);
    if (st == cudaSuccess) {
        normals_kernel<<<(unsigned)std::min<long long>((n + 255) / 256 + 1, 4096), 256>>>(dw, n, d0, d1);
        st = cudaGetLastError();
    }
    if (st == cudaSuccess) st = cudaMemcpy(z0, d0, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    if (st == cudaSuccess) st = cudaMemcpy(z1, d1, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(dw); cudaFree(d0); cudaFree(d1);
    return st == cudaSuccess ? 0 : PTMCMC_ERR_CUDA;
}

int32_t ptmcmc_set_timing(ptmcmc_engine *h, int32_t on)
{
    Engine *e = engine_of(h);
    if (!e) return PTMCMC_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->cfg.timing = on ? 1 : 0;
    return 0;
}

void *ptmcmc_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void ptmcmc_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"
