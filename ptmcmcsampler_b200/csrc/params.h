// Device-visible parameter block of the PT-MCMC engine (passed by value to every kernel).
#pragma once
#ifdef __CUDACC_RTC__  // NVRTC ships no standard headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <cstdint>
#endif

namespace ptm {

constexpr int MAX_CYCLE = 16;
constexpr int MAX_REG_DIM = 32;       // shared-memory (sorted) kernel covers ndim <= 32
constexpr int MAX_GENERIC_DIM = 128;  // local-memory kernel covers ndim <= 128
constexpr int MH_THREADS = 128;

enum : int { JUMP_SCAM = 0, JUMP_AM = 1, JUMP_DE = 2, JUMP_PRIOR = 3, JUMP_EXT0 = 4 };
enum : int { LOGL_EXTERNAL = 0, LOGL_GAUSSIAN = 1, LOGL_CURVED = 2, LOGL_ROSENBROCK = 3, LOGL_USER = 4 };
enum : int { LOGP_EXTERNAL = 0, LOGP_UNIFORM = 1, LOGP_FLAT = 2, LOGP_USER = 3 };

struct DevParams {
    // geometry
    int d, W, T;
    int walker_offset, temp_offset;
    int ngroups, identity_group, njumps;
    unsigned long long seed;
    uint32_t rk[20];  // Philox round keys of the seed (philox_round_keys), read as constant-bank operands
    // chain state, SoA with the walker index fastest: x[T][d][W], lnl/lp[T][W]
    double *x, *lnl, *lp;
    const double *mh_temp, *ladder;  // [T]
    // pooled adaptive factor, concatenated per group: U_g row-major (d_g x d_g), sqrt(S_g)
    const double *U, *sqrtS;
    const int *goff, *gidx, *uoff, *soff;
    // proposal cycle (weight-replicated list of the reference, stored as cumulative weights)
    int ncycle, total_weight;
    int cyc_jump[MAX_CYCLE], cyc_cum[MAX_CYCLE];
    // AM ring [cov_update][d][W] (cold walkers), DE history [burn][W][d] as a ring with head slot
    double *am;
    const double *de;
    long long cov_update, burn, de_head;
    // targets
    int logl_kind, logp_kind, p_inclusive, pad0;
    const double *g_mu, *g_P;  // Gaussian: mean and -0.5*icov folded to an upper-triangular form
    double g_offset, p_inside;
    const double *p_lo, *p_hi;
    // user targets (CUDA source compiled with NVRTC): parameter blob [logl parameters | logp parameters]
    const double *user_par;
    int n_logl_par, n_logp_par;
    // thinned record window: rec_x[row][ntr][W][d], rec_lnl/rec_lnp[row][ntr][W]
    double *rec_x, *rec_lnl, *rec_lnp;
    long long rec_base, rec_cap, thin;
    int ntr, pad1;
    // counters [njumps][T][W]
    unsigned long long *prop, *acc, *swap_acc;
    // test trace: one byte per chain-step [iter-1][T][W]
    unsigned char *trace;
    long long trace_cap;
    // iteration range of this launch; tail = also do iteration it1's buffer/record bookkeeping
    long long it0, it1;
    int tail, pad2;
};

#ifndef __CUDACC_RTC__
// round r of Philox4x32-10 uses the key (k0 + r W0, k1 + r W1), Weyl constants of Salmon et al.
inline void philox_round_keys(unsigned long long seed, uint32_t *rk)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk[2 * r] = k0;
        rk[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
#endif

}  // namespace ptm
