// Fused Metropolis-Hastings kernels: proposal draw (SCAM / AM / DE), built-in log-prior and
// log-likelihood, Hastings test, AM-ring write and thinned record, for a whole segment of
// iterations per launch with the chain state resident on chip.
//
// Follows ref PTMCMCSampler.py: PTMCMCOneStep :601-627, _jump :1048-1067, SCAM :820-876,
// AM :879-933, DE :936-985, updateChains :321-335.
#pragma once
#include "mh_common.cuh"

// This header is also the body of the run-time compiled translation unit of a user target (user_target.cu): the
// caller's CUDA source defines user_logl / user_logp ahead of it, PTMCMC_USER_TARGET is defined, and the kernels get C
// linkage so that the driver API finds them by name.
#ifdef PTMCMC_USER_TARGET
#define PTM_KERNEL extern "C" __global__
#else
#define PTM_KERNEL __global__
#endif

namespace ptm {

// ---------------------------------------------------------------------------------------------
// Generic path (any ndim <= MAX_GENERIC_DIM, arbitrary parameter groups): thread per chain with
// the vectors in (L1-resident) local memory.  Also the building block of the host-callback path.
// ---------------------------------------------------------------------------------------------
__device__ inline double eval_logp_generic(const DevParams &p, const double *q)
{
#ifdef PTMCMC_USER_TARGET
    if (p.logp_kind == LOGP_USER) return ::user_logp(q, p.d, p.user_par + p.n_logl_par);
#endif
    if (p.logp_kind == LOGP_UNIFORM) {
        for (int k = 0; k < p.d; ++k)
            if (!in_box(q[k], p.p_lo[k], p.p_hi[k], p.p_inclusive)) return neg_inf();
        return p.p_inside;
    }
    return 0.0;
}

__device__ inline double eval_logl_generic(const DevParams &p, const double *q)
{
    const int d = p.d;
#ifdef PTMCMC_USER_TARGET
    if (p.logl_kind == LOGL_USER) return ::user_logl(q, d, p.user_par);
#endif
    if (p.logl_kind == LOGL_GAUSSIAN) {
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            double row = 0.0;
            for (int j = i; j < d; ++j) row = fma(__ldg(p.g_P + i * d + j), q[j] - __ldg(p.g_mu + j), row);
            acc = fma(q[i] - __ldg(p.g_mu + i), row, acc);
        }
        return acc + p.g_offset;
    }
    if (p.logl_kind == LOGL_CURVED) {
        double tot = 0.0;
        for (int b = 0; b + 1 < d; b += 2) {
            const double a = q[b], y = q[b + 1];
            const double t0 = 9.0 + 4.0 * a * a + 9.0 * y;
            tot += log(exp(-a * a - t0 * t0) + 0.5 * exp(-8.0 * a * a - 8.0 * (y - 2.0) * (y - 2.0)));
        }
        return tot;
    }
    double tot = 0.0;
    for (int i = 0; i + 1 < d; ++i) {
        const double a = q[i + 1] - q[i] * q[i], b = 1.0 - q[i];
        tot -= 100.0 * a * a + b * b;
    }
    return tot / 20.0;
}

// Draw the proposal of chain (t, w) into q (q holds x on entry).  Returns the jump id; for host
// jumps (id >= JUMP_EXT0) q is left equal to x.
__device__ inline int propose_generic(const DevParams &p, Stream &st, double temp, double beta,
                                      const double *x, double *q, double *dl)
{
    const int jump = pick_jump(p, st);
    if (jump >= JUMP_EXT0) return jump;
    if (jump == JUMP_PRIOR) {
        // a fresh draw from the uniform prior box, one uniform per parameter (the "UniformJump" plugin of
        // ref tests/test_simple.py:44-62 on the device); symmetric: qxy = 0
        for (int k = 0; k < p.d; ++k) {
            const double lo = __ldg(p.p_lo + k), hi = __ldg(p.p_hi + k);
            q[k] = lo + (hi - lo) * word_to_unit(st.next());
        }
        return jump;
    }
    const int g = (int)word_to_int(st.next(), (unsigned long long)p.ngroups);
    const int g0 = p.goff[g], dg = p.goff[g + 1] - g0;
    const int *gi = p.gidx + g0;
    if (jump == JUMP_SCAM || jump == JUMP_AM) {
        const double *U = p.U + p.uoff[g], *sS = p.sqrtS + p.soff[g];
        const double prob = word_to_unit(st.next());
        const double scale = cov_jump_scale(prob, temp);
        if (jump == JUMP_SCAM) {
            const int k = (int)word_to_int(st.next(), (unsigned long long)dg);
            const double cd = 2.4 / sqrt(2.0) * scale;
            double z0, z1;
            word_to_normals(st.next(), z0, z1);
            const double coef = z0 * cd * __ldg(sS + k);
            for (int i = 0; i < dg; ++i) q[gi[i]] = fma(coef, __ldg(U + i * dg + k), q[gi[i]]);
        } else {
            const double cd = 2.4 / sqrt(2.0 * dg) * scale;
            for (int j = 0; j < dg; j += 2) {
                double z0, z1;
                word_to_normals(st.next(), z0, z1);
                dl[j] = z0 * cd * __ldg(sS + j);
                if (j + 1 < dg) dl[j + 1] = z1 * cd * __ldg(sS + j + 1);
            }
            for (int i = 0; i < dg; ++i) {
                double a = x[gi[i]];
                for (int j = 0; j < dg; ++j) a = fma(__ldg(U + i * dg + j), dl[j], a);
                q[gi[i]] = a;
            }
        }
    } else {
        const unsigned long long W = (unsigned long long)p.W;
        const unsigned long long bufsize = (unsigned long long)p.burn * W;
        const unsigned long long mm = word_to_int(st.next(), bufsize);
        unsigned long long nn = word_to_int(st.next(), bufsize);
        while (mm == nn) nn = word_to_int(st.next(), bufsize);
        const double prob = word_to_unit(st.next());
        double scale = 1.0;
        if (!(prob > 0.5)) scale = word_to_unit(st.next()) * 2.4 / sqrt(2.0 * dg) * sqrt(1.0 / beta);
        const unsigned long long sm_ = mm / W, sn_ = nn / W;
        const double *bm = p.de + (((sm_ + p.de_head) % p.burn) * W + (mm - sm_ * W)) * p.d;
        const double *bn = p.de + (((sn_ + p.de_head) % p.burn) * W + (nn - sn_ * W)) * p.d;
        for (int i = 0; i < dg; ++i) {
            const double sigma = __ldg(bm + gi[i]) - __ldg(bn + gi[i]);
            q[gi[i]] = fma(scale, sigma, q[gi[i]]);
        }
    }
    return jump;
}

PTM_KERNEL void __launch_bounds__(MH_THREADS) mh_generic_kernel(const DevParams p)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const uint32_t gw = (uint32_t)(p.walker_offset + w), gt = (uint32_t)(p.temp_offset + t);
    const double temp = p.mh_temp[t];
    const double beta = 1.0 / temp;
    double x[MAX_GENERIC_DIM], q[MAX_GENERIC_DIM], dl[MAX_GENERIC_DIM];
    double *xg = p.x + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) x[k] = xg[(size_t)k * W];
    double lnl = p.lnl[c], lp = p.lp[c];
    const size_t TW = (size_t)T * W;
    for (long long it = p.it0; it <= p.it1; ++it) {
        Stream st(p, PURPOSE_MH, (unsigned long long)it, gw, gt);
        for (int k = 0; k < d; ++k) q[k] = x[k];
        const int jump = propose_generic(p, st, temp, beta, x, q, dl);
        const double lpn = eval_logp_generic(p, q);
        double lnln = 0.0, lnpn = neg_inf();
        if (lpn != neg_inf()) {
            lnln = eval_logl_generic(p, q);
            lnpn = beta * lnln + lpn;
        }
        const double lnp0 = beta * lnl + lp;
        const double diff = lnpn - lnp0;
        const bool accept = hastings_accept(diff, st.next());
        if (accept) {
            for (int k = 0; k < d; ++k) x[k] = q[k];
            lnl = lnln;
            lp = lpn;
            p.acc[jump * TW + c] += 1;
        }
        p.prop[jump * TW + c] += 1;
        if (p.trace && it - 1 < p.trace_cap)
            p.trace[((size_t)(it - 1) * T + t) * W + w] = (unsigned char)(jump | ((int)accept << 7));
        if (it < p.it1 || p.tail)
            bookkeep(p, it, t, w, [&](int k) { return x[k]; }, lnl, lp, beta);
    }
    for (int k = 0; k < d; ++k) xg[(size_t)k * W] = x[k];
    p.lnl[c] = lnl;
    p.lp[c] = lp;
}

// Initial point (ref :478-487): lp = logp(p0); lnlike0 = -inf outside the prior else logl(p0).  A target the host
// evaluates (LOG*_EXTERNAL) keeps the value uploaded beforehand (ptmcmc_set_state_external), so a device prior can be
// combined with a Python likelihood and vice versa.
PTM_KERNEL void __launch_bounds__(MH_THREADS) init_eval_kernel(const DevParams p)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    double x[MAX_GENERIC_DIM];
    const double *xg = p.x + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) x[k] = xg[(size_t)k * W];
    const double lp = (p.logp_kind != LOGP_EXTERNAL) ? eval_logp_generic(p, x) : p.lp[c];
    p.lp[c] = lp;
    p.lnl[c] = (lp == neg_inf()) ? neg_inf() : (p.logl_kind != LOGL_EXTERNAL) ? eval_logl_generic(p, x) : p.lnl[c];
}

// updateChains for iteration `it` of every chain from the state in HBM (used for iteration 0 and by
// the host-callback path).
PTM_KERNEL void __launch_bounds__(MH_THREADS) bookkeep_kernel(const DevParams p, long long it)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const double *xg = p.x + (size_t)t * d * W + w;
    bookkeep(p, it, t, w, [&](int k) { return xg[(size_t)k * W]; }, p.lnl[c], p.lp[c], 1.0 / p.mh_temp[t]);
}

// Reference-style resume (ref :591-599): iterations [it0, it1] take their state from stored rows
// (host layout rows_x[row][T][W][d]); row of iteration it = it / repeat - row_base.  Only the buffer /
// record side effects of the step happen (ref :627).
PTM_KERNEL void __launch_bounds__(MH_THREADS) replay_kernel(const DevParams p, const double *rows_x,
                                                            const double *rows_lnl, const double *rows_lp,
                                                            long long repeat, long long row_base)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const double beta = 1.0 / p.mh_temp[t];
    const size_t TW = (size_t)T * W;
    const double *xr = nullptr;
    double lnl = 0.0, lp = 0.0;
    for (long long it = p.it0; it <= p.it1; ++it) {
        const size_t r = (size_t)(it / repeat - row_base);
        xr = rows_x + (r * TW + c) * d;
        lnl = rows_lnl[r * TW + c];
        lp = rows_lp[r * TW + c];
        bookkeep(p, it, t, w, [&](int k) { return xr[k]; }, lnl, lp, beta);
    }
    double *xg = p.x + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) xg[(size_t)k * W] = xr[k];
    p.lnl[c] = lnl;
    p.lp[c] = lp;
}

// ---- host-callback path (Python logl / logp / custom jumps), one iteration per call pair ------
// q_out [T][W][d] row-major for the host, jump_out [T][W], word_pos [T][W] = stream position
PTM_KERNEL void __launch_bounds__(MH_THREADS) propose_kernel(const DevParams p, double *q_out, int *jump_out,
                                                              unsigned *word_pos)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const double temp = p.mh_temp[t];
    double x[MAX_GENERIC_DIM], q[MAX_GENERIC_DIM], dl[MAX_GENERIC_DIM];
    const double *xg = p.x + (size_t)t * d * W + w;
    for (int k = 0; k < d; ++k) q[k] = x[k] = xg[(size_t)k * W];
    Stream st(p, PURPOSE_MH, (unsigned long long)p.it0, (uint32_t)(p.walker_offset + w),
              (uint32_t)(p.temp_offset + t));
    const int jump = propose_generic(p, st, temp, 1.0 / temp, x, q, dl);
    for (int k = 0; k < d; ++k) q_out[c * d + k] = q[k];
    jump_out[c] = jump;
    word_pos[c] = st.j;
}

// q_in [T][W][d], qxy/lnl_new/lp_new [T][W]
PTM_KERNEL void __launch_bounds__(MH_THREADS) accept_kernel(const DevParams p, const double *q_in, const double *qxy,
                                                             const double *lnl_new, const double *lp_new,
                                                             const int *jump_in, const unsigned *word_pos)
{
    const int d = p.d, W = p.W, T = p.T;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)T * W) return;
    const int t = (int)(c / W), w = (int)(c % W);
    const double beta = 1.0 / p.mh_temp[t];
    Stream st(p, PURPOSE_MH, (unsigned long long)p.it0, (uint32_t)(p.walker_offset + w),
              (uint32_t)(p.temp_offset + t));
    st.seek(word_pos[c]);
    const int jump = jump_in[c];
    // built-in targets are evaluated here; external ones come from the host (ref :605-612)
    double qv[MAX_GENERIC_DIM];
    for (int k = 0; k < d; ++k) qv[k] = q_in[c * d + k];
    const double lpn = (p.logp_kind != LOGP_EXTERNAL) ? eval_logp_generic(p, qv) : lp_new[c];
    double lnln = 0.0, lnpn = neg_inf();
    if (lpn != neg_inf()) {
        lnln = (p.logl_kind != LOGL_EXTERNAL) ? eval_logl_generic(p, qv) : lnl_new[c];
        lnpn = beta * lnln + lpn;
    }
    const double lnp0 = beta * p.lnl[c] + p.lp[c];
    const double diff = lnpn - lnp0 + qxy[c];
    const bool accept = hastings_accept(diff, st.next());
    const size_t TW = (size_t)T * W;
    if (accept) {
        double *xg = p.x + (size_t)t * d * W + w;
        for (int k = 0; k < d; ++k) xg[(size_t)k * W] = qv[k];
        p.lnl[c] = lnln;
        p.lp[c] = lpn;
        p.acc[jump * TW + c] += 1;
    }
    p.prop[jump * TW + c] += 1;
    if (p.trace && p.it0 - 1 < p.trace_cap)
        p.trace[((size_t)(p.it0 - 1) * T + t) * W + w] = (unsigned char)(jump | ((int)accept << 7));
}

}  // namespace ptm
