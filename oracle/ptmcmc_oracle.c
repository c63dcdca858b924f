/*
 * ptmcmc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ptmcmc_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Every function cites the
 * lines of /root/reference/PTMCMCSampler/PTMCMCSampler.py (abbreviated "ref")
 * it follows.  The single semantic extension is the walker axis: W independent
 * ladders share ONE adaptive state (covariance, eigen-factor, DE history) that
 * is pooled over the W cold (T=1) chains, sample order (time slot major,
 * walker minor).  With W == 1 every formula below degenerates to the
 * reference's exactly.
 *
 * Randomness: Philox4x32-10 (Salmon et al., SC'11), counter =
 * (iter, purpose<<24 | block, walker, temperature), key = 64-bit seed.  A chain
 * consumes its 64-bit words strictly sequentially inside one iteration, in the
 * reference's own draw order.
 */
#include "ptmcmc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_CYCLE 32

/* ------------------------------------------------------------------ RNG -- */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t c0, c1base, c2, c3, k0, k1;
    uint32_t j;
    uint32_t blk[4];
} orc_stream;

static void stream_init(orc_stream *st, uint64_t seed, uint32_t purpose, uint64_t iter,
                        uint32_t walker, uint32_t temp)
{
    st->c0 = (uint32_t)iter;
    st->c1base = purpose << 24;
    st->c2 = walker;
    st->c3 = temp;
    st->k0 = (uint32_t)seed;
    st->k1 = (uint32_t)(seed >> 32);
    st->j = 0;
}

static uint64_t stream_word_at(const orc_stream *st, uint32_t j, uint32_t blk[4], int *have_blk)
{
    if (!*have_blk) {
        uint32_t ctr[4] = {st->c0, st->c1base | (j >> 1), st->c2, st->c3};
        uint32_t key[2] = {st->k0, st->k1};
        orc_philox4x32_10(ctr, key, blk);
        *have_blk = 1;
    }
    return (j & 1u) ? ((uint64_t)blk[2] | ((uint64_t)blk[3] << 32))
                    : ((uint64_t)blk[0] | ((uint64_t)blk[1] << 32));
}

static uint64_t stream_next(orc_stream *st)
{
    int have = (st->j & 1u) ? 1 : 0; /* odd word: block already generated */
    uint64_t w = stream_word_at(st, st->j, st->blk, &have);
    st->j++;
    return w;
}

uint64_t orc_draw_word(uint64_t seed, uint32_t purpose, uint64_t iter, uint32_t walker,
                       uint32_t temp, uint32_t j)
{
    orc_stream st;
    uint32_t blk[4];
    int have = 0;
    stream_init(&st, seed, purpose, iter, walker, temp);
    return stream_word_at(&st, j, blk, &have);
}

/* integer in [0, n): high 64 bits of word*n  (stands in for Generator.integers) */
uint64_t orc_word_to_int(uint64_t word, uint64_t n)
{
    return (uint64_t)(((unsigned __int128)word * (unsigned __int128)n) >> 64);
}

/* double in [0,1) with 53 random bits (stands in for Generator.random/uniform) */
double orc_word_to_unit(uint64_t word)
{
    return (double)(word >> 11) * (1.0 / 9007199254740992.0);
}

/* Box-Muller pair from one word (stands in for Generator.standard_normal): radius from the high 32
 * bits, angle from the low 32 bits, evaluated in SINGLE precision with individually rounded operations
 * (fmaf, *, -, sqrtf, integer ops; compiled with -ffp-contract=off) -- the same sequence as
 * ptmcmcsampler_b200/csrc/rng.cuh, so both sides agree to the bit.  u1 = (2 hi + 1) 2^-33 = m 2^e with a
 * 24-bit m; the angle keeps 24 of its 30 bits after the exact quadrant reduction. */
void orc_word_to_normals(uint64_t word, double *z0, double *z1)
{
    uint32_t hi = (uint32_t)(word >> 32), lo = (uint32_t)word;
    uint64_t n = ((uint64_t)hi << 1) | 1ull;
    int lz = __builtin_clzll(n);
    uint32_t mant = (uint32_t)((n << lz) >> 40);
    int e = 30 - lz;
    float m = (float)mant * 1.1920929e-07f;
    if (mant > 11863283u) {
        m = m * 0.5f;
        e += 1;
    }
    float t = m - 1.0f;
    float p = 0.0874394551f;
    p = fmaf(p, t, -0.143773302f);
    p = fmaf(p, t, 0.149490952f);
    p = fmaf(p, t, -0.165606961f);
    p = fmaf(p, t, 0.199569777f);
    p = fmaf(p, t, -0.250021547f);
    p = fmaf(p, t, 0.333341837f);
    p = fmaf(p, t, -0.499999881f);
    p = fmaf(p, t, 1.0f);
    float pt = p * t;
    float lnu = fmaf((float)e, 0.693147182f, pt);
    float r = sqrtf(-2.0f * lnu);
    uint32_t quad = lo >> 30, g = (lo >> 6) & 0xFFFFFFu;
    float f = (float)g * 2.98023224e-08f;
    int sw = g > 0x800000u;
    float x = sw ? 0.5f - f : f;
    float y = x * x;
    float sp = -0.589076877f;
    sp = fmaf(sp, y, 2.54976702f);
    sp = fmaf(sp, y, -5.16770792f);
    sp = fmaf(sp, y, 3.14159274f);
    float sx = sp * x;
    float cx = 0.231329247f;
    cx = fmaf(cx, y, -1.3350445f);
    cx = fmaf(cx, y, 4.05870724f);
    cx = fmaf(cx, y, -4.93480206f);
    cx = fmaf(cx, y, 1.0f);
    float s0 = sw ? cx : sx, c0 = sw ? sx : cx, s, c;
    switch (quad) {
    case 0: s = s0; c = c0; break;
    case 1: s = c0; c = -s0; break;
    case 2: s = -s0; c = -c0; break;
    default: s = -c0; c = s0; break;
    }
    *z0 = (double)(r * c);
    *z1 = (double)(r * s);
}

/* batch form for the distribution tests (tests/test_oracle_golden.py, tests/test_gpu_parity.py) */
void orc_word_to_normals_many(const uint64_t *words, int64_t n, double *z0, double *z1)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) orc_word_to_normals(words[i], &z0[i], &z1[i]);
}

static uint64_t draw_int(orc_stream *st, uint64_t n) { return orc_word_to_int(stream_next(st), n); }
static double draw_unit(orc_stream *st) { return orc_word_to_unit(stream_next(st)); }
static double draw_normal(orc_stream *st)
{
    double z0, z1;
    orc_word_to_normals(stream_next(st), &z0, &z1);
    return z0;
}

/* -------------------------------------------------- symmetric factor (U,S) */

/* Stand-in for np.linalg.svd of a symmetric PSD block (ref :145, :560, :803):
 * Jacobi eigen-decomposition in round-robin (parallel) order, eigenvalues by descending magnitude,
 * S = |lambda|, each eigenvector's largest-magnitude component made positive. */
void orc_sym_factor(int n, const double *a_in, double *U, double *S)
{
    double *a = (double *)malloc(sizeof(double) * n * n);
    double *v = (double *)malloc(sizeof(double) * n * n);
    memcpy(a, a_in, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) v[i * n + j] = (i == j) ? 1.0 : 0.0;
    /* parallel (round-robin) ordering, the same phases and fused operations as the CUDA jacobi_block:
     * a sweep is m-1 rounds of the circle tournament; the pairs of a round are disjoint, so the column phase
     * (A J, V J) and the row phase (J^T A) touch every entry once and the execution order is irrelevant */
    int m = (n + 1) & ~1, npair = m >> 1;
    int *pp = (int *)malloc(sizeof(int) * npair), *qq = (int *)malloc(sizeof(int) * npair);
    int *rot = (int *)malloc(sizeof(int) * npair);
    double *sns = (double *)malloc(sizeof(double) * npair), *taus = (double *)malloc(sizeof(double) * npair);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += fabs(a[p * n + q]);
        if (off == 0.0) break;
        for (int step = 0; step < m - 1; ++step) {
            for (int k = 0; k < npair; ++k) {
                int i = (k == 0) ? m - 1 : (step + k) % (m - 1);
                int j = (k == 0) ? step : (step + m - 1 - k) % (m - 1);
                int p = i < j ? i : j, q = i < j ? j : i;
                pp[k] = p; qq[k] = q; rot[k] = 0; sns[k] = 0.0; taus[k] = 0.0;
                if (q >= n) continue;
                double apq = a[p * n + q], app = a[p * n + p], aqq = a[q * n + q];
                double g = 100.0 * fabs(apq);
                if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
                    rot[k] = 2;
                } else if (apq != 0.0) {
                    double h = aqq - app, t;
                    if (fabs(h) + g == fabs(h)) {
                        t = apq / h;
                    } else {
                        double theta = 0.5 * h / apq;
                        t = 1.0 / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                        if (theta < 0.0) t = -t;
                    }
                    double c = 1.0 / sqrt(fma(t, t, 1.0));
                    sns[k] = t * c;
                    taus[k] = sns[k] / (1.0 + c);
                    rot[k] = 1;
                }
            }
            for (int k = 0; k < npair; ++k) { /* column phase */
                if (rot[k] != 1) continue;
                int p = pp[k], q = qq[k];
                double sn = sns[k], tau = taus[k];
                for (int r = 0; r < n; ++r) {
                    double g = a[r * n + p], h = a[r * n + q];
                    a[r * n + p] = fma(-sn, fma(g, tau, h), g);
                    a[r * n + q] = fma(sn, fma(-h, tau, g), h);
                    g = v[r * n + p]; h = v[r * n + q];
                    v[r * n + p] = fma(-sn, fma(g, tau, h), g);
                    v[r * n + q] = fma(sn, fma(-h, tau, g), h);
                }
            }
            for (int k = 0; k < npair; ++k) { /* row phase */
                if (rot[k] != 1) continue;
                int p = pp[k], q = qq[k];
                double sn = sns[k], tau = taus[k];
                for (int c = 0; c < n; ++c) {
                    double g = a[p * n + c], h = a[q * n + c];
                    a[p * n + c] = fma(-sn, fma(g, tau, h), g);
                    a[q * n + c] = fma(sn, fma(-h, tau, g), h);
                }
            }
            for (int k = 0; k < npair; ++k)
                if (rot[k] != 0) { a[pp[k] * n + qq[k]] = 0.0; a[qq[k] * n + pp[k]] = 0.0; }
        }
    }
    free(pp); free(qq); free(rot); free(sns); free(taus);
    /* order by descending |lambda|, ties by original index (stable selection) */
    int *ord = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    for (int i = 0; i < n; ++i) {
        int best = i;
        for (int j = i + 1; j < n; ++j)
            if (fabs(a[ord[j] * n + ord[j]]) > fabs(a[ord[best] * n + ord[best]])) best = j;
        int tmp = ord[best];
        for (int j = best; j > i; --j) ord[j] = ord[j - 1];
        ord[i] = tmp;
    }
    for (int k = 0; k < n; ++k) {
        int src = ord[k];
        S[k] = fabs(a[src * n + src]);
        int big = 0;
        for (int r = 1; r < n; ++r)
            if (fabs(v[r * n + src]) > fabs(v[big * n + src])) big = r;
        double sg = (v[big * n + src] < 0.0) ? -1.0 : 1.0;
        for (int r = 0; r < n; ++r) U[r * n + k] = sg * v[r * n + src];
    }
    free(ord); free(a); free(v);
}

/* ref :699-720 temperatureLadder */
void orc_temperature_ladder(int ndim, int ntemps, double tmin, double tmax, double *ladder)
{
    if (ntemps > 1) {
        double tstep;
        if (!(tmax > 0.0)) tstep = 1.0 + sqrt(2.0 / ndim);
        else tstep = exp(log(tmax / tmin) / (ntemps - 1));
        for (int i = 0; i < ntemps; ++i) ladder[i] = tmin * pow(tstep, (double)i);
    } else {
        ladder[0] = 1.0;
    }
}

/* ------------------------------------------------------------- sampler -- */

struct orc_sampler {
    orc_config c;
    int d, W, T, ngroups, njumps;
    double *ladder, *mh_temp;
    int *goff, *gidx;
    int ncycle, cyc_jump[ORC_MAX_CYCLE], cyc_w[ORC_MAX_CYCLE], de_in_cycle;
    double *logl_par, *logp_par;
    /* chain state [T][W][d], [T][W] */
    double *x, *lnl, *lp;
    /* adaptive state, pooled over cold walkers */
    double *cov, *mu, *m2;
    int64_t nsamp;
    double *U, *S;     /* concatenated per group */
    int *uoff, *soff;
    double *am, *de;   /* [cU][W][d], [burn][W][d] */
    /* records */
    int ntr;
    int64_t rows;
    double *rec_x, *rec_lnl, *rec_lnp;
    /* counters */
    int64_t *prop, *acc, *swap_acc, swap_proposed;
    int64_t iter;
    /* injected factors */
    int ninj, inj_next;
    double *inj_U, *inj_S;
    /* trace */
    uint8_t *trace; int64_t trace_iters, trace_pos;
    int16_t *swapmaps; int64_t swap_events, swap_pos;
    int64_t adapt_done, de_done; /* boundaries whose maintenance already ran */
    /* ladder sharding */
    int Tg, sharded, pending_swap, swept;
    int *smap;          /* [T][W] source code of every local position: 0..T-1 local rung, T carry_in */
    int *carry_code;    /* [W] */
    double *carry_L;    /* [W] */
    const double *carry_in;
};

static int usize(const orc_sampler *s) { return s->uoff[s->ngroups]; }

static void factor_groups(orc_sampler *s)
{
    /* ref :138-145, :797-803: per-group sub-covariance then factorise */
    if (s->ninj > 0 && s->inj_next < s->ninj) {
        memcpy(s->U, s->inj_U + (size_t)s->inj_next * usize(s), sizeof(double) * usize(s));
        memcpy(s->S, s->inj_S + (size_t)s->inj_next * s->soff[s->ngroups],
               sizeof(double) * s->soff[s->ngroups]);
        s->inj_next++;
        return;
    }
    for (int g = 0; g < s->ngroups; ++g) {
        int dg = s->goff[g + 1] - s->goff[g];
        const int *gi = s->gidx + s->goff[g];
        double *sub = (double *)malloc(sizeof(double) * dg * dg);
        for (int ii = 0; ii < dg; ++ii)
            for (int jj = 0; jj < dg; ++jj) sub[ii * dg + jj] = s->cov[gi[ii] * s->d + gi[jj]];
        orc_sym_factor(dg, sub, s->U + s->uoff[g], s->S + s->soff[g]);
        free(sub);
    }
}

orc_sampler *orc_create(const orc_config *cfg)
{
    orc_sampler *s = (orc_sampler *)calloc(1, sizeof(orc_sampler));
    s->c = *cfg;
    int d = s->d = cfg->ndim, W = s->W = cfg->nwalkers, T = s->T = cfg->ntemps;
    s->ladder = (double *)malloc(sizeof(double) * T);
    s->mh_temp = (double *)malloc(sizeof(double) * T);
    memcpy(s->ladder, cfg->ladder, sizeof(double) * T);
    memcpy(s->mh_temp, cfg->mh_temp ? cfg->mh_temp : cfg->ladder, sizeof(double) * T);
    /* ref :129-131 default = one group of all indices */
    if (cfg->ngroups <= 0 || !cfg->group_offsets) {
        s->ngroups = 1;
        s->goff = (int *)malloc(sizeof(int) * 2);
        s->goff[0] = 0; s->goff[1] = d;
        s->gidx = (int *)malloc(sizeof(int) * d);
        for (int i = 0; i < d; ++i) s->gidx[i] = i;
    } else {
        s->ngroups = cfg->ngroups;
        s->goff = (int *)malloc(sizeof(int) * (s->ngroups + 1));
        memcpy(s->goff, cfg->group_offsets, sizeof(int) * (s->ngroups + 1));
        int tot = s->goff[s->ngroups];
        s->gidx = (int *)malloc(sizeof(int) * tot);
        memcpy(s->gidx, cfg->group_indices, sizeof(int) * tot);
    }
    s->uoff = (int *)malloc(sizeof(int) * (s->ngroups + 1));
    s->soff = (int *)malloc(sizeof(int) * (s->ngroups + 1));
    s->uoff[0] = s->soff[0] = 0;
    for (int g = 0; g < s->ngroups; ++g) {
        int dg = s->goff[g + 1] - s->goff[g];
        s->uoff[g + 1] = s->uoff[g] + dg * dg;
        s->soff[g + 1] = s->soff[g] + dg;
    }
    s->ncycle = cfg->ncycle;
    s->njumps = 3;
    for (int i = 0; i < cfg->ncycle; ++i) {
        s->cyc_jump[i] = cfg->cycle_jump[i];
        s->cyc_w[i] = cfg->cycle_weight[i];
        if (cfg->cycle_jump[i] + 1 > s->njumps) s->njumps = cfg->cycle_jump[i] + 1;
    }
    s->de_in_cycle = 0;
    /* parameter blobs */
    if (cfg->logl_kind == ORC_LOGL_GAUSSIAN) {
        size_t n = (size_t)d + (size_t)d * d + 1;
        s->logl_par = (double *)malloc(sizeof(double) * n);
        memcpy(s->logl_par, cfg->logl_params, sizeof(double) * n);
    }
    if (cfg->logp_kind == ORC_LOGP_UNIFORM) {
        size_t n = 2 * (size_t)d + 2;
        s->logp_par = (double *)malloc(sizeof(double) * n);
        memcpy(s->logp_par, cfg->logp_params, sizeof(double) * n);
    }
    size_t C = (size_t)T * W;
    s->x = (double *)calloc(C * d, sizeof(double));
    s->lnl = (double *)calloc(C, sizeof(double));
    s->lp = (double *)calloc(C, sizeof(double));
    s->cov = (double *)malloc(sizeof(double) * d * d);
    memcpy(s->cov, cfg->cov, sizeof(double) * d * d);
    s->mu = (double *)calloc(d, sizeof(double));       /* ref :147-148 */
    s->m2 = (double *)calloc((size_t)d * d, sizeof(double));
    s->U = (double *)calloc(usize(s), sizeof(double));
    s->S = (double *)calloc(s->soff[s->ngroups], sizeof(double));
    factor_groups(s);                                   /* ref :138-145 */
    s->am = (double *)calloc((size_t)cfg->cov_update * W * d, sizeof(double)); /* ref :220 */
    s->de = (double *)calloc((size_t)cfg->burn * W * d, sizeof(double));       /* ref :221 */
    s->ntr = cfg->record_hot ? T : 1;
    s->rec_x = (double *)calloc((size_t)cfg->max_rows * s->ntr * W * d, sizeof(double));
    s->rec_lnl = (double *)calloc((size_t)cfg->max_rows * s->ntr * W, sizeof(double));
    s->rec_lnp = (double *)calloc((size_t)cfg->max_rows * s->ntr * W, sizeof(double));
    s->prop = (int64_t *)calloc(C * s->njumps, sizeof(int64_t));
    s->acc = (int64_t *)calloc(C * s->njumps, sizeof(int64_t));
    s->swap_acc = (int64_t *)calloc(C, sizeof(int64_t));
    s->adapt_done = s->de_done = -1;
    s->Tg = cfg->ntemps_global > 0 ? cfg->ntemps_global : T;
    s->sharded = s->Tg > T;
    s->smap = (int *)calloc(C, sizeof(int));
    s->carry_code = (int *)calloc(W, sizeof(int));
    s->carry_L = (double *)calloc(W, sizeof(double));
    return s;
}

void orc_destroy(orc_sampler *s)
{
    if (!s) return;
    free(s->ladder); free(s->mh_temp); free(s->goff); free(s->gidx); free(s->uoff); free(s->soff);
    free(s->logl_par); free(s->logp_par); free(s->x); free(s->lnl); free(s->lp); free(s->cov);
    free(s->mu); free(s->m2); free(s->U); free(s->S); free(s->am); free(s->de); free(s->rec_x);
    free(s->rec_lnl); free(s->rec_lnp); free(s->prop); free(s->acc); free(s->swap_acc);
    free(s->inj_U); free(s->inj_S);
    free(s->smap); free(s->carry_code); free(s->carry_L);
    free(s);
}

/* built-in targets; formulas follow the reference's example likelihoods */
static double eval_logp(const orc_sampler *s, const double *x)
{
    switch (s->c.logp_kind) {
    case ORC_LOGP_UNIFORM: { /* examples/simple.py:38-44, curved_likelihood.ipynb lnpriorfn */
        const double *lo = s->logp_par, *hi = lo + s->d;
        double inside = hi[s->d], inclusive = hi[s->d + 1];
        for (int k = 0; k < s->d; ++k) {
            if (inclusive != 0.0) { if (!(lo[k] <= x[k] && hi[k] >= x[k])) return -INFINITY; }
            else { if (!(lo[k] < x[k] && hi[k] > x[k])) return -INFINITY; }
        }
        return inside;
    }
    case ORC_LOGP_FLAT: return 0.0;
    default: return s->c.ext_logp(x, s->d, s->c.user);
    }
}

static double eval_logl(const orc_sampler *s, const double *x)
{
    int d = s->d;
    switch (s->c.logl_kind) {
    case ORC_LOGL_GAUSSIAN: { /* examples/simple.py:34-36: -dot(diff, dot(icov, diff))/2 */
        const double *mu = s->logl_par, *A = mu + d;
        double offset = A[(size_t)d * d], quad = 0.0;
        for (int i = 0; i < d; ++i) {
            double row = 0.0;
            for (int j = 0; j < d; ++j) row += A[i * d + j] * (x[j] - mu[j]);
            quad += (x[i] - mu[i]) * row;
        }
        return -quad / 2.0 + offset;
    }
    case ORC_LOGL_CURVED: { /* curved_likelihood.ipynb lnlikefn, summed over 2-D blocks */
        double tot = 0.0;
        for (int b = 0; b + 1 < d; b += 2) {
            double a = x[b], y = x[b + 1];
            double t0 = 9.0 + 4.0 * a * a + 9.0 * y;
            double ll = exp(-a * a - t0 * t0) + 0.5 * exp(-8.0 * a * a - 8.0 * (y - 2.0) * (y - 2.0));
            tot += log(ll);
        }
        return tot;
    }
    case ORC_LOGL_ROSENBROCK: {
        double tot = 0.0;
        for (int i = 0; i + 1 < d; ++i) {
            double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
            tot -= 100.0 * a * a + b * b;
        }
        return tot / 20.0;
    }
    default: return s->c.ext_logl(x, d, s->c.user);
    }
}

static void record_row(orc_sampler *s, int64_t row, int w)
{
    if (row >= s->c.max_rows) return;
    int d = s->d, W = s->W;
    for (int t = 0; t < s->ntr; ++t) {
        size_t ch = (size_t)t * W + w;
        size_t r = ((size_t)row * s->ntr + t) * W + w;
        memcpy(s->rec_x + r * d, s->x + ch * d, sizeof(double) * d);
        s->rec_lnl[r] = s->lnl[ch];
        s->rec_lnp[r] = 1.0 / s->mh_temp[t] * s->lnl[ch] + s->lp[ch];
    }
}

/* ref :321-339 updateChains (buffer + thinned record; file output lives in the host shim) */
static void update_chains(orc_sampler *s, int64_t iter, int w)
{
    int d = s->d, W = s->W;
    if (s->c.temp_offset == 0)
        memcpy(s->am + ((size_t)(iter % s->c.cov_update) * W + w) * d, s->x + (size_t)w * d,
               sizeof(double) * d);
    if (iter % s->c.thin == 0) record_row(s, iter / s->c.thin, w);
}

int orc_set_state(orc_sampler *s, const double *x0)
{
    int d = s->d, W = s->W, T = s->T;
    memcpy(s->x, x0, sizeof(double) * (size_t)T * W * d);
    /* ref :478-487 */
    for (int t = 0; t < T; ++t)
        for (int w = 0; w < W; ++w) {
            size_t ch = (size_t)t * W + w;
            double lp = eval_logp(s, s->x + ch * d);
            s->lp[ch] = lp;
            s->lnl[ch] = (lp == -INFINITY) ? -INFINITY : eval_logl(s, s->x + ch * d);
        }
    s->iter = 0;
    s->rows = 1;
    for (int w = 0; w < W; ++w) update_chains(s, 0, w); /* ref :491 */
    return 0;
}

/* ref :769-803 _updateRecursive; pooled sample order (slot, walker) */
static void update_recursive(orc_sampler *s, int64_t iter, int64_t mem)
{
    int d = s->d, W = s->W;
    int64_t it = iter - mem;
    if (it == 0) {
        memset(s->m2, 0, sizeof(double) * d * d);
        memset(s->mu, 0, sizeof(double) * d);
    }
    int64_t n_end = 0;
#pragma omp parallel num_threads(s->c.nthreads > 0 ? s->c.nthreads : 1)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
        int tid = 0, nth = 1;
#endif
        /* every thread replays the (cheap) mean recursion and owns a slice of M2's rows, so
         * the result is bit-identical to the sequential loop */
        double *mu = (double *)malloc(sizeof(double) * d);
        double *diff = (double *)malloc(sizeof(double) * d);
        memcpy(mu, s->mu, sizeof(double) * d);
        int r0 = (int)((int64_t)d * tid / nth), r1 = (int)((int64_t)d * (tid + 1) / nth);
        int64_t n = it * W;
        for (int64_t ii = 0; ii < mem; ++ii)
            for (int w = 0; w < W; ++w) {
                const double *row = s->am + ((size_t)ii * W + w) * d;
                n += 1;
                for (int jj = 0; jj < d; ++jj) {
                    diff[jj] = row[jj] - mu[jj];
                    mu[jj] += diff[jj] / (double)n;
                }
                for (int i = r0; i < r1; ++i)
                    for (int j = 0; j < d; ++j) s->m2[i * d + j] += diff[i] * (row[j] - mu[j]);
            }
#pragma omp barrier
        if (tid == 0) {
            memcpy(s->mu, mu, sizeof(double) * d);
            n_end = n;
        }
        free(mu); free(diff);
    }
    s->nsamp = n_end;
    for (int i = 0; i < d * d; ++i) s->cov[i] = s->m2[i] / (double)(n_end - 1); /* ref :794 */
    factor_groups(s);
}

/* ref :806-817 + shift_array :27-37: shift left by len(AM buffer), append AM buffer */
static int update_de_buffer(orc_sampler *s)
{
    size_t rowlen = (size_t)s->W * s->d;
    int64_t cu = s->c.cov_update, burn = s->c.burn;
    if (cu > burn) return -2; /* the reference raises a broadcast ValueError here */
    memmove(s->de, s->de + (size_t)cu * rowlen, sizeof(double) * (size_t)(burn - cu) * rowlen);
    memcpy(s->de + (size_t)(burn - cu) * rowlen, s->am, sizeof(double) * (size_t)cu * rowlen);
    return 0;
}

/* one Metropolis-Hastings update of chain (t, w); ref :601-622, _jump :1048-1067,
 * SCAM :820-876, AM :879-933, DE :936-985 */
static void mh_step(orc_sampler *s, int64_t iter, int w, int t, double *q, double *y, int64_t trace_pos)
{
    int d = s->d, W = s->W;
    size_t ch = (size_t)t * W + w;
    double *x = s->x + ch * d;
    double temp = s->mh_temp[t];
    double beta = 1.0 / temp;
    orc_stream st;
    stream_init(&st, s->c.seed, ORC_PURPOSE_MH, (uint64_t)iter, (uint32_t)(s->c.walker_offset + w),
                (uint32_t)(s->c.temp_offset + t));
    /* _jump: ind = integers(0, len(propCycle)); propCycle is weight-replicated :1007-1008 */
    int total = 0;
    for (int i = 0; i < s->ncycle; ++i) total += s->cyc_w[i];
    int ind = (int)draw_int(&st, (uint64_t)total);
    int jump = s->cyc_jump[s->ncycle - 1];
    for (int i = 0, cum = 0; i < s->ncycle; ++i) {
        cum += s->cyc_w[i];
        if (ind < cum) { jump = s->cyc_jump[i]; break; }
    }
    memcpy(q, x, sizeof(double) * d);
    double qxy = 0.0;
    if (jump == ORC_JUMP_SCAM || jump == ORC_JUMP_AM) {
        int g = (int)draw_int(&st, (uint64_t)s->ngroups);
        int dg = s->goff[g + 1] - s->goff[g];
        const int *gi = s->gidx + s->goff[g];
        const double *U = s->U + s->uoff[g], *S = s->S + s->soff[g];
        double prob = draw_unit(&st);
        double scale;
        if (prob > 0.97) scale = 10.0;
        else if (prob > 0.9) scale = 0.2;
        else scale = 1.0;
        if (temp <= 100.0) scale *= sqrt(temp);
        if (jump == ORC_JUMP_SCAM) {
            int k = (int)draw_int(&st, (uint64_t)dg); /* np.unique of one index: neff = 1 */
            double cd = 2.4 / sqrt(2.0 * 1.0) * scale;
            double z = draw_normal(&st);
            double coef = z * cd * sqrt(S[k]);
            for (int i = 0; i < dg; ++i) q[gi[i]] += coef * U[i * dg + k];
        } else {
            double cd = 2.4 / sqrt(2.0 * dg) * scale;
            for (int j = 0; j < dg; ++j) {
                double acc = 0.0;
                for (int i = 0; i < dg; ++i) acc += U[i * dg + j] * x[gi[i]];
                y[j] = acc; /* y = U^T x */
            }
            for (int j = 0; j < dg; j += 2) {
                double z0, z1;
                orc_word_to_normals(stream_next(&st), &z0, &z1);
                y[j] = y[j] + z0 * cd * sqrt(S[j]);
                if (j + 1 < dg) y[j + 1] = y[j + 1] + z1 * cd * sqrt(S[j + 1]);
            }
            for (int i = 0; i < dg; ++i) {
                double acc = 0.0;
                for (int j = 0; j < dg; ++j) acc += U[i * dg + j] * y[j];
                q[gi[i]] = acc; /* q = U y */
            }
        }
    } else if (jump == ORC_JUMP_DE) {
        int g = (int)draw_int(&st, (uint64_t)s->ngroups);
        int dg = s->goff[g + 1] - s->goff[g];
        const int *gi = s->gidx + s->goff[g];
        uint64_t bufsize = (uint64_t)s->c.burn * (uint64_t)W;
        uint64_t mm = draw_int(&st, bufsize);
        uint64_t nn = draw_int(&st, bufsize);
        while (mm == nn) nn = draw_int(&st, bufsize);
        double prob = draw_unit(&st);
        double scale;
        if (prob > 0.5) scale = 1.0;
        else scale = draw_unit(&st) * 2.4 / sqrt(2.0 * dg) * sqrt(1.0 / beta);
        const double *bm = s->de + mm * d, *bn = s->de + nn * d;
        for (int i = 0; i < dg; ++i) {
            double sigma = bm[gi[i]] - bn[gi[i]];
            q[gi[i]] += scale * sigma;
        }
    } else if (jump == ORC_JUMP_PRIOR) {
        /* draw from the uniform prior box, one stream.random() per parameter (the UniformJump plugin of ref
         * tests/test_simple.py:44-62 with the sampler's stream); symmetric, qxy = 0 */
        const double *lo = s->logp_par, *hi = lo + d;
        for (int k = 0; k < d; ++k) q[k] = lo[k] + (hi[k] - lo[k]) * draw_unit(&st);
    } else {
        s->c.ext_jump(jump - ORC_JUMP_EXT0, x, d, iter, beta, s->c.walker_offset + w,
                      s->c.temp_offset + t, q, &qxy, s->c.user);
    }
    s->prop[ch * s->njumps + jump] += 1; /* ref :602 */
    /* ref :605-612 */
    double lpnew = eval_logp(s, q), newlnlike = 0.0, newlnprob;
    if (lpnew == -INFINITY) newlnprob = -INFINITY;
    else {
        newlnlike = eval_logl(s, q);
        newlnprob = 1.0 / temp * newlnlike + lpnew;
    }
    double lnprob0 = 1.0 / temp * s->lnl[ch] + s->lp[ch];
    /* ref :615-622 */
    double diff = newlnprob - lnprob0 + qxy;
    double u = draw_unit(&st);
    int accepted = 0;
    if (diff > log(u)) {
        memcpy(x, q, sizeof(double) * d);
        s->lnl[ch] = newlnlike;
        s->lp[ch] = lpnew;
        s->acc[ch * s->njumps + jump] += 1;
        accepted = 1;
    }
    if (s->trace && trace_pos < s->trace_iters)
        s->trace[((size_t)trace_pos * s->T + t) * W + w] = (uint8_t)(jump | (accepted << 7));
}

/* ref :631-697 PTswap for the ladder of walker w */
static void pt_swap(orc_sampler *s, int64_t iter, int w, int *map, double *tmpx, double *tmpl,
                    int64_t swap_pos)
{
    int d = s->d, W = s->W, T = s->T;
    const double *Ts = s->ladder;
    orc_stream st;
    stream_init(&st, s->c.seed, ORC_PURPOSE_SWAP, (uint64_t)iter, (uint32_t)(s->c.walker_offset + w), 0);
    for (int j = 0; j < T; ++j) map[j] = j;
    for (int sc = T - 2; sc >= 0; --sc) {
        double La = s->lnl[(size_t)map[sc] * W + w], Lb = s->lnl[(size_t)map[sc + 1] * W + w];
        double lar = -La / Ts[sc];
        lar += -Lb / Ts[sc + 1];
        lar += Lb / Ts[sc];
        lar += La / Ts[sc + 1];
        double ratio = exp(lar);
        double u = draw_unit(&st);
        if (u <= ratio) {
            int tmp = map[sc]; map[sc] = map[sc + 1]; map[sc + 1] = tmp;
            s->swap_acc[(size_t)sc * W + w] += 1;
        }
    }
    for (int j = 0; j < T; ++j) {
        size_t src = (size_t)map[j] * W + w;
        memcpy(tmpx + (size_t)j * d, s->x + src * d, sizeof(double) * d);
        tmpl[j] = s->lnl[src];
        tmpl[T + j] = s->lp[src];
    }
    for (int j = 0; j < T; ++j) {
        size_t dst = (size_t)j * W + w;
        memcpy(s->x + dst * d, tmpx + (size_t)j * d, sizeof(double) * d);
        s->lnl[dst] = tmpl[j];
        s->lp[dst] = tmpl[T + j]; /* lnprob = lnlike/temp + logp(p0) is re-derived on use, ref :695 */
    }
    if (s->swapmaps && swap_pos < s->swap_events)
        for (int j = 0; j < T; ++j) s->swapmaps[((size_t)swap_pos * W + w) * T + j] = (int16_t)map[j];
}

static int64_t next_boundary(const orc_sampler *s, int64_t it)
{
    /* smallest b >= it such that iteration b+1 starts with a covariance or DE update */
    int64_t cu = s->c.cov_update, burn = s->c.burn;
    int64_t b1 = ((it + cu - 1) / cu) * cu, b2 = ((it + burn - 1) / burn) * burn;
    return b1 < b2 ? b1 : b2;
}

/* maintenance at the start of iteration it0: ref :545-560 covariance update, :563-585 DE buffer
 * update + DE joins the cycle.  Idempotent per boundary so that orc_maintain can run it early. */
static int maintenance(orc_sampler *s, int64_t it0)
{
    int64_t b = it0 - 1;
    if (b % s->c.cov_update == 0 && b != 0 && s->c.temp_offset == 0 && s->adapt_done != b) {
        update_recursive(s, b, s->c.cov_update);
        s->adapt_done = b;
    }
    if (b % s->c.burn == 0 && b != 0 && s->de_done != b) {
        int rc = update_de_buffer(s);
        if (rc) return rc;
        s->de_done = b;
        if (!s->de_in_cycle && s->c.de_weight > 0) {
            s->cyc_jump[s->ncycle] = ORC_JUMP_DE;
            s->cyc_w[s->ncycle] = s->c.de_weight;
            s->ncycle++;
            s->de_in_cycle = 1;
        }
    }
    return 0;
}

int orc_maintain(orc_sampler *s) { return s->pending_swap ? -4 : maintenance(s, s->iter + 1); }

/* Walker sharding: the covariance update split around an exchange of batch moments, the multi-process
 * form of rank 0's send(cov) (ref :545-560).  begin: batch = {n, mean[d], M2c[d*d]} of this shard's AM
 * ring if an update is due at the current iteration (returns 1, else 0); finish: Chan-merge a (pooled)
 * batch into the running moments, refresh cov and the factor. */
int orc_adapt_begin(orc_sampler *s, double *batch)
{
    int d = s->d, W = s->W;
    int64_t b = s->iter, cu = s->c.cov_update;
    if (b == 0 || b % cu != 0 || s->adapt_done == b) return 0;
    double n = (double)cu * W;
    double *mean = batch + 1, *m2 = batch + 1 + d;
    memset(batch, 0, sizeof(double) * (size_t)(1 + d + d * d));
    for (int64_t i = 0; i < cu * W; ++i)
        for (int k = 0; k < d; ++k) mean[k] += s->am[(size_t)i * d + k];
    for (int k = 0; k < d; ++k) mean[k] /= n;
    for (int64_t i = 0; i < cu * W; ++i) {
        const double *row = s->am + (size_t)i * d;
        for (int a = 0; a < d; ++a)
            for (int c = 0; c < d; ++c) m2[a * d + c] += (row[a] - mean[a]) * (row[c] - mean[c]);
    }
    batch[0] = n;
    return 1;
}

int orc_adapt_finish(orc_sampler *s, const double *batch)
{
    int d = s->d;
    int64_t b = s->iter, cu = s->c.cov_update;
    if (b == 0 || b % cu != 0 || s->adapt_done == b) return -4;
    double na = (b - cu == 0) ? 0.0 : (double)s->nsamp, nb = batch[0], ntot = na + nb;
    const double *mb = batch + 1, *m2b = batch + 1 + d;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double mai = na > 0 ? s->mu[i] : 0.0, maj = na > 0 ? s->mu[j] : 0.0;
            double m2a = na > 0 ? s->m2[i * d + j] : 0.0;
            s->m2[i * d + j] = m2a + m2b[i * d + j] + (mb[i] - mai) * (mb[j] - maj) * (na * nb / ntot);
        }
    for (int k = 0; k < d; ++k) {
        double ma = na > 0 ? s->mu[k] : 0.0;
        s->mu[k] = ma + (mb[k] - ma) * (nb / ntot);
    }
    s->nsamp = (int64_t)ntot;
    for (int i = 0; i < d * d; ++i) s->cov[i] = s->m2[i] / (ntot - 1.0);
    factor_groups(s);
    s->adapt_done = b;
    return 0;
}

/* ref :495-528 driver loop and :530-629 PTMCMCOneStep */
int orc_run(orc_sampler *s, int64_t niter)
{
    int d = s->d, W = s->W, T = s->T;
    int64_t end = s->iter + niter;
    int nth = s->c.nthreads > 0 ? s->c.nthreads : 1;
    if (s->c.logl_kind == ORC_LOGL_EXTERNAL || s->c.logp_kind == ORC_LOGP_EXTERNAL || s->njumps > ORC_JUMP_EXT0) nth = 1;
    if (s->pending_swap) return -4; /* a sharded swap must be completed first */
    while (s->iter < end) {
        int64_t it0 = s->iter + 1;
        int mrc = maintenance(s, it0);
        if (mrc) return mrc;
        int64_t seg_end = next_boundary(s, it0);
        if (seg_end > end) seg_end = end;
        if (s->sharded) { /* stop at the swap iteration: the exchange is driven from outside */
            int64_t nsw = ((it0 + s->c.tskip - 1) / s->c.tskip) * s->c.tskip;
            if (nsw < seg_end) seg_end = nsw;
        }
        int64_t trace_base = s->trace_pos, swap_base = s->swap_pos;
#pragma omp parallel num_threads(nth)
        {
            double *q = (double *)malloc(sizeof(double) * d);
            double *y = (double *)malloc(sizeof(double) * d);
            int *map = (int *)malloc(sizeof(int) * T);
            double *tmpx = (double *)malloc(sizeof(double) * (size_t)T * d);
            double *tmpl = (double *)malloc(sizeof(double) * 2 * T);
#pragma omp for schedule(static)
            for (int w = 0; w < W; ++w) {
                int64_t nsw = 0;
                for (int64_t it = it0; it <= seg_end; ++it) {
                    for (int t = 0; t < T; ++t) mh_step(s, it, w, t, q, y, trace_base + (it - it0));
                    if (s->sharded && it % s->c.tskip == 0) continue; /* swap + updateChains: orc_swap_* */
                    if (it % s->c.tskip == 0 && T > 1) { /* ref :624-625 */
                        pt_swap(s, it, w, map, tmpx, tmpl, swap_base + nsw);
                        nsw++;
                    }
                    update_chains(s, it, w);                /* ref :627 */
                }
            }
            free(q); free(y); free(map); free(tmpx); free(tmpl);
        }
        for (int64_t it = it0; it <= seg_end; ++it) {
            if (s->sharded) continue;
            if (it % s->c.tskip == 0 && T > 1) { s->swap_proposed++; s->swap_pos++; }
            if (it % s->c.thin == 0 && it / s->c.thin + 1 > s->rows) s->rows = it / s->c.thin + 1;
        }
        s->trace_pos += seg_end - it0 + 1;
        s->iter = seg_end;
        if (s->sharded) {
            for (int64_t it = it0; it <= seg_end; ++it)
                if (it % s->c.tskip != 0 && it % s->c.thin == 0 && it / s->c.thin + 1 > s->rows)
                    s->rows = it / s->c.thin + 1;
            if (seg_end % s->c.tskip == 0) {
                s->pending_swap = 1;
                s->swept = 0;
                if (seg_end < end) return -4;
            }
        }
    }
    if (s->rows > s->c.max_rows) s->rows = s->c.max_rows;
    return 0;
}

/* ------------------------------------------------ ladder-sharded swap -- */

int64_t orc_swap_msg_doubles(const orc_sampler *s) { return (int64_t)(s->d + 3) * s->W; }
int orc_swap_pending(const orc_sampler *s) { return s->pending_swap; }
double *orc_am_ring(orc_sampler *s) { return s->am; }

static void pack_state(const orc_sampler *s, double *msg, int w, const double *x, double lnl, double lp,
                       double origin)
{
    int d = s->d, W = s->W;
    for (int k = 0; k < d; ++k) msg[(size_t)k * W + w] = x[k];
    msg[(size_t)d * W + w] = lnl;
    msg[(size_t)(d + 1) * W + w] = lp;
    msg[(size_t)(d + 2) * W + w] = origin;
}

void orc_swap_pack_top(const orc_sampler *s, double *msg)
{
    int d = s->d, W = s->W, T = s->T;
    for (int w = 0; w < W; ++w) {
        size_t ch = (size_t)(T - 1) * W + w;
        pack_state(s, msg, w, s->x + ch * d, s->lnl[ch], s->lp[ch], (double)(s->c.temp_offset + T - 1));
    }
}

/* acceptance of the pair (lower rung at Ta with lnL La, upper rung at Tb with lnL Lb), ref :673-679 */
static int swap_accept(double La, double Lb, double Ta, double Tb, double u)
{
    double lar = -La / Ta;
    lar += -Lb / Tb;
    lar += Lb / Ta;
    lar += La / Tb;
    return u <= exp(lar);
}

static double swap_uniform(const orc_sampler *s, int64_t iter, int w, int sc_global)
{
    /* the sweep draws one uniform per pair, hottest pair first: pair sc uses word Tg-2-sc */
    return orc_word_to_unit(orc_draw_word(s->c.seed, ORC_PURPOSE_SWAP, (uint64_t)iter,
                                          (uint32_t)(s->c.walker_offset + w), 0, (uint32_t)(s->Tg - 2 - sc_global)));
}

int orc_swap_sweep(orc_sampler *s, const double *carry_in, double *carry_out)
{
    int d = s->d, W = s->W, T = s->T, off = s->c.temp_offset;
    if (!s->pending_swap || s->swept) return -4;
    int hottest = (off + T == s->Tg), coldest = (off == 0);
    if ((carry_in == NULL) != hottest || (carry_out == NULL) != coldest) return -1;
    int64_t iter = s->iter;
    for (int w = 0; w < W; ++w) {
        int carry = T - 1;
        double Lcarry = s->lnl[(size_t)(T - 1) * W + w];
        if (carry_in) { /* boundary pair: our top rung against the carry of the hotter shard */
            double La = Lcarry, Lb = carry_in[(size_t)d * W + w];
            if (swap_accept(La, Lb, s->ladder[T - 1], s->c.ladder_above, swap_uniform(s, iter, w, off + T - 1))) {
                s->swap_acc[(size_t)(T - 1) * W + w] += 1;
                carry = T; /* the foreign state keeps travelling down */
                Lcarry = Lb;
            }
        }
        for (int sc = T - 2; sc >= 0; --sc) {
            double La = s->lnl[(size_t)sc * W + w];
            if (swap_accept(La, Lcarry, s->ladder[sc], s->ladder[sc + 1], swap_uniform(s, iter, w, off + sc))) {
                s->smap[(size_t)(sc + 1) * W + w] = sc;
                s->swap_acc[(size_t)sc * W + w] += 1;
            } else {
                s->smap[(size_t)(sc + 1) * W + w] = carry;
                carry = sc;
                Lcarry = La;
            }
        }
        s->carry_code[w] = carry;
        s->carry_L[w] = Lcarry;
        if (carry_out) {
            if (carry == T)
                for (int k = 0; k < d + 3; ++k) carry_out[(size_t)k * W + w] = carry_in[(size_t)k * W + w];
            else {
                size_t ch = (size_t)carry * W + w;
                pack_state(s, carry_out, w, s->x + ch * d, s->lnl[ch], s->lp[ch], (double)(off + carry));
            }
        }
    }
    s->carry_in = carry_in;
    s->swept = 1;
    return 0;
}

int orc_swap_finish(orc_sampler *s, const double *below_top)
{
    int d = s->d, W = s->W, T = s->T, off = s->c.temp_offset;
    if (!s->pending_swap || !s->swept) return -4;
    if ((below_top == NULL) != (off == 0)) return -1;
    int64_t iter = s->iter;
    const double *cin = s->carry_in;
    double *nx = (double *)malloc(sizeof(double) * (size_t)T * d);
    double *nl = (double *)malloc(sizeof(double) * 2 * T);
    for (int w = 0; w < W; ++w) {
        /* position 0: boundary pair with the colder shard's top rung, decided identically there */
        int code0 = s->carry_code[w];
        if (below_top) {
            double La = below_top[(size_t)d * W + w], Lb = s->carry_L[w];
            if (swap_accept(La, Lb, s->c.ladder_below, s->ladder[0], swap_uniform(s, iter, w, off - 1))) code0 = T + 1;
        }
        s->smap[w] = code0;
        for (int t = 0; t < T; ++t) {
            int code = s->smap[(size_t)t * W + w];
            const double *msg = (code == T) ? cin : (code == T + 1) ? below_top : NULL;
            int origin;
            if (msg) {
                for (int k = 0; k < d; ++k) nx[(size_t)t * d + k] = msg[(size_t)k * W + w];
                nl[t] = msg[(size_t)d * W + w];
                nl[T + t] = msg[(size_t)(d + 1) * W + w];
                origin = (int)msg[(size_t)(d + 2) * W + w];
            } else {
                size_t ch = (size_t)code * W + w;
                memcpy(nx + (size_t)t * d, s->x + ch * d, sizeof(double) * d);
                nl[t] = s->lnl[ch];
                nl[T + t] = s->lp[ch];
                origin = off + code;
            }
            if (s->swapmaps && s->swap_pos < s->swap_events)
                s->swapmaps[((size_t)s->swap_pos * W + w) * T + t] = (int16_t)origin;
        }
        for (int t = 0; t < T; ++t) {
            size_t ch = (size_t)t * W + w;
            memcpy(s->x + ch * d, nx + (size_t)t * d, sizeof(double) * d);
            s->lnl[ch] = nl[t];
            s->lp[ch] = nl[T + t];
        }
        update_chains(s, iter, w); /* ref :627 */
    }
    free(nx); free(nl);
    if (iter % s->c.thin == 0 && iter / s->c.thin + 1 > s->rows) s->rows = iter / s->c.thin + 1;
    if (s->rows > s->c.max_rows) s->rows = s->c.max_rows;
    s->swap_proposed++;
    s->swap_pos++;
    s->pending_swap = 0;
    s->swept = 0;
    s->carry_in = NULL;
    return 0;
}

int64_t orc_iteration(const orc_sampler *s) { return s->iter; }
int64_t orc_rows(const orc_sampler *s) { return s->rows; }
int32_t orc_njumps(const orc_sampler *s) { return s->njumps; }

void orc_get_state(const orc_sampler *s, double *x, double *lnl, double *lnprior, double *lnprob)
{
    size_t C = (size_t)s->T * s->W;
    if (x) memcpy(x, s->x, sizeof(double) * C * s->d);
    if (lnl) memcpy(lnl, s->lnl, sizeof(double) * C);
    if (lnprior) memcpy(lnprior, s->lp, sizeof(double) * C);
    if (lnprob)
        for (int t = 0; t < s->T; ++t)
            for (int w = 0; w < s->W; ++w) {
                size_t ch = (size_t)t * s->W + w;
                lnprob[ch] = 1.0 / s->mh_temp[t] * s->lnl[ch] + s->lp[ch];
            }
}

void orc_get_chain(const orc_sampler *s, double *chain, double *lnl, double *lnprob)
{
    size_t n = (size_t)s->rows * s->ntr * s->W;
    if (chain) memcpy(chain, s->rec_x, sizeof(double) * n * s->d);
    if (lnl) memcpy(lnl, s->rec_lnl, sizeof(double) * n);
    if (lnprob) memcpy(lnprob, s->rec_lnp, sizeof(double) * n);
}

void orc_get_adapt(const orc_sampler *s, double *cov, double *mu, double *m2, int64_t *nsamp)
{
    if (cov) memcpy(cov, s->cov, sizeof(double) * s->d * s->d);
    if (mu) memcpy(mu, s->mu, sizeof(double) * s->d);
    if (m2) memcpy(m2, s->m2, sizeof(double) * s->d * s->d);
    if (nsamp) *nsamp = s->nsamp;
}

void orc_get_factor(const orc_sampler *s, double *U, double *S)
{
    if (U) memcpy(U, s->U, sizeof(double) * usize(s));
    if (S) memcpy(S, s->S, sizeof(double) * s->soff[s->ngroups]);
}

void orc_set_factor(orc_sampler *s, const double *U, const double *S)
{
    memcpy(s->U, U, sizeof(double) * usize(s));
    memcpy(s->S, S, sizeof(double) * s->soff[s->ngroups]);
}

void orc_inject_factors(orc_sampler *s, int n, const double *Us, const double *Ss)
{
    free(s->inj_U); free(s->inj_S);
    s->ninj = n; s->inj_next = 0;
    s->inj_U = (double *)malloc(sizeof(double) * (size_t)n * usize(s));
    s->inj_S = (double *)malloc(sizeof(double) * (size_t)n * s->soff[s->ngroups]);
    memcpy(s->inj_U, Us, sizeof(double) * (size_t)n * usize(s));
    memcpy(s->inj_S, Ss, sizeof(double) * (size_t)n * s->soff[s->ngroups]);
}

void orc_get_buffers(const orc_sampler *s, double *am, double *de)
{
    size_t rowlen = (size_t)s->W * s->d;
    if (am) memcpy(am, s->am, sizeof(double) * (size_t)s->c.cov_update * rowlen);
    if (de) memcpy(de, s->de, sizeof(double) * (size_t)s->c.burn * rowlen);
}

void orc_get_counters(const orc_sampler *s, int64_t *prop, int64_t *acc, int64_t *swap_acc,
                      int64_t *swap_proposed)
{
    size_t C = (size_t)s->T * s->W;
    if (prop) memcpy(prop, s->prop, sizeof(int64_t) * C * s->njumps);
    if (acc) memcpy(acc, s->acc, sizeof(int64_t) * C * s->njumps);
    if (swap_acc) memcpy(swap_acc, s->swap_acc, sizeof(int64_t) * C);
    if (swap_proposed) *swap_proposed = s->swap_proposed;
}

void orc_set_trace(orc_sampler *s, uint8_t *trace, int64_t trace_iters, int16_t *swapmaps,
                   int64_t swap_events)
{
    s->trace = trace; s->trace_iters = trace_iters; s->trace_pos = 0;
    s->swapmaps = swapmaps; s->swap_events = swap_events; s->swap_pos = 0;
}
