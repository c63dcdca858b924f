/*
 * ptmcmc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the parallel-tempering MCMC hot path of
 * nanograv/PTMCMCSampler (reference @ dd837f9), lifted from one chain per
 * temperature to W walkers x T temperatures, driven by a counter-based RNG
 * (Philox4x32-10) so that the CUDA engine and this file consume identical
 * random numbers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (ptmcmcsampler_b200/) never imports, links or calls it.
 *
 * Parity pin: the reference ships no golden vectors.  This
 * restatement is pinned instead against the UNMODIFIED reference run in the
 * build container with its `stream` attribute replaced by a generator that
 * returns this file's draws (tests/golden/make_golden.py); per-iteration
 * states, jump choices, accept flags, swap maps, covariance and DE buffers
 * are committed under tests/golden/ and checked by tests/test_oracle_golden.py.
 */
#ifndef PTMCMC_ORACLE_H
#define PTMCMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_JUMP_SCAM = 0, ORC_JUMP_AM = 1, ORC_JUMP_DE = 2, ORC_JUMP_PRIOR = 3, ORC_JUMP_EXT0 = 4 };
enum { ORC_LOGL_EXTERNAL = 0, ORC_LOGL_GAUSSIAN = 1, ORC_LOGL_CURVED = 2, ORC_LOGL_ROSENBROCK = 3 };
enum { ORC_LOGP_EXTERNAL = 0, ORC_LOGP_UNIFORM = 1, ORC_LOGP_FLAT = 2 };
enum { ORC_PURPOSE_MH = 0, ORC_PURPOSE_SWAP = 1 };

typedef double (*orc_logfn)(const double *x, int ndim, void *user);
/* external jump: fills q[ndim] and *qxy */
typedef void (*orc_jumpfn)(int ext_index, const double *x, int ndim, int64_t iter, double beta,
                           int walker, int temp, double *q, double *qxy, void *user);

typedef struct orc_config {
    int32_t ndim, nwalkers, ntemps;
    int32_t walker_offset;    /* global id of local walker 0 (RNG key)        */
    int32_t temp_offset;      /* global index of local temperature 0 (RNG key) */
    uint64_t seed;
    const double *ladder;     /* [ntemps] temperatures used by the swap        */
    const double *mh_temp;    /* [ntemps] temperatures used by the MH step     */
    const double *cov;        /* [ndim*ndim] initial proposal covariance       */
    int32_t ngroups;
    const int32_t *group_offsets; /* [ngroups+1] */
    const int32_t *group_indices; /* [group_offsets[ngroups]] */
    int32_t ncycle;           /* proposal-cycle segments before DE is added    */
    const int32_t *cycle_jump;    /* [ncycle] jump ids                         */
    const int32_t *cycle_weight;  /* [ncycle] integer weights                  */
    int32_t de_weight;        /* weight of the DE segment appended at burn+1   */
    int64_t cov_update, burn, tskip, thin;
    int32_t logl_kind; const double *logl_params;
    int32_t logp_kind; const double *logp_params;
    int32_t record_hot;       /* 0: record T=1 rung only, 1: all rungs         */
    int64_t max_rows;         /* record capacity (rows incl. row 0)            */
    int32_t nthreads;         /* OpenMP threads over walkers (1 = scalar port) */
    orc_logfn ext_logl; orc_logfn ext_logp; orc_jumpfn ext_jump; void *user;
    /* ladder sharding (one shard = ntemps contiguous rungs starting at temp_offset of a ladder of
     * ntemps_global rungs; 0 or ntemps = not sharded).  ladder_above / ladder_below are the
     * temperatures of the neighbouring shards' adjacent rungs. */
    int32_t ntemps_global;
    double ladder_above, ladder_below;
} orc_config;

typedef struct orc_sampler orc_sampler;

orc_sampler *orc_create(const orc_config *cfg);
void orc_destroy(orc_sampler *s);
/* x0 is [ntemps][nwalkers][ndim]; evaluates the initial point and records row 0 */
int orc_set_state(orc_sampler *s, const double *x0);
int orc_run(orc_sampler *s, int64_t niter);
int64_t orc_iteration(const orc_sampler *s);
void orc_get_state(const orc_sampler *s, double *x, double *lnl, double *lnprior, double *lnprob);
int64_t orc_rows(const orc_sampler *s);
/* chain[rows][ntr][W][d], lnl/lnprob[rows][ntr][W]; ntr = record_hot ? T : 1 */
void orc_get_chain(const orc_sampler *s, double *chain, double *lnl, double *lnprob);
void orc_get_adapt(const orc_sampler *s, double *cov, double *mu, double *m2, int64_t *nsamp);
/* U: concatenated row-major d_g x d_g blocks; S: concatenated d_g vectors */
void orc_get_factor(const orc_sampler *s, double *U, double *S);
void orc_set_factor(orc_sampler *s, const double *U, const double *S);
/* after every covariance update take the factor from this table instead of the
 * built-in eigensolver (used to remove LAPACK's sign ambiguity in golden tests) */
void orc_inject_factors(orc_sampler *s, int n, const double *Us, const double *Ss);
void orc_get_buffers(const orc_sampler *s, double *am /*[cU][W][d]*/, double *de /*[burn][W][d]*/);
/* prop/acc [T][W][njumps] ; swap_acc [T][W] */
void orc_get_counters(const orc_sampler *s, int64_t *prop, int64_t *acc, int64_t *swap_acc,
                      int64_t *swap_proposed);
int32_t orc_njumps(const orc_sampler *s);
/* trace: one byte per chain-step = jump id | accepted<<7, layout [iter][T][W];
 * swap maps int16 [event][W][T] */
void orc_set_trace(orc_sampler *s, uint8_t *trace, int64_t trace_iters, int16_t *swapmaps,
                   int64_t swap_events);

/* --- ladder sharding: the swap sweep (ref :631-697) cut at shard boundaries.  A message is
 *     (ndim+3)*nwalkers doubles: x[ndim][W], lnl[W], lnprior[W], origin rung[W].
 *     Per swap iteration (orc_run stops there): pack_top -> hotter neighbour; sweep(carry from the
 *     hotter neighbour or NULL, carry for the colder neighbour or NULL); finish(top rung of the colder
 *     neighbour or NULL) applies the permutation and the iteration's updateChains. --- */
int64_t orc_swap_msg_doubles(const orc_sampler *s);
int orc_swap_pending(const orc_sampler *s);
void orc_swap_pack_top(const orc_sampler *s, double *msg);
int orc_swap_sweep(orc_sampler *s, const double *carry_in, double *carry_out);
int orc_swap_finish(orc_sampler *s, const double *below_top);
/* AM ring [cov_update][W][d] in place (ladder sharding: broadcast from the cold shard) */
double *orc_am_ring(orc_sampler *s);
/* run the covariance / DE maintenance due at the start of the next iteration now (idempotent) */
int orc_maintain(orc_sampler *s);
/* walker sharding: covariance update split around an exchange of batch moments {n, mean[d], M2c[d*d]} */
int orc_adapt_begin(orc_sampler *s, double *batch);
int orc_adapt_finish(orc_sampler *s, const double *batch);

/* --- RNG primitives, exported so that the golden harness can feed the very
 *     same draws to the unmodified reference --- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t orc_draw_word(uint64_t seed, uint32_t purpose, uint64_t iter, uint32_t walker,
                       uint32_t temp, uint32_t j);
uint64_t orc_word_to_int(uint64_t word, uint64_t n);
double orc_word_to_unit(uint64_t word);
void orc_word_to_normals(uint64_t word, double *z0, double *z1);
void orc_word_to_normals_many(const uint64_t *words, int64_t n, double *z0, double *z1);

/* symmetric eigen-factorisation used for U,S (round-robin Jacobi, sorted, sign-fixed) */
void orc_sym_factor(int n, const double *a, double *U, double *S);
void orc_temperature_ladder(int ndim, int ntemps, double tmin, double tmax, double *ladder);

#ifdef __cplusplus
}
#endif
#endif
