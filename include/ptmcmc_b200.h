/*
 * ptmcmc_b200.h -- C ABI of the B200-native parallel-tempering MCMC engine.
 *
 * This is the drop-in boundary for the hot path of nanograv/PTMCMCSampler
 * (reference PTMCMCSampler/PTMCMCSampler.py, "ref" below): the per-chain
 * Metropolis-Hastings step, the adaptive covariance / eigen-factor update, the
 * DE history and the inter-temperature swap, for W walkers x T temperatures
 * resident on one GPU.  Plain C types only: no torch, no Python objects, no
 * exceptions cross this line.  The host-side mirror of the reference API
 * (ptmcmcsampler_b200.PTSampler) binds exactly these symbols with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every call returns 0 on success, <0 on error; ptmcmc_last_error() gives
 *     the message (engine-local; ptmcmc_create_error() for a failed create).
 *   - all host buffers are caller-owned, C-contiguous; the engine never keeps a
 *     host pointer after the call returns.
 *   - chain state arrays are [T][W][d] (temperature, walker, parameter).
 *   - one host thread per engine; calls are asynchronous on the engine's CUDA
 *     stream unless they copy to host (getters synchronise).
 */
#ifndef PTMCMC_B200_H
#define PTMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTMCMC_ABI_VERSION 4

/* jump ids: the reference's built-in proposals (ref :820-985); PRIOR = a draw from the uniform prior
 * box, the device-side form of the "UniformJump" plugin of ref tests/test_simple.py:44-62 (needs
 * PTMCMC_LOGP_UNIFORM); ids >= PTMCMC_JUMP_EXT0 are host-side (Python) proposals registered with
 * addProposalToCycle (ref :988-1014) */
enum { PTMCMC_JUMP_SCAM = 0, PTMCMC_JUMP_AM = 1, PTMCMC_JUMP_DE = 2, PTMCMC_JUMP_PRIOR = 3, PTMCMC_JUMP_EXT0 = 4 };
/* built-in log-likelihoods (ref examples/simple.py:34-36, examples/curved_likelihood.ipynb) */
enum { PTMCMC_LOGL_EXTERNAL = 0, PTMCMC_LOGL_GAUSSIAN = 1, PTMCMC_LOGL_CURVED = 2, PTMCMC_LOGL_ROSENBROCK = 3,
       PTMCMC_LOGL_USER = 4 /* CUDA source supplied by the caller, compiled with NVRTC (logl_source) */ };
/* built-in log-priors (ref examples/simple.py:38-44) */
enum { PTMCMC_LOGP_EXTERNAL = 0, PTMCMC_LOGP_UNIFORM = 1, PTMCMC_LOGP_FLAT = 2,
       PTMCMC_LOGP_USER = 3 /* CUDA source supplied by the caller (logp_source) */ };

enum {
    PTMCMC_OK = 0,
    PTMCMC_ERR_ARG = -1,       /* bad argument / configuration                       */
    PTMCMC_ERR_DE_SHAPE = -2,  /* covUpdate > burn at a DE update (ref :817 ValueError) */
    PTMCMC_ERR_CUDA = -3,
    PTMCMC_ERR_STATE = -4,     /* call out of order (e.g. run before set_state)       */
    PTMCMC_ERR_CAPACITY = -5   /* record window full: fetch rows and release them     */
};

#define PTMCMC_MAX_CYCLE 16

/* Replaces the keyword arguments of PTSampler.__init__ (ref :75-93) and
 * PTSampler.sample / initialize (ref :157-181, :374-399) that drive the hot path. */
typedef struct ptmcmc_config {
    int32_t abi_version;        /* PTMCMC_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    int32_t ndim;               /* ref ndim */
    int32_t nwalkers;           /* W: independent ladders resident on this device */
    int32_t ntemps;             /* T: ref nchain = comm.Get_size() (:97) */
    int32_t walker_offset;      /* global id of local walker 0 (keys the RNG; walker sharding) */
    int32_t temp_offset;        /* global index of local rung 0 (keys the RNG; ladder sharding) */
    int32_t ntemps_global;      /* rungs of the whole ladder when it is sharded over engines (0 = ntemps) */
    uint64_t seed;              /* ref seed (:92); Philox key */
    const double *ladder;       /* [T] swap temperatures, ref self.ladder (:274-275, :658) */
    const double *mh_temp;      /* [T] MH temperatures, ref self.temp (:278-282; 1e80 for hotChain); NULL = ladder */
    const double *cov;          /* [d*d] initial proposal covariance, ref cov (:134) */
    int32_t ngroups;            /* ref groups (:129-131); 0 = one group of all parameters */
    int32_t reserved1;
    const int32_t *group_offsets;  /* [ngroups+1] CSR offsets */
    const int32_t *group_indices;  /* [group_offsets[ngroups]] parameter indices */
    int32_t ncycle;             /* proposal-cycle segments in registration order (ref :1007-1008) */
    int32_t de_weight;          /* ref DEweight: DE segment appended at iteration burn+1 (:563-585) */
    int32_t cycle_jump[PTMCMC_MAX_CYCLE];
    int32_t cycle_weight[PTMCMC_MAX_CYCLE];
    int64_t cov_update;         /* ref covUpdate */
    int64_t burn;               /* ref burn (DE history length, per walker) */
    int64_t tskip;              /* ref Tskip */
    int64_t thin;               /* ref thin */
    int32_t logl_kind;          /* PTMCMC_LOGL_* */
    int32_t logp_kind;          /* PTMCMC_LOGP_* */
    const double *logl_params;  /* GAUSSIAN: mu[d], icov[d*d] row-major, offset ; others: NULL */
    const double *logp_params;  /* UNIFORM: lo[d], hi[d], value_inside, inclusive(0/1) */
    int32_t record_hot;         /* 0: record the T=1 rung only; 1: all rungs (ref writeHotChains) */
    int32_t trace;              /* 1: keep a per-iteration jump/accept byte and swap maps (tests) */
    int64_t record_rows;        /* capacity of the device-resident record window, in rows */
    int64_t trace_iters;        /* capacity of the trace, in iterations (0 if trace == 0) */
    int32_t timing;             /* 1: bracket every launch with CUDA events (ptmcmc_get_timing) */
    int32_t reserved2;
    double ladder_above;        /* ladder sharding: temperature of rung temp_offset+ntemps (hotter neighbour) */
    double ladder_below;        /* ladder sharding: temperature of rung temp_offset-1 (colder neighbour) */
    /* User targets on the device (the reference's arbitrary logl / logp callables, ref :108-109, :605-612,
     * :1072-1086, as CUDA source instead of Python).  logl_source defines
     *     __device__ double user_logl(const double *x, int ndim, const double *par);
     * logp_source defines  __device__ double user_logp(const double *x, int ndim, const double *par);
     * returning -inf outside the prior support (ref :607-608).  Either may be NULL when the matching kind is a
     * built-in.  The sources are compiled with NVRTC for the device's architecture together with the MH
     * kernel at ptmcmc_create (cached by content hash); a compile error fails the create with the NVRTC log.
     * par points at a device copy of user_params. */
    const char *logl_source;
    const char *logp_source;
    const double *user_params;  /* [n_logl_user_params + n_logp_user_params]: user_logl gets the first block, user_logp the second */
    int32_t n_logl_user_params;
    int32_t n_logp_user_params;
} ptmcmc_config;

typedef struct ptmcmc_engine ptmcmc_engine;

/* kernel classes for ptmcmc_get_timing */
enum {
    PTMCMC_K_MH = 0, PTMCMC_K_SWAP = 1, PTMCMC_K_ADAPT = 2, PTMCMC_K_DE = 3, PTMCMC_K_INIT = 4,
    PTMCMC_K_PROPOSE = 5, PTMCMC_K_ACCEPT = 6, PTMCMC_K_NCLASSES = 8
};
typedef struct ptmcmc_timing {
    int64_t launches[PTMCMC_K_NCLASSES]; /* kernels launched since create / last reset     */
    double ms[PTMCMC_K_NCLASSES];        /* device time per class (only when cfg.timing=1) */
    int64_t chain_steps;                 /* MH chain-steps executed                        */
} ptmcmc_timing;

int32_t ptmcmc_abi_version(void);
int32_t ptmcmc_device_count(void);
const char *ptmcmc_create_error(void);
const char *ptmcmc_last_error(const ptmcmc_engine *e);

/* PTSampler.__init__ + initialize (ref :75-155, :157-319): allocate device state, factor cov */
ptmcmc_engine *ptmcmc_create(const ptmcmc_config *cfg);
void ptmcmc_destroy(ptmcmc_engine *e);

/* Initial point (ref sample :471-493): evaluate logp/logl on device, record row 0, fill AM slot 0.
 * x0 is [T][W][d]. */
int32_t ptmcmc_set_state(ptmcmc_engine *e, const double *x0);
/* Same with host-evaluated values (Python logl/logp): lnl, lnprior are [T][W].  A target the device knows (a built-in or
 * user source) is evaluated on the device and the host value ignored, so device and host targets can be mixed. */
int32_t ptmcmc_set_state_external(ptmcmc_engine *e, const double *x0, const double *lnl,
                                  const double *lnprior);

/* The hot loop (ref :495-528 driver + :530-629 PTMCMCOneStep) for niter iterations, built-in
 * proposals and targets only, entirely on device.  The random streams are keyed by (seed, iteration, purpose, walker,
 * rung) with the iteration in one 32-bit counter word: a run ends at iteration 2^32 - 1 (PTMCMC_ERR_ARG beyond). */
int32_t ptmcmc_run(ptmcmc_engine *e, int64_t niter);

/* Slow path for Python callables, one iteration per propose/accept pair.
 * propose: start iteration iter+1: covariance / DE maintenance (ref :545-585), draw the jump and
 *   the proposal for every chain (ref :601, _jump :1048-1067).  q [T][W][d] receives the proposals
 *   (q = x for chains whose jump id >= PTMCMC_JUMP_EXT0: the host fills those), jump [T][W] the ids.
 * accept: finish the iteration with host-supplied q, qxy, log-likelihood and log-prior of q
 *   ([T][W] each; lnprior = -inf marks "logl not evaluated", ref :607-608): Hastings test
 *   (ref :614-622), swap (ref :624-625, :631-697), buffers and record (ref :627). */
int32_t ptmcmc_propose(ptmcmc_engine *e, double *q, int32_t *jump);
int32_t ptmcmc_accept(ptmcmc_engine *e, const double *q, const double *qxy, const double *lnl,
                      const double *lnprior);

/* The same round trip through engine-owned page-locked buffers, with ONE host synchronisation per iteration:
 * ptmcmc_callback_buffers hands out q [T][W][d], jump [T][W], qxy / lnl / lnprior [T][W] and x [T][W][d] (valid until
 * ptmcmc_destroy).  ptmcmc_propose_pinned fills q and jump (and x, the current points, when want_x != 0) and returns
 * once they have landed; the host writes its proposals / qxy / target values in place; ptmcmc_accept_pinned uploads
 * what the host owns (q only if q_modified; lnl / lnprior only for PTMCMC_LOG*_EXTERNAL targets) and enqueues the rest
 * of the iteration without waiting for it. */
int32_t ptmcmc_callback_buffers(ptmcmc_engine *e, double **q, int32_t **jump, double **qxy, double **lnl,
                                double **lnprior, double **x);
int32_t ptmcmc_propose_pinned(ptmcmc_engine *e, int32_t want_x);
int32_t ptmcmc_accept_pinned(ptmcmc_engine *e, int32_t q_modified);

int64_t ptmcmc_iteration(const ptmcmc_engine *e);
int32_t ptmcmc_sync(ptmcmc_engine *e);

/* ref p0 / lnlike0 / lnprob0 of every chain; any pointer may be NULL */
int32_t ptmcmc_get_state(ptmcmc_engine *e, double *x, double *lnl, double *lnprior, double *lnprob);

/* ref _chain/_lnlike/_lnprob (:208-212, :331-335).  Rows are numbered iter/thin; the device keeps
 * a window [row_base, row_base+record_rows).  chain is [nrows][ntr][W][d], lnl/lnprob [nrows][ntr][W],
 * ntr = record_hot ? T : 1. */
int64_t ptmcmc_rows(const ptmcmc_engine *e);
int64_t ptmcmc_row_base(const ptmcmc_engine *e);
int32_t ptmcmc_get_chain(ptmcmc_engine *e, int64_t row0, int64_t nrows, double *chain, double *lnl,
                         double *lnprob);
int32_t ptmcmc_release_rows(ptmcmc_engine *e, int64_t upto_row);

/* ref cov / mu / M2 (:147-148, :769-794) and the per-group factor U, S (:797-803) */
int32_t ptmcmc_get_adapt(ptmcmc_engine *e, double *cov, double *mu, double *m2, int64_t *nsamp);
int32_t ptmcmc_get_factor(ptmcmc_engine *e, double *U, double *S);
int32_t ptmcmc_set_factor(ptmcmc_engine *e, const double *U, const double *S);
/* ref _AMbuffer [covUpdate][W][d] and _DEbuffer [burn][W][d] (:219-221) */
int32_t ptmcmc_get_buffers(ptmcmc_engine *e, double *am, double *de);

/* Pooled adaptation across devices (walker sharding).  begin: if a covariance update is due at
 * the current iteration, compute this device's batch moments into batch_out = {n, mean[d],
 * M2c[d*d]} and return 1 (0 if none is due).  finish: apply the (all-reduced) batch. */
int32_t ptmcmc_adapt_begin(ptmcmc_engine *e, double *batch_out);
int32_t ptmcmc_adapt_finish(ptmcmc_engine *e, const double *batch_in);

/* The same without a host round trip (the collective runs on device memory, e.g. NCCL all-gather on the engine's
 * stream).  begin_dev: dev_batch receives the device address of this engine's batch buffer (ndoubles doubles), filled
 * in stream order when an update is due (returns 1, else 0).  finish_dev: dev_parts holds nparts such batches back to
 * back in DEVICE memory; they are merged in part order by a kernel (Chan's formula, as the host path) and applied;
 * nsamples = the samples all parts hold together (known to the host: covUpdate x walkers per part). */
int32_t ptmcmc_adapt_begin_dev(ptmcmc_engine *e, void **dev_batch, int64_t *ndoubles);
int32_t ptmcmc_adapt_finish_dev(ptmcmc_engine *e, const double *dev_parts, int32_t nparts, int64_t nsamples);
/* device addresses of the eigen-factor (U concatenated per group, S) so that the shard holding T=1 can broadcast it in
 * place (the multi-device form of the reference's send(cov) + per-rank SVD, ref :545-560); ptmcmc_factor_refresh
 * re-derives sqrt(S) and the tensor-core operand images after the buffers were written from outside */
int32_t ptmcmc_factor_dev(ptmcmc_engine *e, void **dev_U, int64_t *usize, void **dev_S, int64_t *ssize);
int32_t ptmcmc_factor_refresh(ptmcmc_engine *e);

/* Ladder sharding (ntemps_global > ntemps): the swap sweep (ref PTswap :631-697) cut at the shard
 * boundaries; replaces the reference's gather / scatter to rank 0 (ref :660-661, :689-691) by a
 * nearest-neighbour exchange of one rung.  ptmcmc_run stops at every swap iteration (it refuses to
 * cross one) and leaves the swap pending; the caller then moves three messages between neighbours:
 *   1. ptmcmc_swap_pack_top(msg_up)            -> send to the hotter neighbour (no dependency)
 *   2. ptmcmc_swap_sweep(carry_in, carry_out)  carry_in: received from the hotter neighbour (NULL on
 *      the hottest shard); carry_out: to send to the colder neighbour (NULL on the coldest shard)
 *   3. ptmcmc_swap_finish(below_top)           below_top: the colder neighbour's msg_up (NULL on the
 *      coldest shard).  Applies the permutation and the iteration's buffer/record writes (ref :627).
 * Both sides of a boundary evaluate the same acceptance from the counter-based stream.  A message is
 * ptmcmc_swap_msg_doubles() doubles in DEVICE memory: x[ndim][W], lnl[W], lnprior[W], origin rung[W].
 * carry_in must stay valid until ptmcmc_swap_finish returns. */
int64_t ptmcmc_swap_msg_doubles(const ptmcmc_engine *e);
int32_t ptmcmc_swap_pending(const ptmcmc_engine *e);
int32_t ptmcmc_swap_pack_top(ptmcmc_engine *e, double *dev_msg);
int32_t ptmcmc_swap_sweep(ptmcmc_engine *e, const double *dev_carry_in, double *dev_carry_out);
int32_t ptmcmc_swap_finish(ptmcmc_engine *e, const double *dev_below_top);

/* The same three steps with the two messages moved by the kernels themselves through peer memory (NVLink stores into
 * the neighbour's mailbox, a sequence number published by the last block, the consumer's kernel spinning on the flag in
 * its own memory) instead of a send / receive pair per hop: replaces the same reference lines (:660-661, :689-691).
 *   ptmcmc_p2p_open(e, handle, &ptr)   allocates this shard's mailbox; handle receives its 64-byte CUDA IPC handle
 *                                      (NULL to skip), ptr its device address (for shards living in one process)
 *   ptmcmc_p2p_connect(e, above, below, ipc, seq0)  the hotter / colder neighbour's mailbox: IPC handles when ipc != 0,
 *                                      else device addresses; NULL exactly where there is no neighbour.  Collective: every
 *                                      shard connects between the same two swaps and passes the same seq0 >= the largest
 *                                      ptmcmc_p2p_seq() among them (mailboxes are reused by later engines of the process,
 *                                      their flags only ever grow).
 *   ptmcmc_swap_p2p(e, phase)          phase 0: pack, 1: sweep, 2: finish -- only enqueues work; every shard issues the
 *                                      three in this order.  Shards sharing ONE device must issue phase by phase
 *                                      (all packs, all sweeps hottest first, all finishes).
 *   ptmcmc_p2p_error(e)                synchronises; an error if a message did not arrive within about two seconds */
int32_t ptmcmc_p2p_open(ptmcmc_engine *e, void *ipc_handle_out, void **mailbox_out);
int64_t ptmcmc_p2p_seq(const ptmcmc_engine *e);
int32_t ptmcmc_p2p_connect(ptmcmc_engine *e, const void *above, const void *below, int32_t ipc, int64_t seq0);
int32_t ptmcmc_swap_p2p(ptmcmc_engine *e, int32_t phase);
int32_t ptmcmc_p2p_error(ptmcmc_engine *e);
/* device address and length of the AM ring: the cold shard broadcasts it before a DE update, the
 * multi-device form of rank 0's send(_DEbuffer) (ref :563-571) */
int32_t ptmcmc_am_ring(ptmcmc_engine *e, void **dev_ptr, int64_t *ndoubles);
/* run now the covariance / DE maintenance due at the start of the next iteration (ref :545-585);
 * ptmcmc_run will not repeat it.  Lets the caller place collectives around it. */
int32_t ptmcmc_maintain(ptmcmc_engine *e);

/* Checkpoint of the complete sampling state (chain state, adaptive state, AM ring, DE history, counters,
 * iteration number).  The reference resumes by replaying its chain file (ref :290-319, :591-599) and
 * cannot continue bit-exactly because generator state is not saved; here the streams are counter-based,
 * so state + iteration number IS the generator state and a loaded engine continues exactly as the
 * original would have.  The engine that loads must have been created with the same configuration.  The
 * record window is not part of the checkpoint: fetch rows before saving. */
int64_t ptmcmc_state_bytes(const ptmcmc_engine *e);
int32_t ptmcmc_save_state(ptmcmc_engine *e, void *buf, int64_t nbytes);
int32_t ptmcmc_load_state(ptmcmc_engine *e, const void *buf, int64_t nbytes);
/* the seed a checkpoint was written with (the loading engine must be created with it: ptmcmc_load_state refuses a
 * checkpoint whose seed, thin, Tskip, walker offset, ladder or proposal cycle differ from the engine's) */
int32_t ptmcmc_state_seed(const void *buf, int64_t nbytes, uint64_t *seed);
/* Reference-style resume (ref :591-599): advance the engine through `nrows * repeat` iterations whose
 * states are given instead of proposed -- row r is used for `repeat` (= thin) consecutive iterations, the
 * first call starting at iteration 1 with row index (iteration / repeat).  Buffers, covariance / DE
 * maintenance and the record run as in ptmcmc_run; no proposal is drawn and no swap is made.
 * x is [nrows][T][W][d], lnl / lnprior [nrows][T][W]. */
int32_t ptmcmc_replay(ptmcmc_engine *e, int64_t niter, int64_t repeat, int64_t nrows, const double *x,
                      const double *lnl, const double *lnprior);

/* ref jumpDict[name] = [proposed, accepted] (:602, :622) per chain: prop/acc are [njumps][T][W];
 * nswap_accepted per chain [T][W] (:691) and swapProposed (:692) */
int32_t ptmcmc_njumps(const ptmcmc_engine *e);
int32_t ptmcmc_get_counters(ptmcmc_engine *e, int64_t *prop, int64_t *acc, int64_t *swap_acc,
                            int64_t *swap_proposed);
/* test hook: trace bytes [iters][T][W] (jump | accepted<<7) and swap maps [events][W][T] */
int32_t ptmcmc_get_trace(ptmcmc_engine *e, uint8_t *trace, int64_t iters, int16_t *swapmaps,
                         int64_t events);

/* Record sink: overlaps the device->host copy of the thinned record with sampling.  The reference stores every
 * thinned sample into _chain / _lnlike / _lnprob as it goes (ref :331-335); with a sink registered the engine does
 * the same: after every enqueued segment the rows it completed are copied, on a second stream ordered by events,
 * to the caller's arrays chain [row_capacity][ntr][W][d], lnl / lnprob [row_capacity][ntr][W] at their absolute
 * row positions (rows >= row_capacity are dropped).  EXCEPTION to the "no host pointer is kept" rule: the three
 * arrays (page-locked memory, e.g. ptmcmc_host_alloc) must stay valid until ptmcmc_set_sink(e, NULL, ...) or
 * ptmcmc_destroy.  With a sink the device window needs no ptmcmc_release_rows: it is reused as soon as its rows
 * have landed, and ptmcmc_run never fails with PTMCMC_ERR_CAPACITY. */
int32_t ptmcmc_set_sink(ptmcmc_engine *e, double *chain, double *lnl, double *lnprob, int64_t row_capacity);
/* block until every row recorded by the calls made so far has landed in the sink */
int32_t ptmcmc_sink_wait(ptmcmc_engine *e);
/* Asynchronous snapshot of what the reference's writeOutput needs (ref :341-372, :722-766), ordered after the
 * work enqueued so far and before anything enqueued later: host (page-locked, ptmcmc_snapshot_bytes() bytes)
 * receives  int64 {iteration, swapProposed, nsamp, njumps}, int64 prop_sum[nj][T], acc_sum[nj][T] (summed over
 * walkers), prop_w0[nj][T], acc_w0[nj][T] (walker 0), swap_sum[T], swap_w0[T], then doubles cov[d*d], mu[d],
 * M2[d*d], U, S.  Two slots so that the host can write files for one boundary while the next segment runs;
 * ptmcmc_snapshot_wait(slot) blocks until that snapshot (and every sink row before it) has landed. */
int64_t ptmcmc_snapshot_bytes(const ptmcmc_engine *e);
int32_t ptmcmc_snapshot(ptmcmc_engine *e, void *host, int64_t nbytes, int32_t slot);
int32_t ptmcmc_snapshot_wait(ptmcmc_engine *e, int32_t slot);

int32_t ptmcmc_get_timing(ptmcmc_engine *e, ptmcmc_timing *out);
int32_t ptmcmc_reset_timing(ptmcmc_engine *e);
/* switch the per-launch CUDA-event bracketing on or off (same as cfg.timing) */
int32_t ptmcmc_set_timing(ptmcmc_engine *e, int32_t on);
/* name of the kernel ptmcmc_run launches for the MH segments of this engine (for profiles and bench records) */
const char *ptmcmc_mh_kernel_name(ptmcmc_engine *e);
/* compile-only check of user target sources for compute capability cc_major.cc_minor (no device needed): returns the
 * cubin size, or < 0 with the NVRTC log in `log` */
int32_t ptmcmc_user_compile_check(const char *logl_source, const char *logp_source, int32_t cc_major, int32_t cc_minor,
                                  char *log, int64_t log_capacity);
/* measured fp64 FMA throughput of `device` in TFLOP/s (a short DFMA kernel; the compute-side roofline denominator) */
int32_t ptmcmc_measure_fp64_peak(int32_t device, double *tflops);
/* test hook: the device's Box-Muller pairs of n 64-bit words (word_to_normals) on `device`; host buffers */
int32_t ptmcmc_test_normals(int32_t device, const uint64_t *words, int64_t n, double *z0, double *z1);
/* the engine's CUDA stream (cudaStream_t) for callers that time with their own events */
void *ptmcmc_stream(ptmcmc_engine *e);
/* page-locked host memory for the caller-owned buffers (faster DMA); plain malloc memory works too */
void *ptmcmc_host_alloc(int64_t bytes);
void ptmcmc_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* PTMCMC_B200_H */
