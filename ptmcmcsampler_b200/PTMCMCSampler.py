"""Host-side mirror of ``PTMCMCSampler.PTMCMCSampler.PTSampler`` (nanograv/PTMCMCSampler).

Same constructor, ``sample()`` signature, plugin surface (``addProposalToCycle``,
``addAuxilaryJump``) and public attributes as the reference (ref PTMCMCSampler.py:40-1069); the
hot path runs in the CUDA engine behind ``include/ptmcmc_b200.h``.  What stays in Python is what
the reference also does on the host around the step: configuration, the driver loop's chunking,
chain / jump files and the progress line.

Differences a reference user should know:

* one process drives all temperatures (``ntemps=``) and, new, ``nwalkers`` independent ladders;
  ``comm`` is accepted for signature parity only.  The reference's one-MPI-rank-per-temperature layout
  is covered by ``dist_group=..., shard="ladder"`` (one process per GPU, several rungs each);
* ``logl`` / ``logp`` may be objects from :mod:`ptmcmcsampler_b200.likelihoods` (fully on device)
  or plain callables (evaluated on the host once per iteration, as are custom Python jumps);
* random numbers come from a counter-based Philox stream keyed by ``seed`` instead of PCG64, so
  runs agree with the reference in distribution, not draw by draw;
* the gradient proposals (NUTS / HMC / MALA, ref nutsjump.py) are host-side plugins built from the user's Python
  gradients (``ptmcmcsampler_b200.nutsjump``): registered when both ``logl_grad`` and ``logp_grad`` are given, as in the
  reference, and executed between the engine's propose / accept calls.
"""
import os
import sys
import time

import numpy as np

from . import _cabi
from . import nompi4py as MPI
from .likelihoods import DeviceLogLikelihood, DeviceLogPrior

__all__ = ["PTSampler", "shift_array"]


def shift_array(arr, num, fill_value=0.0):
    """Shift rows of ``arr`` by ``num`` places, filling the vacated rows (ref :27-37)."""
    arr = np.asarray(arr)
    out = np.full_like(arr, fill_value)
    if num > 0:
        out[num:] = arr[:-num]
    elif num < 0:
        out[:num] = arr[-num:]
    else:
        out[...] = arr
    return out


def integrated_time(x, c=5.0):
    """Integrated autocorrelation time of a 1-D series: tau = 1 + 2 sum_{t <= M} rho(t) with the smallest window
    M >= c tau (Sokal's self-consistent rule, the estimate the ``acor`` package returns)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    x = x - x.mean()
    if n < 4 or not np.any(x):
        return 1.0
    f = np.fft.rfft(x, 2 * int(2 ** np.ceil(np.log2(n))))
    acf = np.fft.irfft(f * np.conjugate(f))[:n]
    acf /= acf[0]
    taus = 2.0 * np.cumsum(acf) - 1.0
    window = np.arange(n) >= c * taus
    m = int(np.argmax(window)) if window.any() else n - 1
    return float(max(1.0, taus[m]))


class _function_wrapper(object):
    """Binds ``args`` / ``kwargs`` to a user callable (ref :1072-1086)."""

    def __init__(self, f, args, kwargs):
        self.f, self.args, self.kwargs = f, args, kwargs

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


class _BuiltinJump(object):
    """Names one of the device-side proposals in ``propCycle`` / ``jumpDict`` (ref :820-985).  A plain
    object rather than a bound method, so the sampler is not part of a reference cycle and its
    page-locked result arrays and device memory are released as soon as it goes out of scope."""

    def __init__(self, name, jid, doc):
        self.__name__, self.jid, self.__doc__ = name, jid, doc

    def __call__(self, x, iter, beta):
        raise NotImplementedError("the %s proposal is drawn on the device; this handle only names it" % self.__name__)

    def __repr__(self):
        return "<device proposal %s>" % self.__name__


class PTSampler(object):
    """Parallel-tempering MCMC sampler with adaptive (AM / SCAM) and differential-evolution
    proposals; API of the reference's ``PTSampler`` (ref :40-155).

    Extra keyword arguments (all optional, defaults reproduce a single reference chain):

    :param ntemps: number of temperatures (the reference takes it from ``comm.Get_size()``)
    :param nwalkers: independent ladders run side by side; the adaptive covariance and the DE
        history are pooled over the walkers' T=1 chains
    :param device: CUDA device ordinal
    :param record_rows: rows of the thinned record kept on the device between flushes
    :param walker_offset: global id of this process's walker 0 (walker sharding over several GPUs)
    :param dist_group: ``torch.distributed`` process group this sampler is sharded over (``True`` =
        the default group); ``None`` keeps this sampler independent
    :param vectorized: Python ``logl`` / ``logp`` (and custom jumps flagged ``func.vectorized = True``)
        take all chains at once: ``logl(X[n, ndim]) -> [n]``, ``logp(X) -> [n]`` (``-inf`` where the prior
        rejects), ``jump(X[n, ndim], iter, beta[n]) -> (Q[n, ndim], qxy[n])``, auxiliary jumps
        ``aux(X, Q, iter, beta) -> (Q, qxy)``.  One host call per iteration instead of one per chain.
    :param checkpoint: write ``<outDir>/engine_state.npy`` (the complete device state) at every
        ``isave``; with ``resume=True`` such a file is preferred over replaying the chain file and the run
        continues exactly where it stopped (the draws are counter-based, so state + iteration is all
        there is to restore)
    :param shard: with ``dist_group``: ``"walkers"`` (default) -- every rank runs its own ``nwalkers``
        complete ladders and only the proposal covariance is pooled; ``"ladder"`` -- the ``ntemps`` rungs
        are split contiguously over the ranks (rank 0 holds T=1, like the reference's MPI layout with
        several rungs per rank) and the swap exchanges the boundary rung between neighbours
    """

    def __init__(self, ndim, logl, logp, cov, groups=None, loglargs=[], loglkwargs={}, logpargs=[],
                 logpkwargs={}, logl_grad=None, logp_grad=None, comm=MPI.COMM_WORLD, outDir="./chains",
                 verbose=True, resume=False, seed=None, ntemps=None, nwalkers=1, device=0, record_rows=None,
                 walker_offset=0, dist_group=None, shard="walkers", checkpoint=False, vectorized=False):
        self.comm = comm
        self.MPIrank = 0
        if comm is not None and hasattr(comm, "Get_size") and comm.Get_size() > 1:
            raise NotImplementedError(
                "one process drives every temperature on the GPU: pass ntemps=%d instead of launching "
                "one MPI rank per temperature" % comm.Get_size())
        self.nchain = int(ntemps) if ntemps else 1
        self.nwalkers = int(nwalkers)
        self.device = int(device)
        self.walker_offset = int(walker_offset)
        self._record_rows = record_rows
        self.dist_group = dist_group  # torch.distributed group (walker or ladder sharding)
        if shard not in ("walkers", "ladder"):
            raise ValueError("shard must be 'walkers' or 'ladder'")
        self.shard = shard
        self._group = None if dist_group is True else dist_group
        self._shard_rank, self._shard_world = 0, 1
        if dist_group is not None and shard == "ladder":
            import torch.distributed as dist

            self._shard_rank, self._shard_world = dist.get_rank(self._group), dist.get_world_size(self._group)
            self.MPIrank = self._shard_rank  # rank g holds rungs [g*T/G, (g+1)*T/G), ref :94-97, :278
        if dist_group is not None and shard == "walkers":
            import torch.distributed as dist

            r = dist.get_rank(self._group)
            if r > 0:  # every rank holds its own walkers: its files must not collide with another rank's
                outDir = os.path.join(outDir, "rank%d" % r)
        self._lo, self._Tloc = 0, self.nchain
        self._comm = None
        if seed is None and resume:
            # an engine checkpoint continues exactly only under the seed it was written with (the draws are keyed by it)
            for cand in ("engine_state.npy", "engine_state_%d.npy" % self._shard_rank):
                path = os.path.join(outDir, cand)
                if os.path.isfile(path):
                    seed = _cabi.checkpoint_seed(np.load(path, mmap_mode="r")[:4096])
                    if seed is not None:
                        break
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed)
        self.stream = np.random.default_rng(self.seed)  # host-side draws for user plugins only

        self.ndim = int(ndim)
        self._dev_logl = logl if isinstance(logl, DeviceLogLikelihood) else None
        self._dev_logp = logp if isinstance(logp, DeviceLogPrior) else None
        self.logl = logl if self._dev_logl is not None else _function_wrapper(logl, loglargs, loglkwargs)
        self.logp = logp if self._dev_logp is not None else _function_wrapper(logp, logpargs, logpkwargs)
        # gradient proposals need both gradients, as in the reference (ref :110-115)
        self.logl_grad = self.logp_grad = None
        if logl_grad is not None and logp_grad is not None:
            self.logl_grad = _function_wrapper(logl_grad, loglargs, loglkwargs)
            self.logp_grad = _function_wrapper(logp_grad, logpargs, logpkwargs)

        self.outDir = outDir
        self.verbose = verbose
        self.resume = resume
        self.checkpoint = bool(checkpoint)
        self.vectorized = bool(vectorized)
        if not os.path.exists(self.outDir):
            try:
                os.makedirs(self.outDir)
            except OSError:
                pass

        # parameter groups (ref :129-131) and the initial eigen-factor of each block (ref :133-145)
        self.groups = groups if groups is not None else [np.arange(0, self.ndim)]
        self.cov = cov
        cov_arr = np.asarray(cov, dtype=np.float64)
        if cov_arr.shape != (self.ndim, self.ndim):
            raise ValueError("cov must be (ndim, ndim)")
        self.U = [[]] * len(self.groups)
        self.S = [[]] * len(self.groups)
        self.M2 = np.zeros((self.ndim, self.ndim))
        self.mu = np.zeros(self.ndim)

        self.propCycle = []
        self.jumpDict = {}
        self.aux = []
        self._ext_jumps = []  # user callables in registration order -> engine jump ids 3, 4, ...
        self._engine = None
        # handles of the built-in proposals, same names as the reference's methods (ref :820, :879, :936)
        self.covarianceJumpProposalSCAM = _BuiltinJump(
            "covarianceJumpProposalSCAM", _cabi.JUMP_SCAM, "single-component adaptive jump (ref :820-876)")
        self.covarianceJumpProposalAM = _BuiltinJump(
            "covarianceJumpProposalAM", _cabi.JUMP_AM, "adaptive-Metropolis jump (ref :879-933)")
        self.DEJump = _BuiltinJump("DEJump", _cabi.JUMP_DE, "differential-evolution jump (ref :936-985)")
        # beyond the reference: a draw from the uniform prior box on the device (the UniformJump plugin of the
        # reference's tests/test_simple.py:44-62); add with addProposalToCycle(sampler.priorDrawJump, weight)
        self.priorDrawJump = _BuiltinJump("priorDrawJump", _cabi.JUMP_PRIOR, "draw from the UniformPrior box")

    # ------------------------------------------------------------------ plugin surface --------
    def addProposalToCycle(self, func, weight):
        """Add ``func(x, iter, beta) -> (q, qxy)`` to the cycle with integer ``weight`` (ref :988-1014)."""
        if weight == 0:  # silently ignored, as in the reference (:1001-1004)
            return
        self.propCycle.extend([func] * int(weight))
        if func.__name__ not in self.jumpDict:
            self.jumpDict[func.__name__] = [0, 0]
            open(os.path.join(self.outDir, func.__name__ + "_jump.txt"), "w").close()

    def addAuxilaryJump(self, func):
        """Add ``func(x, q, iter, beta) -> (q, qxy_aux)`` applied after every proposal (ref :1017-1028)."""
        self.aux.append(func)

    def randomizeProposalCycle(self):
        """Kept for API parity; the reference's shuffled copy is never read (ref :1031-1045)."""
        index = np.arange(len(self.propCycle))
        self.stream.shuffle(index)
        self.randomizedPropCycle = [self.propCycle[i] for i in index]

    def temperatureLadder(self, Tmin, Tmax=None, tstep=None):
        """Geometric ladder ``Tmin * tstep**i`` (ref :699-720)."""
        if self.nchain > 1:
            if tstep is None and Tmax is None:
                tstep = 1 + np.sqrt(2 / self.ndim)
            elif tstep is None:
                tstep = np.exp(np.log(Tmax / Tmin) / (self.nchain - 1))
            return Tmin * tstep ** np.arange(self.nchain, dtype=np.float64)
        return np.array([1])

    # ------------------------------------------------------------------ engine plumbing -------
    def _cycle_segments(self):
        """propCycle (weight-replicated list) -> [(jump id, weight)] in order."""
        segs = []
        for f in self.propCycle:
            if isinstance(f, _BuiltinJump):
                jid = f.jid
            else:
                if f not in self._ext_jumps:
                    self._ext_jumps.append(f)
                jid = _cabi.JUMP_EXT0 + self._ext_jumps.index(f)
            if segs and segs[-1][0] == jid:
                segs[-1][1] += 1
            else:
                segs.append([jid, 1])
        return [(j, w) for j, w in segs]

    @property
    def _external(self):
        return self._dev_logl is None or self._dev_logp is None or bool(self._ext_jumps) or bool(self.aux)

    def _make_engine(self, maxIter):
        d = self.ndim
        rows_total = int(maxIter / self.thin) + 1
        rr = self._record_rows or min(rows_total, max(int(self.isave // self.thin) + 2, 64))
        rr = max(rr, int(self.isave // self.thin) + 2)
        mh_temp = np.array(self.ladder, dtype=np.float64)
        if self._hotChain:
            mh_temp[-1] = 1e80  # ref :281-282
        shard_kw = {}
        ladder = np.asarray(self.ladder, np.float64)
        if self._shard_world > 1:
            from . import distributed

            shard_kw = distributed.ladder_shard_kwargs(ladder, self._shard_world, self._shard_rank)
            self._Tloc, ladder = shard_kw.pop("ntemps"), shard_kw.pop("ladder")
            self._lo = shard_kw["temp_offset"]
            mh_temp = mh_temp[self._lo:self._lo + self._Tloc].copy()
        self._mh_temp = mh_temp
        self._engine = _cabi.Engine(
            d, self.nwalkers, self._Tloc, np.asarray(self.cov, dtype=np.float64), ladder,
            mh_temp=mh_temp, seed=self.seed, **shard_kw, groups=None if self._default_groups() else self.groups,
            cycle=self._cycle_segments(), de_weight=self.DEweight, cov_update=self.covUpdate, burn=self.burn,
            tskip=self.Tskip, thin=self.thin,
            logl_kind=self._dev_logl.kind if self._dev_logl is not None else _cabi.LOGL_EXTERNAL,
            logl_params=self._dev_logl.params(d) if self._dev_logl is not None else None,
            logp_kind=self._dev_logp.kind if self._dev_logp is not None else _cabi.LOGP_EXTERNAL,
            logp_params=self._dev_logp.params(d) if self._dev_logp is not None else None,
            logl_source=getattr(self._dev_logl, "source", None), logp_source=getattr(self._dev_logp, "source", None),
            logl_user_params=getattr(self._dev_logl, "user_params", None),
            logp_user_params=getattr(self._dev_logp, "user_params", None),
            record_hot=self.writeHotChains, record_rows=rr, device=self.device, walker_offset=self.walker_offset)
        if self._shard_world > 1:
            from . import distributed

            if self._external:
                raise NotImplementedError("ladder sharding needs device targets and the built-in proposals")
            self._comm = distributed.LadderComm(self._engine, self._group)
        # the thinned T=1 record streams into the result arrays while the engine runs (hot-rung records and
        # engine checkpoints are taken synchronously at every write instead)
        self._async = self._engine.ntr == 1 and not self._external and not self.checkpoint
        if self._engine.ntr == 1:
            self._engine.set_sink(self._chain_all, self._lnlike_all, self._lnprob_all)
        self._pull_factor()

    def _default_groups(self):
        return len(self.groups) == 1 and np.array_equal(np.asarray(self.groups[0]), np.arange(self.ndim))

    def _pull_factor(self):
        U, S = self._engine.factor()
        uo = so = 0
        for g, grp in enumerate(self.groups):
            n = len(grp)
            self.U[g] = U[uo:uo + n * n].reshape(n, n).copy()
            self.S[g] = S[so:so + n].copy()
            uo += n * n
            so += n

    def _pull_adapt(self):
        cov, mu, m2, _ = self._engine.adapt()
        np.asarray(self.cov)[:, :] = cov  # in place, like the reference (:794)
        self.mu, self.M2 = mu, m2
        self._pull_factor()

    # ------------------------------------------------------------------ initialize ------------
    def initialize(self, Niter, ladder=None, Tmin=1, Tmax=None, Tskip=100, isave=1000, covUpdate=1000,
                   SCAMweight=30, AMweight=20, DEweight=50, NUTSweight=20, HMCweight=20, MALAweight=0,
                   burn=50000, HMCstepsize=0.1, HMCsteps=300, maxIter=None, thin=10, i0=0, neff=None,
                   writeHotChains=False, hotChain=False):
        """Set the run's knobs, register the default proposals, build the device engine (ref :157-319)."""
        if maxIter is None:
            maxIter = Niter
        self.ladder = ladder
        self.covUpdate, self.burn, self.Tskip, self.thin, self.isave = int(covUpdate), int(burn), int(Tskip), int(thin), int(isave)
        self.SCAMweight, self.AMweight, self.DEweight = int(SCAMweight), int(AMweight), int(DEweight)
        self.Niter, self.neff, self.tstart = Niter, neff, 0

        N = int(maxIter / thin) + 1
        W = self.nwalkers
        # page-locked so that the device record is DMA'd straight into these arrays
        self._chain_all = _cabi.pinned_empty((N, W, self.ndim))
        self._lnlike_all = _cabi.pinned_empty((N, W))
        self._lnprob_all = _cabi.pinned_empty((N, W))
        # rows are filled as the device record is pulled; the unwritten tail is zeroed in _finish()
        self._chain, self._lnlike, self._lnprob = self._chain_all[:, 0], self._lnlike_all[:, 0], self._lnprob_all[:, 0]
        self._hot_rows = None
        self.ind_next_write = 0
        self._rows_pulled = 0
        self._counters_at = None
        self._acc_offset = 0.0
        self._full_counters = None
        self._acc_w0 = self._swap_w0 = None
        self.naccepted = 0
        self.swapProposed = 0
        self.nswap_accepted = 0

        # gradient-based jumps first, as the reference registers them (ref :226-258); host-side plugins built from the
        # user's gradient callables (nutsjump.py)
        if self.logl_grad is not None and self.logp_grad is not None:
            from . import nutsjump

            gkw = dict(rng=self.stream, batched_gradients=self.vectorized)
            cov0 = np.array(self.cov, dtype=np.float64)
            self.addProposalToCycle(nutsjump.MALAJump(self.logl_grad, self.logp_grad, cov0, self.burn, **gkw), int(MALAweight))
            self.addProposalToCycle(nutsjump.HMCJump(self.logl_grad, self.logp_grad, cov0, self.burn, stepsize=HMCstepsize,
                                                     nminsteps=2, nmaxsteps=HMCsteps, **gkw), int(HMCweight))
            self.addProposalToCycle(nutsjump.NUTSJump(self.logl_grad, self.logp_grad, cov0, self.burn, delta=0.6, **gkw),
                                    int(NUTSweight))
        self.addProposalToCycle(self.covarianceJumpProposalSCAM, self.SCAMweight)
        self.addProposalToCycle(self.covarianceJumpProposalAM, self.AMweight)
        if len(self.propCycle) == 0:
            raise ValueError("No jump proposals specified!")
        self.randomizeProposalCycle()

        if self.ladder is None:
            self.ladder = self.temperatureLadder(Tmin, Tmax=Tmax)
        self.ladder = np.asarray(self.ladder)
        if len(self.ladder) != self.nchain:
            raise ValueError("ladder has %d entries for %d temperatures" % (len(self.ladder), self.nchain))
        self.temp = self.ladder[self.MPIrank]
        self._hotChain = bool(hotChain) and self.nchain > 1
        self.writeHotChains = bool(writeHotChains)
        self._hot_fnames = [self.outDir + "/chain_{0}.txt".format(t) for t in self.ladder]
        if self._hotChain:
            self._hot_fnames[-1] = self.outDir + "/chain_hot.txt"
        if self._shard_world > 1:  # this rank's rungs; its first rung plays the reference's self.temp
            from . import distributed

            lo, hi = distributed.ladder_slice(self.nchain, self._shard_world, self._shard_rank)
            self.temp = self.ladder[lo]
            self._hot_fnames = self._hot_fnames[lo:hi]
            self.fname = self._hot_fnames[0]
        else:
            self.fname = self.outDir + "/chain_{0}.txt".format(self.temp)
        self._writes_primary = self.MPIrank == 0 or self.writeHotChains  # ref :346

        # resume (ref :289-319): an engine checkpoint if there is one, else the reference's replay of the chain file
        self.resumeLength = 0
        self.resumechain = None
        self._resumed_at = 0
        self._last_neff = 0.0
        self._state_file = os.path.join(self.outDir, "engine_state.npy" if self._shard_world == 1
                                        else "engine_state_%d.npy" % self._shard_rank)
        self._resume_state = self.resume and os.path.isfile(self._state_file)
        if self.resume and not self._resume_state and os.path.isfile(self.fname):
            if self.verbose:
                print("Resuming run from chain file {0}".format(self.fname))
            if self.nwalkers != 1 or self._shard_world != 1:
                raise NotImplementedError("a chain file holds one walker: resuming several walkers or a sharded "
                                          "run needs the engine checkpoint (checkpoint=True)")
            try:
                self.resumechain = np.loadtxt(self.fname, ndmin=2)
                self.resumeLength = self.resumechain.shape[0]
            except ValueError as error:
                print("Reading old chain files failed with error", error)
                raise Exception("Couldn't read old chain to resume")
            if self.isave != self.thin and self.resumeLength % (self.isave / self.thin) != 1:
                raise Exception("Old chain has {0} rows, which is not the initial sample plus a multiple of "
                                "isave/thin = {1}".format(self.resumeLength, self.isave // self.thin))
            if self.verbose:
                print("Resuming with", self.resumeLength, "samples from file representing",
                      (self.resumeLength - 1) * self.thin + 1, "original samples")
        elif not self._resume_state:
            if self._writes_primary:
                open(self.fname, "w").close()
            if self.writeHotChains:
                for f in self._hot_fnames[1:]:
                    open(f, "w").close()
        self._make_engine(maxIter)
        self._buffers = None

    # ------------------------------------------------------------------ output ----------------
    def _pull_rows(self):
        """Copy newly recorded rows from the device window into _chain/_lnlike/_lnprob."""
        eng = self._engine
        rows = eng.rows
        if rows <= self._rows_pulled:
            return
        r0 = self._rows_pulled
        n = min(rows - r0, self._chain_all.shape[0] - r0)
        if eng.ntr == 1:
            # T=1 rung only: the device rows stream straight into the (pinned) result arrays (record sink)
            eng.sink_wait()
            self._rows_pulled = min(rows, self._chain_all.shape[0])
            return
        else:
            ch, lnl, lnp = eng.chain(r0, n)
            self._chain_all[r0:r0 + n] = ch[:, 0]
            self._lnlike_all[r0:r0 + n] = lnl[:, 0]
            self._lnprob_all[r0:r0 + n] = lnp[:, 0]
            if self._hot_rows is None:
                self._hot_rows = []
            self._hot_rows.append((r0, ch[:, :, 0].copy(), lnl[:, :, 0].copy(), lnp[:, :, 0].copy()))
        self._rows_pulled = rows
        eng.release_rows(rows)

    _JUMP_NAMES = {_cabi.JUMP_SCAM: "covarianceJumpProposalSCAM", _cabi.JUMP_AM: "covarianceJumpProposalAM",
                   _cabi.JUMP_DE: "DEJump", _cabi.JUMP_PRIOR: "priorDrawJump"}

    def _jump_names(self):
        names = dict(self._JUMP_NAMES)
        for k, f in enumerate(self._ext_jumps):
            names[_cabi.JUMP_EXT0 + k] = f.__name__
        return names

    def _set_counter_summary(self, prop_sum, acc_sum, acc_w0, swap_sum, swap_w0, nsw):
        """jumpDict / naccepted / swap statistics from per-rung sums over walkers ([njumps][T]) and walker 0's values."""
        for jid, name in self._jump_names().items():
            if jid < prop_sum.shape[0] and (name in self.jumpDict or prop_sum[jid, 0] > 0):
                # T=1 rung, summed over walkers (one walker: the reference's rank-0 jumpDict)
                self.jumpDict[name] = [int(prop_sum[jid, 0]), int(acc_sum[jid, 0])]
        self._acc_w0 = acc_w0.sum(axis=0)        # walker 0, per local rung
        self._swap_w0 = np.asarray(swap_w0)
        self.naccepted = acc_sum[:, 0].sum() / float(self.nwalkers) + self._acc_offset
        self.swapProposed = nsw
        self.nswap_accepted = swap_sum[0] / float(self.nwalkers)

    def _pull_counters(self):
        """Synchronous pull of every per-chain counter (the asynchronous path uses ptmcmc_snapshot's summary)."""
        it = self._engine.iteration
        if getattr(self, "_counters_at", None) == it and self._full_counters is not None:
            return
        self._counters_at = it
        if it == 0:  # nothing proposed yet: no device round trip
            shp = (self._Tloc, self.nwalkers, self._engine.njumps)
            prop, acc = np.zeros(shp, dtype=np.int64), np.zeros(shp, dtype=np.int64)
            sw, nsw = np.zeros(shp[:2], dtype=np.int64), 0
        else:
            prop, acc, sw, nsw = self._engine.counters()
        self._full_counters = (prop, acc, sw)
        self._prop, self._acc, self._swap_acc = prop, acc, sw
        # prop / acc are [T][W][njumps] views of the engine's [njumps][T][W] arrays: reduce along the contiguous layout
        pj, aj = prop.transpose(2, 0, 1), acc.transpose(2, 0, 1)
        self._set_counter_summary(pj.sum(axis=2), aj.sum(axis=2), aj[:, :, 0], sw.sum(axis=1), sw[:, 0], nsw)

    @property
    def naccepted_all(self):
        """Accepted MH steps of every chain, [T][W] (fetched from the device on access)."""
        self._pull_counters()
        return np.add.reduce(self._full_counters[1].transpose(2, 0, 1), axis=0)

    @property
    def nswap_accepted_all(self):
        """Accepted swaps with the next-hotter rung of every chain, [T][W] (fetched from the device on access)."""
        self._pull_counters()
        return self._full_counters[2]

    def updateChains(self, p0, lnlike0, lnprob0, iter):
        """The reference's per-iteration buffer/record hook (ref :321-339) is fused into the kernels;
        kept as a no-op for code that calls it around ``PTMCMCOneStep``."""
        return None

    def writeOutput(self, iter):
        """Write chain rows, covariance and jump statistics (ref :341-372)."""
        if self._async:
            self._engine.snapshot(0)
            self._write_boundary(iter, 0)
            return
        self._pull_rows()
        self._pull_counters()
        if iter // self.thin >= self.ind_next_write:
            self._writeToFile(iter)
            if iter > 0 and self.MPIrank == 0:
                self._pull_adapt()
                np.save(self.outDir + "/cov.npy", np.asarray(self.cov))
            if iter > 0 and self.checkpoint:
                tmp = self._state_file + ".tmp.npy"
                np.save(tmp, self._engine.save_state())
                os.replace(tmp, self._state_file)
            self._progress(iter)

    def _progress(self, iter):
        """Progress line (ref :353-372); on resume the percentage is of the new work, as in the reference (:359-367)."""
        if not self.verbose:
            return
        if iter > 0:
            sys.stdout.write("\r")
        done0 = getattr(self, "_resumed_at", 0)
        if done0 and self.Niter > done0:
            percent = (iter - done0) / (self.Niter - done0) * 100
        else:
            percent = iter / self.Niter * 100
        acceptance = self.naccepted / iter if iter > 0 else 0
        sys.stdout.write("Finished %2.2f percent in %f s Acceptance rate = %g" % (percent, time.time() - self.tstart, acceptance))
        sys.stdout.flush()

    def _write_boundary(self, iter, slot):
        """Finish the write of iteration ``iter`` from snapshot ``slot`` (taken on the engine's stream right after that
        iteration): the engine may already be running the next segment."""
        snap = self._engine.snapshot_result(slot)   # waits for the snapshot and every record row before it
        self._rows_pulled = min(iter // self.thin + 1, self._chain_all.shape[0])
        self._full_counters = None
        self._counters_at = None
        self._set_counter_summary(snap["prop_sum"], snap["acc_sum"], snap["acc_w0"], snap["swap_sum"], snap["swap_w0"],
                                  snap["swap_proposed"])
        if iter // self.thin >= self.ind_next_write:
            self._writeToFile(iter)
            if iter > 0:
                self._apply_adapt(snap["cov"], snap["mu"], snap["m2"], snap["U"], snap["S"])
                if self.MPIrank == 0:
                    np.save(self.outDir + "/cov.npy", np.asarray(self.cov))
            self._progress(iter)

    def _apply_adapt(self, cov, mu, m2, U, S):
        if self.MPIrank == 0:
            np.asarray(self.cov)[:, :] = cov  # in place, like the reference (:794)
            self.mu, self.M2 = mu, m2
        uo = so = 0
        for g, grp in enumerate(self.groups):
            n = len(grp)
            self.U[g] = U[uo:uo + n * n].reshape(n, n).copy()
            self.S[g] = S[so:so + n].copy()
            uo += n * n
            so += n

    def _writeToFile(self, iter):
        """Chain file: ndim columns ``%22.22f`` then lnprob, lnlike, acceptance rate, PT swap
        acceptance, rates as of the time of writing (ref :722-766).  Walker 0 is written in the
        reference's layout; all walkers stay available in ``_chain_all``."""
        write_end = iter // self.thin + 1
        rows = range(self.ind_next_write, min(write_end, self._rows_pulled))
        acc_rate = (self._acc_w0[0] + self._acc_offset) / iter if iter > 0 else 0
        pt_acc = 1  # the hottest chain has no hotter partner (ref :737-739)
        if self._lo < self.nchain - 1 and self.swapProposed != 0:
            pt_acc = self._swap_w0[0] / self.swapProposed
        if self._writes_primary:
            tail = "\t%f\t%f\t%f\t%f\n"
            fmt = "\t".join(["%22.22f"] * self.ndim) + tail
            with open(self.fname, "a+") as fh:
                fh.write("".join([fmt % (tuple(self._chain[ind]) + (self._lnprob[ind], self._lnlike[ind], acc_rate, pt_acc))
                                  for ind in rows]))
        if self.writeHotChains and self._hot_rows:
            for t in range(1, self._Tloc):
                a_t = self._acc_w0[t] / iter if iter > 0 else 0
                p_t = 1
                if self._lo + t < self.nchain - 1 and self.swapProposed != 0:
                    p_t = self._swap_w0[t] / self.swapProposed
                with open(self._hot_fnames[t], "a+") as fh:
                    for r0, ch, lnl, lnp in self._hot_rows:
                        for i in range(ch.shape[0]):
                            fh.write("\t".join(["%22.22f" % v for v in ch[i, t]]))
                            fh.write("\t%f\t%f\t%f\t%f\n" % (lnp[i, t], lnl[i, t], a_t, p_t))
            self._hot_rows = []
        self.ind_next_write = write_end
        if self.MPIrank != 0:
            return
        # jump statistics, T=1 chain only (ref :751-766)
        njumps = len(self.propCycle)
        with open(self.outDir + "/jumps.txt", "w") as fout:
            seen = []
            for jump in self.propCycle:
                if jump not in seen:
                    seen.append(jump)
                    fout.write("%s %4.2g\n" % (jump.__name__, self.propCycle.count(jump) / njumps))
        for name in self.jumpDict:
            with open(self.outDir + "/" + name + "_jump.txt", "a+") as fout:
                fout.write("%g\n" % (self.jumpDict[name][1] / max(1, self.jumpDict[name][0])))

    # ------------------------------------------------------------------ sampling --------------
    def _full_p0(self, p0):
        p0 = np.asarray(p0, dtype=np.float64)
        T, W, d = self._Tloc, self.nwalkers, self.ndim
        if p0.shape == (self.nchain, W, d) and self._shard_world > 1:
            return np.ascontiguousarray(p0[self._lo:self._lo + T])
        if p0.shape == (d,):
            return np.broadcast_to(p0, (T, W, d)).copy()
        if p0.shape == (W, d):
            return np.broadcast_to(p0[None], (T, W, d)).copy()
        if p0.shape == (T, W, d):
            return np.ascontiguousarray(p0)
        raise ValueError("p0 must have shape (ndim,), (nwalkers, ndim) or (ntemps, nwalkers, ndim)")

    def _init_state(self, x0):
        """Initial point of every chain (ref :471-487): on the device for device targets; Python targets are evaluated
        here and uploaded (a device prior may be combined with a Python likelihood and vice versa: the engine evaluates
        what it knows and keeps the host's values for the rest)."""
        if self._dev_logl is not None and self._dev_logp is not None:
            self._engine.set_state(x0)
        else:
            lnl, lp = self._host_eval(x0)
            self._engine.set_state_external(x0, lnl, lp)

    def _host_eval(self, x):
        """logp / logl of every chain on the host (plain Python callables), ref :478-487, :605-612."""
        T, W = x.shape[:2]
        lp = np.zeros((T, W))
        lnl = np.zeros((T, W))
        if self.vectorized:
            return self._host_eval_vectorized(x.reshape(T * W, -1), lnl, lp)
        for t in range(T):
            for w in range(W):
                if self._dev_logp is None:
                    lp[t, w] = self.logp(x[t, w])
                if self._dev_logl is None and (self._dev_logp is not None or lp[t, w] != -np.inf):
                    lnl[t, w] = self.logl(x[t, w])
        return lnl, lp

    def _host_eval_vectorized(self, X, lnl, lp):
        """All chains in one call; logl only sees the points the prior accepts (ref :607-612)."""
        if self._dev_logp is None:
            lp.ravel()[:] = np.asarray(self.logp(X), dtype=np.float64)
        if self._dev_logl is None:
            # (with a device prior the host cannot know which points it rejects: logl is evaluated everywhere and the
            # device ignores it where its prior returns -inf)
            ok = (lp.ravel() != -np.inf) if self._dev_logp is None else np.ones(len(X), bool)
            if ok.all():
                lnl.ravel()[:] = np.asarray(self.logl(X), dtype=np.float64)
            elif ok.any():
                lnl.ravel()[ok] = np.asarray(self.logl(X[ok]), dtype=np.float64)
        return lnl, lp

    def _maybe_add_de(self, iter):
        """DE joins the cycle at iteration burn+1 (ref :563-585)."""
        if (iter - 1) == self.burn and self.DEJump not in self.propCycle and self.DEweight > 0:
            if self.verbose:
                print("Adding DE jump with weight {0}".format(self.DEweight))
            self.addProposalToCycle(self.DEJump, self.DEweight)
            self.randomizeProposalCycle()

    def _step_external(self, iter):
        """One iteration with Python callables in the loop (ref :601-612, _jump :1048-1067).  Proposals, jump ids and
        (for custom jumps) the current points arrive in the engine's page-locked buffers; the host fills in what it owns
        -- custom / auxiliary jumps, ``qxy``, Python ``logl`` / ``logp`` -- in place; the rest of the iteration is enqueued
        without waiting.  One host synchronisation per iteration."""
        eng = self._engine
        cb = eng.callback_buffers()
        need_x = bool(self._ext_jumps) or bool(self.aux)
        eng.propose_pinned(want_x=need_x)
        q, jump, qxy, lnl, lp, x = cb["q"], cb["jump"], cb["qxy"], cb["lnl"], cb["lp"], cb["x"]
        T, W = jump.shape
        if need_x:
            qxy[...] = 0.0
        if self.vectorized:
            d = self.ndim
            X, Q, J = x.reshape(T * W, d), q.reshape(T * W, d), jump.ravel()
            betas = np.repeat(1.0 / self._mh_temp, W)
            for k, f in enumerate(self._ext_jumps):
                sel = np.nonzero(J == _cabi.JUMP_EXT0 + k)[0]
                if len(sel) == 0:
                    continue
                if getattr(f, "vectorized", False):
                    Q[sel], qxy.ravel()[sel] = f(X[sel].copy(), iter, betas[sel])
                else:
                    for i in sel:
                        Q[i], qxy.ravel()[i] = f(X[i].copy(), iter, betas[i])
            for aux in self.aux:
                if getattr(aux, "vectorized", False):
                    Q[:], add = aux(X.copy(), Q.copy(), iter, betas)
                    qxy.ravel()[:] += add
                else:
                    for i in range(T * W):
                        Q[i], add = aux(X[i].copy(), Q[i].copy(), iter, betas[i])
                        qxy.ravel()[i] += add
            self._host_eval_vectorized(Q, lnl, lp)
            eng.accept_pinned(q_modified=need_x)
            return
        for t in range(T):
            beta = 1 / self._mh_temp[t]
            for w in range(W):
                jid = jump[t, w]
                if jid >= _cabi.JUMP_EXT0:
                    qq, lq = self._ext_jumps[jid - _cabi.JUMP_EXT0](x[t, w].copy(), iter, beta)
                    q[t, w], qxy[t, w] = qq, lq
                for aux in self.aux:
                    qq, lq = aux(x[t, w].copy(), q[t, w].copy(), iter, beta)
                    q[t, w] = qq
                    qxy[t, w] += lq
                if self._dev_logp is None:
                    lp[t, w] = self.logp(q[t, w])
                if self._dev_logl is None and (self._dev_logp is not None or lp[t, w] != -np.inf):
                    lnl[t, w] = self.logl(q[t, w])
        eng.accept_pinned(q_modified=need_x)

    def _advance(self, n, iter0):
        """Run ``n`` iterations starting after ``iter0``."""
        if not self._external:
            if self._comm is not None:
                from . import distributed

                distributed.run_ladder(self._engine, n, self._comm, self.Tskip)
            elif self.dist_group is not None:
                from . import distributed

                distributed.run(self._engine, n, None if self.dist_group is True else self.dist_group)
            else:
                self._engine.run(n)
            self._maybe_add_de_bulk(iter0, n)
            return
        for it in range(iter0 + 1, iter0 + n + 1):
            self._maybe_add_de(it)
            self._step_external(it)

    def _maybe_add_de_bulk(self, iter0, n):
        if iter0 < self.burn + 1 <= iter0 + n:
            self._maybe_add_de(self.burn + 1)

    def sample(self, p0, Niter, ladder=None, Tmin=1, Tmax=None, Tskip=100, isave=1000, covUpdate=1000,
               SCAMweight=20, AMweight=20, DEweight=20, NUTSweight=20, MALAweight=20, HMCweight=20, burn=10000,
               HMCstepsize=0.1, HMCsteps=300, maxIter=None, thin=10, i0=0, neff=None, writeHotChains=False,
               hotChain=False):
        """Run ``Niter`` iterations from ``p0`` (ref :374-528).  ``p0`` may be one point (used for
        every chain), one per walker, or ``(ntemps, nwalkers, ndim)``."""
        if maxIter is None:
            maxIter = Niter
        if isave % thin != 0:
            raise ValueError("isave = %d is not a multiple of thin =  %d" % (isave, thin))
        if Niter % thin != 0:
            print("Niter = %d is not a multiple of thin = %d.  The last %d samples will be lost"
                  % (Niter, thin, Niter % thin))
        _dbg = os.environ.get("PTMCMC_E2E_DEBUG")
        _t = [time.perf_counter()]
        if i0 == 0:
            self.initialize(Niter, ladder=ladder, Tmin=Tmin, Tmax=Tmax, Tskip=Tskip, isave=isave,
                            covUpdate=covUpdate, SCAMweight=SCAMweight, AMweight=AMweight, DEweight=DEweight,
                            NUTSweight=NUTSweight, MALAweight=MALAweight, HMCweight=HMCweight, burn=burn,
                            HMCstepsize=HMCstepsize, HMCsteps=HMCsteps, maxIter=maxIter, thin=thin, i0=i0,
                            neff=neff, writeHotChains=writeHotChains, hotChain=hotChain)
            x0 = self._full_p0(p0)
            if self._resume_state or self.resumeLength > 0:
                i0 = self._resume(x0)
            else:
                self._init_state(x0)
        elif self._engine is None:
            raise ValueError("i0 != 0 requires a sampler that has already been initialised")
        self.tstart = time.time()
        _t.append(time.perf_counter())
        if not (self._resume_state or self.resumeLength > 0):
            self.writeOutput(i0)  # row 0 (ref :491 -> updateChains -> writeOutput at iter 0)
        _t.append(time.perf_counter())

        iter = i0
        pending, slot = None, 0
        stopped_early = False
        while iter < self.Niter:
            # advance to the next multiple of isave, the reference's write cadence (ref :338-339)
            nxt = min(self.Niter, (iter // self.isave + 1) * self.isave)
            if self.neff:
                nxt = min(nxt, (iter // 1000 + 1) * 1000)  # the effective-sample check runs every 1000 iterations (ref :511)
            self._advance(nxt - iter, iter)
            iter = nxt
            if self.neff and iter % 1000 == 0 and iter > 2 * self.burn and iter < self.Niter and self._neff_reached(iter):
                stopped_early = True   # ref :510-521
                self.Niter = iter
            if iter % self.isave == 0 or iter >= self.Niter:
                if self._async:
                    # the engine returns as soon as the segment is enqueued: snapshot what this write needs in stream
                    # order, and write the PREVIOUS boundary's files while the device works on
                    self._engine.snapshot(slot)
                    if pending is not None:
                        self._write_boundary(*pending)
                    pending, slot = (iter, slot), slot ^ 1
                else:
                    self.writeOutput(iter)
        _t.append(time.perf_counter())
        if pending is not None:
            self._write_boundary(*pending)
        if self._comm is not None:
            self._comm.check(self._engine)  # ladder shards exchanging through peer memory: every message arrived
        _t.append(time.perf_counter())
        self._finish()
        if _dbg:
            _t.append(time.perf_counter())
            sys.stderr.write("sample(): init+set_state %.1f, row0 %.1f, enqueue %.1f, last boundary %.1f, finish %.1f ms\n"
                             % tuple(1e3 * (b - a) for a, b in zip(_t[:-1], _t[1:])))
        if self.verbose:
            print("\nRun Complete with {0} effective samples".format(int(self._last_neff)) if stopped_early
                  else "\nRun Complete")

    def _neff_reached(self, iter):
        """Effective number of samples of walker 0's T=1 chain past burn-in (ref :510-521).  The reference uses the optional
        ``acor`` C extension; the same estimate (integrated autocorrelation time, self-consistent window) is computed here
        with numpy, on the thinned rows, in units of iterations."""
        self._pull_rows()
        r0, r1 = self.burn // self.thin, min(iter // self.thin, self._rows_pulled)
        if r1 - r0 < 16:
            return False
        tau = max(integrated_time(self._chain[r0:r1, k]) for k in range(self.ndim)) * self.thin
        self._last_neff = iter / max(1.0, tau)
        return int(self._last_neff) >= self.neff

    def _resume(self, x0):
        """Bring the engine to the iteration a previous run stopped at; returns that iteration."""
        eng = self._engine
        if self._resume_state:
            # exact continuation: the checkpoint holds every array the step reads, the draws are counter-based
            if self._dev_logl is not None and self._dev_logp is not None:
                eng.set_state(x0)
            else:  # Python targets: the checkpoint replaces every value, nothing needs evaluating
                eng.set_state_external(x0, np.zeros(x0.shape[:2]), np.zeros(x0.shape[:2]))
            eng.load_state(np.load(self._state_file))
            it = eng.iteration
            if self.verbose:
                print("Resuming from engine checkpoint {0} at iteration {1}".format(self._state_file, it))
            self._rows_pulled = self.ind_next_write = it // self.thin + 1
            self._resumed_at = it
            # rows before the checkpoint: walker 0's come back from its chain file (as the reference refills its arrays
            # on resume, ref :291-299); the other walkers' rows are not stored anywhere and read as zeros
            n0 = min(self._rows_pulled, self._chain_all.shape[0])
            for a in (self._chain_all, self._lnlike_all, self._lnprob_all):
                a[:n0] = 0.0
            if os.path.isfile(self.fname):
                try:
                    old = np.loadtxt(self.fname, ndmin=2)[:n0]
                    self._chain_all[:len(old), 0] = old[:, :self.ndim]
                    self._lnprob_all[:len(old), 0] = old[:, -4]
                    self._lnlike_all[:len(old), 0] = old[:, -3]
                except ValueError:
                    pass
            return it
        # the reference's replay (ref :474-476, :591-599): row 0 is the initial point, every stored row is
        # used for `thin` iterations; buffers, covariance and DE history are rebuilt through the normal
        # update path.  Hot rungs have no file here: they are held at the given p0 during the replay.
        rc, R, d, T = self.resumechain, self.resumeLength, self.ndim, self._Tloc
        temp = float(self._mh_temp[0])
        rows_x = np.repeat(x0[None], R, axis=0)               # [R][T][1][d]
        rows_x[:, 0, 0, :] = rc[:, :d]
        rows_lnl = np.zeros((R, T, 1))
        rows_lp = np.zeros((R, T, 1))
        self._init_state(x0)
        st = eng.state()
        rows_lnl[:] = st[1][None]
        rows_lp[:] = st[2][None]
        rows_lnl[:, 0, 0] = rc[:, -3]
        rows_lp[:, 0, 0] = rc[:, -4] - rc[:, -3] / temp       # lnprob = lnlike / temp + logp (ref :487)
        # restart from row 0 exactly as stored
        eng.set_state_external(rows_x[0], rows_lnl[0], rows_lp[0])
        total, done = R * self.thin - 1, 0
        chunk = max(self.thin, (self.isave // self.thin) * self.thin)
        while done < total:
            n = min(chunk, total - done)
            first = (done + 1) // self.thin
            last = (done + n) // self.thin
            eng.replay(n, self.thin, rows_x[first:last + 1], rows_lnl[first:last + 1], rows_lp[first:last + 1])
            done += n
            self._pull_rows()
        self._maybe_add_de_bulk(0, total)                      # DE joined the cycle during the replay (ref :563-585)
        self.ind_next_write = R                                # these rows are already in the file (ref :476)
        self._acc_offset = total * float(rc[-1, -2])           # ref :599
        self._resumed_at = total
        return total

    def _finish(self):
        if not self._async:
            self._pull_rows()
            self._pull_counters()
            if self.MPIrank == 0:
                self._pull_adapt()
            else:
                self._pull_factor()
        elif self._engine.iteration == 0:
            self._pull_counters()
        for a in (self._chain_all, self._lnlike_all, self._lnprob_all):
            a[self._rows_pulled:] = 0.0  # rows never reached (the reference's arrays start as zeros, :208-212)
        self._buffers = None

    def close(self):
        """Release the device engine (the result arrays stay valid)."""
        if self._engine is not None:
            self._engine.sync()
            self._engine.clear_sink()
            self._engine.close()
            self._engine = None

    def _fetch_buffers(self):
        if self._buffers is None:
            am, de = self._engine.buffers()
            self._buffers = (am[:, 0], de[:, 0]) if self.nwalkers == 1 else (am, de)
        return self._buffers

    @property
    def _AMbuffer(self):
        """ref _AMbuffer (:220): the last covUpdate T=1 samples, fetched from the device on access."""
        return self._fetch_buffers()[0]

    @property
    def _DEbuffer(self):
        """ref _DEbuffer (:221): the DE history, fetched from the device on access."""
        return self._fetch_buffers()[1]

    def PTMCMCOneStep(self, p0, lnlike0, lnprob0, iter):
        """One iteration of every chain on the engine (ref :530-629); returns walker 0's T=1 state."""
        if self._engine is None:
            raise ValueError("call initialize() / sample() first")
        self._advance(1, self._engine.iteration)
        x, lnl, lp, lnp = self._engine.state()
        return x[0, 0], lnl[0, 0], lnp[0, 0]

    def write_walker_chain(self, walker, fname=None):
        """Write the recorded T=1 chain of ``walker`` in the reference's chain-file format
        (ref :736-745; ``sample`` itself writes walker 0 only).  Acceptance columns are the walker's final
        rates, as in a file written in one block."""
        if fname is None:
            fname = os.path.join(self.outDir, "chain_1_walker%d.txt" % walker)
        if getattr(self, "_resumed_at", 0) and walker != 0:
            raise ValueError("rows before the checkpoint exist for walker 0 only (its chain file); walker %d's are not stored" % walker)
        it = self._engine.iteration
        acc = (self.naccepted_all[0, walker] / it) if it > 0 else 0
        pt_acc = 1
        if self._lo < self.nchain - 1 and self.swapProposed != 0:
            pt_acc = self.nswap_accepted_all[0, walker] / self.swapProposed
        with open(fname, "w") as fh:
            for ind in range(self._rows_pulled):
                fh.write("\t".join(["%22.22f" % v for v in self._chain_all[ind, walker]]))
                fh.write("\t%f\t%f\t%f\t%f\n" % (self._lnprob_all[ind, walker], self._lnlike_all[ind, walker], acc, pt_acc))
        return fname

    # convenience accessors beyond the reference ------------------------------------------------
    @property
    def engine(self):
        return self._engine

    def get_state(self):
        """Current ``(x[T][W][d], lnlike[T][W], lnprior[T][W], lnprob[T][W])`` of every chain."""
        return self._engine.state()
