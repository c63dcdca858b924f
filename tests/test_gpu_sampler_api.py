"""The reference-facing API on the GPU: tests written like the reference's own
(tests/test_simple.py) plus statistical parity against the reference's golden bands."""
import os
import shutil

import numpy as np
import pytest

from ptmcmcsampler_b200 import PTMCMCSampler, nompi4py as MPIDUMMY
from ptmcmcsampler_b200.likelihoods import GaussianLikelihood, UniformPrior

from _helpers import load

pytestmark = pytest.mark.gpu


class GaussianLikelihoodPy(object):
    """The reference's test fixture (ref tests/test_simple.py:14-41), plain Python callables."""

    def __init__(self, ndim=2, pmin=-10, pmax=10):
        self.a = np.ones(ndim) * pmin
        self.b = np.ones(ndim) * pmax
        self.mu = np.random.uniform(pmin, pmax, ndim)
        cov = 0.5 - np.random.rand(ndim**2).reshape((ndim, ndim))
        cov = np.triu(cov)
        cov += cov.T - np.diag(cov.diagonal())
        self.cov = np.dot(cov, cov)
        self.icov = np.linalg.inv(self.cov)

    def lnlikefn(self, x):
        diff = x - self.mu
        return -np.dot(diff, np.dot(self.icov, diff)) / 2.0

    def lnpriorfn(self, x):
        if np.all(self.a <= x) and np.all(self.b >= x):
            return 0.0
        return -np.inf


class UniformJump(object):
    def __init__(self, pmin, pmax):
        self.pmin, self.pmax = pmin, pmax

    def jump(self, x, it, beta):
        return np.random.uniform(self.pmin, self.pmax, len(x)), 0


@pytest.fixture
def outdir(tmp_path):
    d = str(tmp_path / "chains")
    yield d
    shutil.rmtree(d, ignore_errors=True)


def test_simple_python_callables_and_custom_jump(outdir):
    """ref tests/test_simple.py::test_simple, shortened: Python logl/logp + UniformJump plugin."""
    np.random.seed(0)
    ndim, pmin, pmax = 20, 0.0, 10.0
    glo = GaussianLikelihoodPy(ndim=ndim, pmin=pmin, pmax=pmax)
    p0 = np.random.uniform(pmin, pmax, ndim)
    cov = np.eye(ndim) * 0.1**2
    sampler = PTMCMCSampler.PTSampler(ndim, glo.lnlikefn, glo.lnpriorfn, np.copy(cov), outDir=outdir,
                                      comm=MPIDUMMY.COMM_WORLD, verbose=False, seed=1)
    ujump = UniformJump(pmin, pmax)
    sampler.addProposalToCycle(ujump.jump, 5)
    sampler.sample(p0, 1500, burn=500, thin=1, covUpdate=500, SCAMweight=20, AMweight=20, DEweight=20)
    assert os.path.isfile(os.path.join(outdir, "chain_1.txt"))
    data = np.loadtxt(os.path.join(outdir, "chain_1.txt"))
    assert data.shape == (1501, ndim + 4)
    assert np.allclose(data[:, :ndim], sampler._chain, atol=0)
    for name in ("covarianceJumpProposalSCAM", "covarianceJumpProposalAM", "DEJump", "jump"):
        assert os.path.isfile(os.path.join(outdir, name + "_jump.txt"))
        assert sampler.jumpDict[name][0] > 0
    assert sum(v[0] for v in sampler.jumpDict.values()) == 1500
    assert os.path.isfile(os.path.join(outdir, "cov.npy")) and os.path.isfile(os.path.join(outdir, "jumps.txt"))
    # lnlike column is the user's function of the recorded point
    assert np.allclose(data[-1, ndim + 1], glo.lnlikefn(data[-1, :ndim]), rtol=1e-6, atol=1e-5)
    assert sampler._AMbuffer.shape == (500, ndim) and sampler._DEbuffer.shape == (500, ndim)


def test_error_conventions(outdir):
    lk = GaussianLikelihood(np.zeros(3), cov=np.eye(3))
    pr = UniformPrior(-5, 5)
    s = PTMCMCSampler.PTSampler(3, lk, pr, np.eye(3), outDir=outdir, verbose=False, seed=2)
    with pytest.raises(ValueError):  # ref :434-435
        s.sample(np.zeros(3), 100, isave=15, thin=10)
    s = PTMCMCSampler.PTSampler(3, lk, pr, np.eye(3), outDir=outdir, verbose=False, seed=2)
    with pytest.raises(ValueError):  # ref :267-268
        s.sample(np.zeros(3), 100, SCAMweight=0, AMweight=0)
    s = PTMCMCSampler.PTSampler(3, lk, pr, np.eye(3), outDir=outdir, verbose=False, seed=2)
    with pytest.raises(ValueError):  # covUpdate > burn: ref :817 broadcast error at the first DE update
        s.sample(np.zeros(3), 100, covUpdate=50, burn=20, thin=1, isave=10)


def _run_device(g, W, T, N, outdir, seed, thin=1):
    d = int(g["d"])
    lk = GaussianLikelihood(g["pb_mu"], icov=g["pb_icov"])
    pr = UniformPrior(g["pb_lo"], g["pb_hi"])
    s = PTMCMCSampler.PTSampler(d, lk, pr, np.eye(d) * 0.01, outDir=outdir, verbose=False, seed=seed, ntemps=T,
                                nwalkers=W)
    lo = max(float(g["pb_lo"][0]), 0.0)
    hi = min(float(g["pb_hi"][0]), 10.0)
    p0 = np.random.default_rng(seed).uniform(lo, hi, (T, W, d))
    s.sample(p0, N, burn=int(g["kw_burn"]), thin=thin, covUpdate=int(g["kw_covUpdate"]), SCAMweight=20, AMweight=20,
             DEweight=20, isave=N, Tskip=int(g["kw_Tskip"]), writeHotChains=T > 1)
    return s


def _band(ref_values, engine_value, nsig=3.0, floor=0.0):
    """engine value inside mean +- nsig * (standard error of the reference's repeats + floor)."""
    m = ref_values.mean(axis=0)
    se = ref_values.std(axis=0, ddof=1) / np.sqrt(ref_values.shape[0])
    return np.abs(engine_value - m) <= nsig * np.sqrt(se**2 + floor**2)


@pytest.mark.parametrize("name", ["stats_t1_d8", "stats_t1_d8_box"])
def test_posterior_matches_reference_band(name, outdir):
    """Posterior mean / variance / acceptance of the T=1 chain against the reference's own runs
    (mean over repeats +- 3 standard errors; the engine's error is negligible with 4096 walkers)."""
    g = load(name)
    N, W, thin = int(g["N"]), 1024, 10   # same length as the reference runs
    s = _run_device(g, W, 1, N, outdir, seed=31, thin=thin)
    x = s._chain_all[int(N * float(g["burn_frac"])) // thin:]  # [n][W][d]
    mean, var = x.mean(axis=(0, 1)), x.var(axis=(0, 1))
    assert np.all(_band(g["means"][:, 0], mean, floor=0.01)), (mean, g["means"][:, 0].mean(0))
    assert np.all(_band(g["vars"][:, 0], var, floor=0.02 * var.max()))
    acc = s.naccepted / N
    assert _band(g["acc"][:, 0], acc, floor=0.01)
    jacc = np.array([s.jumpDict[k][1] / s.jumpDict[k][0] for k in
                     ("covarianceJumpProposalSCAM", "covarianceJumpProposalAM", "DEJump")])
    assert np.all(_band(g["jump_acc"][:, 0], jacc, floor=0.02)), (jacc, g["jump_acc"][:, 0].mean(0))
    if name == "stats_t1_d8":
        # analytic known answer: untruncated Gaussian target
        cov = g["pb_cov"]
        assert np.allclose(mean, g["pb_mu"], atol=0.03)
        assert np.allclose(var, np.diag(cov), rtol=0.03)


def test_tempered_posteriors_match_reference_band(outdir):
    g = load("stats_t4_d8")
    T, N, W = 4, int(g["N"]), 1024
    s = _run_device(g, W, T, N, outdir, seed=41, thin=10)
    x, lnl, lp, lnp = s.get_state()
    # cov at temperature T is T * Sigma for the untruncated Gaussian (SURVEY.md 8c KAT 1)
    var_ratio = x.var(axis=1) / np.diag(g["pb_cov"])[None, :]
    assert np.allclose(var_ratio.mean(axis=1), g["ladder"], rtol=0.06), var_ratio.mean(axis=1)
    swap = s.nswap_accepted_all.mean(axis=1) / s.swapProposed
    assert np.all(_band(g["swap"][:, :T - 1], swap[:T - 1], floor=0.02)), (swap, g["swap"].mean(0))
    acc = s.naccepted_all.mean(axis=1) / N
    assert np.all(_band(g["acc"], acc, floor=0.015)), (acc, g["acc"].mean(0))
    for t in range(1, T):
        assert os.path.isfile(os.path.join(outdir, "chain_{0}.txt".format(s.ladder[t])))


def _small_sampler(outdir, W, T, seed=5, **kw):
    d = 6
    rng = np.random.default_rng(3)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    lk, pr = GaussianLikelihood(np.full(d, 5.0), cov=cov), UniformPrior(-50, 60)
    s = PTMCMCSampler.PTSampler(d, lk, pr, np.eye(d) * 0.01, outDir=outdir, verbose=False, seed=seed, ntemps=T,
                                nwalkers=W, **kw)
    p0 = np.random.default_rng(5 if seed is None else seed).uniform(0, 10, (T, W, d))
    return s, p0


def test_vectorized_python_callables_match_per_chain_calls(tmp_path):
    """vectorized=True: logl / logp / custom jump take every chain at once; same chain as the per-chain protocol."""
    d, W, T = 5, 12, 2
    rng = np.random.default_rng(0)
    A = rng.standard_normal((d, d))
    icov = np.linalg.inv(A @ A.T + 0.5 * np.eye(d))
    mu = np.full(d, 5.0)

    def logl1(x):
        return -0.5 * (x - mu) @ icov @ (x - mu)

    def logp1(x):
        return 0.0 if np.all((x >= 0) & (x <= 10)) else -np.inf

    def loglv(X):
        D = X - mu
        return -0.5 * np.einsum("ni,ij,nj->n", D, icov, D)

    def logpv(X):
        return np.where(np.all((X >= 0) & (X <= 10), axis=1), 0.0, -np.inf)

    def jump1(x, it, beta):
        return x + 0.1 * np.sin(np.arange(d) + it), 0.0

    def jumpv(X, it, beta):
        return X + 0.1 * np.sin(np.arange(d) + it)[None, :], np.zeros(len(X))

    jumpv.vectorized = True
    jumpv.__name__ = "jump1"
    p0 = rng.uniform(3, 7, (T, W, d))
    runs = []
    for vec in (False, True):
        s = PTMCMCSampler.PTSampler(d, loglv if vec else logl1, logpv if vec else logp1, np.eye(d) * 0.05,
                                    outDir=str(tmp_path / ("v%d" % vec)), verbose=False, seed=4, ntemps=T, nwalkers=W,
                                    vectorized=vec)
        s.addProposalToCycle(jumpv if vec else jump1, 10)
        s.sample(p0, 150, burn=50, covUpdate=50, Tskip=10, thin=1, isave=50)
        runs.append(s)
    assert np.allclose(runs[0]._chain_all, runs[1]._chain_all, rtol=1e-12, atol=1e-12)
    assert runs[0].jumpDict["jump1"] == runs[1].jumpDict["jump1"] and runs[0].jumpDict["jump1"][0] > 0


def test_walker_chain_export(tmp_path):
    s, p0 = _small_sampler(str(tmp_path / "w"), 8, 2)
    s.sample(p0, 200, burn=100, covUpdate=50, Tskip=10, thin=5, isave=100)
    f = s.write_walker_chain(3)
    data = np.loadtxt(f)
    assert data.shape == (41, 10) and np.allclose(data[:, :6], s._chain_all[:, 3], atol=0)
    w0 = np.loadtxt(os.path.join(str(tmp_path / "w"), "chain_1.0.txt"))
    assert np.array_equal(np.loadtxt(s.write_walker_chain(0))[:, :8], w0[:, :8])


def test_checkpoint_resume_is_exact(tmp_path):
    """Engine checkpoint: a run stopped at 300 and resumed to 600 equals the straight 600-iteration run bit
    for bit (the reference cannot do this: it does not save generator state)."""
    kw = dict(burn=100, covUpdate=50, Tskip=10, thin=5, isave=100)
    a, p0 = _small_sampler(str(tmp_path / "a"), 16, 3)
    a.sample(p0, 600, **kw)
    b, _ = _small_sampler(str(tmp_path / "b"), 16, 3, checkpoint=True)
    b.sample(p0, 300, **kw)
    assert os.path.isfile(os.path.join(str(tmp_path / "b"), "engine_state.npy"))
    c, _ = _small_sampler(str(tmp_path / "b"), 16, 3, resume=True)
    c.sample(p0, 600, **kw)
    for x, y in zip(a.get_state(), c.get_state()):
        assert np.array_equal(x, y)
    assert np.array_equal(a._chain_all[61:], c._chain_all[61:])
    assert np.array_equal(np.asarray(a.cov), np.asarray(c.cov))
    fa = np.loadtxt(os.path.join(str(tmp_path / "a"), "chain_1.0.txt"))
    fc = np.loadtxt(os.path.join(str(tmp_path / "b"), "chain_1.0.txt"))
    assert fa.shape == fc.shape == (121, 10) and np.array_equal(fa[:, :8], fc[:, :8])
    # rows before the checkpoint: walker 0's come back from its chain file, nothing stale is exposed
    assert np.array_equal(c._chain_all[:61, 0], fc[:61, :6]) and np.all(c._chain_all[:61, 1:] == 0)
    with pytest.raises(ValueError):
        c.write_walker_chain(3)


def test_checkpoint_resume_recovers_the_seed_and_rejects_a_changed_schedule(tmp_path):
    """seed=None on resume takes the seed from the checkpoint (the draws are keyed by it); a checkpoint written under another
    thin / Tskip / ladder is refused instead of silently continuing a different run."""
    from ptmcmcsampler_b200 import _cabi

    kw = dict(burn=100, covUpdate=50, Tskip=10, thin=5, isave=100)
    a, p0 = _small_sampler(str(tmp_path / "a"), 8, 3, seed=77)
    a.sample(p0, 400, **kw)
    b, _ = _small_sampler(str(tmp_path / "b"), 8, 3, seed=77, checkpoint=True)
    b.sample(p0, 200, **kw)
    c, _ = _small_sampler(str(tmp_path / "b"), 8, 3, seed=None, resume=True)
    assert c.seed == 77
    c.sample(p0, 400, **kw)
    for x, y in zip(a.get_state(), c.get_state()):
        assert np.array_equal(x, y)
    bad, _ = _small_sampler(str(tmp_path / "b"), 8, 3, seed=78, resume=True)
    with pytest.raises(_cabi.EngineError):
        bad.sample(p0, 600, **kw)
    bad2, _ = _small_sampler(str(tmp_path / "b"), 8, 3, seed=77, resume=True)
    with pytest.raises(_cabi.EngineError):
        bad2.sample(p0, 600, **dict(kw, Tskip=20))


def test_reference_style_resume_replays_the_chain_file(tmp_path):
    """resume=True without a checkpoint: the chain file is replayed, each row for `thin` iterations, through
    the normal buffer / covariance / DE path (ref :474-476, :591-599), then sampling continues."""
    out = str(tmp_path / "r")
    kw = dict(burn=100, covUpdate=50, Tskip=10, thin=2, isave=100)
    a, p0 = _small_sampler(out, 1, 1)
    a.sample(p0, 400, **kw)
    first = np.loadtxt(os.path.join(out, "chain_1.txt"))
    assert first.shape[0] == 201
    b, _ = _small_sampler(out, 1, 1, resume=True, seed=9)
    b.sample(p0, 800, **kw)
    data = np.loadtxt(os.path.join(out, "chain_1.txt"))
    assert data.shape[0] == 401 and np.array_equal(data[:201], first)
    assert b.resumeLength == 201
    assert np.allclose(b._chain[:201], first[:, :6], atol=0)          # replayed rows are in _chain
    assert np.allclose(b._chain[201:], data[201:, :6], atol=0)
    # the adaptive state was rebuilt from the replayed rows: the DE history holds file rows
    am, de = b._engine.buffers()
    assert np.isfinite(np.asarray(b.cov)).all() and np.abs(de).sum() > 0
    with pytest.raises(Exception):  # misaligned file (ref :301-309)
        bad = np.loadtxt(os.path.join(out, "chain_1.txt"))[:-3]
        np.savetxt(os.path.join(out, "chain_1.txt"), bad)
        c, _ = _small_sampler(out, 1, 1, resume=True)
        c.sample(p0, 1000, **kw)


def _py_targets(d, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    mu, icov = 5.0 * np.ones(d), np.linalg.inv(cov)
    lo, hi = -50.0 * np.ones(d), 60.0 * np.ones(d)

    def logl(x):
        r = x - mu
        return -0.5 * float(r @ (icov @ r))

    def logp(x):
        return 0.0 if np.all(lo <= x) and np.all(x <= hi) else -np.inf

    return mu, icov, logl, logp


def test_device_and_python_targets_can_be_mixed(tmp_path):
    """A device prior with a Python likelihood, and the reverse, walk the trajectory of the all-device sampler."""
    d, W, T, N = 6, 5, 3, 200
    mu, icov, logl, logp = _py_targets(d, 4)
    p0 = np.random.default_rng(3).uniform(0, 10, (T, W, d))
    kw = dict(burn=100, covUpdate=50, Tskip=10, thin=5, isave=100)
    runs = []
    for name, lk, pr in (("dd", GaussianLikelihood(mu, icov=icov), UniformPrior(-50.0, 60.0)),
                         ("pd", logl, UniformPrior(-50.0, 60.0)), ("dp", GaussianLikelihood(mu, icov=icov), logp)):
        s = PTMCMCSampler.PTSampler(d, lk, pr, 0.05 * np.eye(d), outDir=str(tmp_path / name), verbose=False, seed=21,
                                    ntemps=T, nwalkers=W)
        s.sample(p0, N, **kw)
        runs.append(s)
    for s in runs[1:]:
        assert np.allclose(s._chain_all, runs[0]._chain_all, rtol=1e-9, atol=1e-9)
        assert np.allclose(s._lnlike_all, runs[0]._lnlike_all, rtol=1e-9, atol=1e-9)
        assert s.jumpDict == runs[0].jumpDict


def test_resume_with_python_callables(tmp_path):
    """Reference-style resume (replay of the chain file, ref :290-319, :591-599) and the engine checkpoint with Python
    logl / logp, as the reference allows for any callable."""
    d, N = 4, 400
    mu, icov, logl, logp = _py_targets(d, 7)
    p0 = np.random.default_rng(5).uniform(0, 10, d)
    kw = dict(burn=100, covUpdate=50, thin=2, isave=100)
    out = str(tmp_path / "replay")
    a = PTMCMCSampler.PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=out, verbose=False, seed=3)
    a.sample(p0, N, **kw)
    first = np.loadtxt(os.path.join(out, "chain_1.txt"))
    b = PTMCMCSampler.PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=out, verbose=False, seed=4, resume=True)
    b.sample(p0, 2 * N, **kw)
    data = np.loadtxt(os.path.join(out, "chain_1.txt"))
    assert data.shape[0] == 2 * N // 2 + 1 and np.array_equal(data[:len(first)], first) and b.resumeLength == len(first)
    # exact continuation from an engine checkpoint
    out2 = str(tmp_path / "ckpt")
    full = PTMCMCSampler.PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=str(tmp_path / "full"), verbose=False, seed=3)
    full.sample(p0, 2 * N, **kw)
    c = PTMCMCSampler.PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=out2, verbose=False, seed=3, checkpoint=True)
    c.sample(p0, N, **kw)
    e = PTMCMCSampler.PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=out2, verbose=False, seed=3, resume=True)
    e.sample(p0, 2 * N, **kw)
    for x, y in zip(full.get_state(), e.get_state()):
        assert np.array_equal(x, y)
    assert np.array_equal(np.loadtxt(os.path.join(str(tmp_path / "full"), "chain_1.txt"))[:, :d],
                          np.loadtxt(os.path.join(out2, "chain_1.txt"))[:, :d])
