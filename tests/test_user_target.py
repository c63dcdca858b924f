"""User targets as CUDA source (NVRTC): compile-only checks on the CPU, parity on the GPU -- the same likelihood / prior as
Python callables (the oracle's callback hooks, and PTSampler's host path) and as device source must walk the same
trajectory."""
import numpy as np
import pytest

from oracle import oracle as orc
from ptmcmcsampler_b200 import _cabi
from ptmcmcsampler_b200.likelihoods import SourceLikelihood, SourcePrior

GAUSS_SRC = """
// par = mu[ndim], icov[ndim * ndim] (row-major)
double quad(const double *r, const double *A, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double row = 0.0;
        for (int j = 0; j < n; ++j) row += A[i * n + j] * r[j];
        acc += r[i] * row;
    }
    return acc;
}
double user_logl(const double *x, int ndim, const double *par) {
    double r[64];
    for (int i = 0; i < ndim; ++i) r[i] = x[i] - par[i];
    return -0.5 * quad(r, par + ndim, ndim);
}
"""
BOX_SRC = """
// par = lo[ndim], hi[ndim]
double user_logp(const double *x, int ndim, const double *par) {
    for (int i = 0; i < ndim; ++i)
        if (!(par[i] <= x[i] && x[i] <= par[ndim + i])) return -INFINITY;
    return 0.0;
}
"""


def problem(d, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    return rng.uniform(2, 8, d), np.linalg.inv(cov)


def test_sources_compile_for_sm100a_without_a_device():
    assert _cabi.user_compile_check(GAUSS_SRC, BOX_SRC) > 0
    assert _cabi.user_compile_check(GAUSS_SRC, None) > 0
    assert _cabi.user_compile_check(None, BOX_SRC) > 0


def test_compile_error_carries_the_nvrtc_log():
    with pytest.raises(ValueError) as err:
        _cabi.user_compile_check("double user_logl(const double *x, int n, const double *p) { return nonsense(x); }")
    assert "nonsense" in str(err.value) and "NVRTC" in str(err.value)
    with pytest.raises(ValueError):  # wrong signature: the kernels call user_logl(const double *, int, const double *)
        _cabi.user_compile_check("double user_logl(double x) { return x; }")


@pytest.mark.gpu
@pytest.mark.parametrize("d,W,T", [(6, 40, 3), (20, 33, 2), (3, 5, 1)])
def test_source_targets_match_oracle_callbacks(d, W, T):
    mu, icov = problem(d, d)
    lo, hi = 3.0 * np.ones(d) if d == 6 else -50.0 * np.ones(d), 7.0 * np.ones(d) if d == 6 else 60.0 * np.ones(d)

    def logl(x):
        r = x - mu
        return -0.5 * float(r @ (icov @ r))

    def logp(x):
        return 0.0 if np.all(lo <= x) and np.all(x <= hi) else -np.inf

    niter, tskip = 240, 10
    ladder = orc.temperature_ladder(d, T)
    kw = dict(seed=5 + d, cov_update=40, burn=80, tskip=tskip, thin=5)
    o = orc.Oracle(d, W, T, 0.05 * np.eye(d), ladder=ladder, logl_kind=orc.LOGL_EXTERNAL, logp_kind=orc.LOGP_EXTERNAL,
                   ext_logl=logl, ext_logp=logp, max_rows=niter // 5 + 1, **kw)
    g = _cabi.Engine(d, W, T, 0.05 * np.eye(d), ladder, logl_kind=_cabi.LOGL_USER, logp_kind=_cabi.LOGP_USER,
                     logl_source=GAUSS_SRC, logp_source=BOX_SRC, logl_user_params=np.concatenate([mu, icov.ravel()]),
                     logp_user_params=np.concatenate([lo, hi]), record_rows=niter // 5 + 1, trace_iters=niter, **kw)
    assert "NVRTC" in g.mh_kernel_name
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    otrace, oswap = o.set_trace(niter, niter // tskip if T > 1 else 0)
    o.set_state(x0)
    g.set_state(x0)
    o.run(niter)
    g.run(niter)
    gtrace, gswap = g.trace(niter, niter // tskip if T > 1 else 0)
    assert np.array_equal(gtrace, otrace)
    if T > 1:
        assert np.array_equal(gswap, oswap[:niter // tskip])
    for a, b in zip(o.state(), g.state()):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9, equal_nan=True)
    assert np.allclose(o.adapt()[0], g.adapt()[0], rtol=1e-8, atol=1e-12)


@pytest.mark.gpu
def test_sampler_source_target_equals_python_callables(tmp_path):
    """PTSampler with the target as device source against the same sampler with Python callables on the host path (one
    host round trip per iteration): same seed, same trajectory."""
    from ptmcmcsampler_b200 import PTSampler

    d, W, T, N = 5, 6, 3, 300
    mu, icov = problem(d, 2)
    lo, hi = -50.0 * np.ones(d), 60.0 * np.ones(d)

    def logl(x):
        r = x - mu
        return -0.5 * float(r @ (icov @ r))

    def logp(x):
        return 0.0 if np.all(lo <= x) and np.all(x <= hi) else -np.inf

    p0 = np.random.default_rng(3).uniform(0, 10, (T, W, d))
    kw = dict(burn=100, covUpdate=50, Tskip=10, thin=5, isave=100)
    a = PTSampler(d, SourceLikelihood(GAUSS_SRC, params=np.concatenate([mu, icov.ravel()])),
                  SourcePrior(BOX_SRC, params=np.concatenate([lo, hi])), 0.05 * np.eye(d), outDir=str(tmp_path / "a"),
                  verbose=False, seed=11, ntemps=T, nwalkers=W)
    a.sample(p0, N, **kw)
    b = PTSampler(d, logl, logp, 0.05 * np.eye(d), outDir=str(tmp_path / "b"), verbose=False, seed=11, ntemps=T, nwalkers=W)
    b.sample(p0, N, **kw)
    assert np.allclose(a._chain_all, b._chain_all, rtol=1e-9, atol=1e-9)
    assert np.allclose(a._lnlike_all, b._lnlike_all, rtol=1e-9, atol=1e-9)
    assert a.jumpDict == b.jumpDict and a.swapProposed == b.swapProposed
    assert open(str(tmp_path / "a" / "chain_1.0.txt")).read() == open(str(tmp_path / "b" / "chain_1.0.txt")).read()
