"""Gradient proposals (ref nutsjump.py) and the effective-sample stop (ref PTMCMCSampler.py:510-521).

CPU part: each proposal, driven by a ten-line Metropolis-Hastings loop that applies ``qxy`` exactly as the sampler does
(ref :614-616), must sample a correlated Gaussian correctly -- this is what pins the Hastings corrections.  GPU part:
the reference-facing call with ``logl_grad`` / ``logp_grad`` (ref tests/test_nuts.py)."""
import numpy as np
import pytest

from ptmcmcsampler_b200 import nutsjump
from ptmcmcsampler_b200.PTMCMCSampler import integrated_time


def target(d=3, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    mu, icov = np.arange(d, dtype=float), np.linalg.inv(cov)

    def logl_grad(x):
        r = x - mu
        g = -icov @ r
        return 0.5 * float(r @ g), g

    def logp_grad(x):
        return 0.0, np.zeros(d)

    return mu, cov, logl_grad, logp_grad


def mh_chain(jump, logl_grad, x0, n, beta, rng):
    """Metropolis-Hastings with a plugin proposal, the sampler's accept rule: diff = lnprob(q) - lnprob(x) + qxy."""
    x, lp = x0.copy(), beta * logl_grad(x0)[0]
    out, acc = np.empty((n, len(x0))), 0
    for it in range(n):
        q, qxy = jump(x, it, beta)
        lq = beta * logl_grad(q)[0]
        if lq - lp + qxy > np.log(rng.random()):
            x, lp, acc = q, lq, acc + 1
        out[it] = x
    return out, acc / n


@pytest.mark.parametrize("kind,n,beta", [("mala", 40000, 1.0), ("hmc", 6000, 1.0), ("hmc", 6000, 0.5), ("nuts", 2500, 1.0)])
def test_gradient_proposals_sample_the_target(kind, n, beta):
    mu, cov, logl_grad, logp_grad = target()
    rng = np.random.default_rng(5)
    mm = 0.6 * cov + 0.1 * np.eye(3)   # a proposal covariance that is not the target's
    if kind == "mala":
        j = nutsjump.MALAJump(logl_grad, logp_grad, mm, rng=rng)
    elif kind == "hmc":
        j = nutsjump.HMCJump(logl_grad, logp_grad, mm, stepsize=0.25, nminsteps=3, nmaxsteps=12, rng=rng)
    else:
        j = nutsjump.NUTSJump(logl_grad, logp_grad, mm, nburn=300, rng=rng)
    chain, acc = mh_chain(j, logl_grad, mu + 1.0, n, beta, rng)
    x = chain[n // 5:]
    tau = max(integrated_time(x[:, k]) for k in range(3))
    neff = len(x) / tau
    assert acc > (0.99 if kind == "nuts" else 0.3)
    assert np.all(np.abs(x.mean(0) - mu) < 5 * np.sqrt(np.diag(cov) / beta / neff)), (x.mean(0), neff)
    assert np.allclose(np.cov(x.T), cov / beta, rtol=0, atol=8 * np.abs(cov / beta).max() * np.sqrt(2.0 / neff)), neff


def test_hmc_batched_equals_per_chain_statistics():
    """vectorized protocol: all chains integrate in lock step with per-chain trajectory lengths; energy error stays small."""
    mu, cov, logl_grad, logp_grad = target(4, 2)
    icov = np.linalg.inv(cov)

    def logl_grad_b(X):
        R = X - mu
        G = -R @ icov
        return 0.5 * np.einsum("ni,ni->n", R, G), G

    def logp_grad_b(X):
        return np.zeros(len(X)), np.zeros_like(X)

    rng = np.random.default_rng(3)
    j = nutsjump.HMCJump(logl_grad_b, logp_grad_b, cov, stepsize=0.1, nminsteps=5, nmaxsteps=20, rng=rng, batched_gradients=True)
    X = rng.multivariate_normal(mu, cov, 64)
    Q, qxy = j(X, 0, np.ones(64))
    assert Q.shape == X.shape and qxy.shape == (64,)
    dH = logl_grad_b(Q)[0] - logl_grad_b(X)[0] + qxy           # energy error of a leapfrog trajectory: small
    assert np.all(np.abs(dH) < 0.05) and np.abs(Q - X).max() > 0.1
    q1, e1 = j(X[0], 0, 1.0)
    assert q1.shape == (4,) and np.isscalar(e1)


def test_hmc_compat_reproduces_the_reference_quirks():
    """compat=True: the trajectory stops after the first leapfrog and qxy is the full Hamiltonian difference (ref :283-287)."""
    mu, cov, logl_grad, logp_grad = target()
    calls = [0]

    def counting(x):
        calls[0] += 1
        return logl_grad(x)

    j = nutsjump.HMCJump(counting, logp_grad, cov, stepsize=0.1, nminsteps=10, nmaxsteps=20, compat=True,
                         rng=np.random.default_rng(1))
    j(mu + 0.3, 0, 1.0)
    assert calls[0] == 2   # the start point and ONE leapfrog


def test_integrated_time_of_an_ar1_series():
    rng = np.random.default_rng(0)
    phi, n = 0.9, 200000
    x = np.empty(n)
    x[0] = 0.0
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    assert abs(integrated_time(x) - (1 + phi) / (1 - phi)) < 2.0
    assert integrated_time(rng.standard_normal(10000)) < 1.3


@pytest.mark.gpu
def test_sampler_with_gradients_and_neff(tmp_path):
    """ref tests/test_nuts.py shape: logl_grad / logp_grad given, HMC / NUTS / MALA in the cycle next to SCAM / AM / DE;
    neff stops the run early once the T=1 chain holds enough effective samples."""
    from ptmcmcsampler_b200 import PTSampler

    d = 3
    mu, cov, logl_grad, logp_grad = target(d)

    def logl(x):
        return logl_grad(x)[0]

    def logp(x):
        return 0.0

    s = PTSampler(d, logl, logp, 0.5 * np.eye(d), logl_grad=logl_grad, logp_grad=logp_grad, outDir=str(tmp_path / "g"),
                  verbose=False, seed=3)
    s.sample(mu + 0.5, 60000, burn=500, covUpdate=500, thin=1, isave=1000, SCAMweight=10, AMweight=10, DEweight=10,
             NUTSweight=5, HMCweight=10, MALAweight=5, HMCstepsize=0.2, HMCsteps=10, neff=800)
    assert {"HMCJump", "NUTSJump", "MALAJump"} <= set(s.jumpDict)
    assert all(s.jumpDict[k][0] > 0 for k in ("HMCJump", "NUTSJump", "MALAJump"))
    assert s.jumpDict["HMCJump"][1] / s.jumpDict["HMCJump"][0] > 0.5
    it = s._engine.iteration
    assert it < 60000 and it % 1000 == 0 and s._last_neff >= 800
    x = s._chain[500:it]
    assert np.all(np.abs(x.mean(0) - mu) < 0.3) and np.allclose(np.cov(x.T), cov, atol=0.5 * np.abs(cov).max())
