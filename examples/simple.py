#!/usr/bin/env python
"""The reference's examples/simple.py (20-dim Gaussian in a [0, 10] box, AM + SCAM + DE) on the GPU engine.

Three ways to run the same problem, from "change one import" to "everything on the device":

    python examples/simple.py callables     # plain Python logl / logp, one chain, as in the reference
    python examples/simple.py device        # device target descriptors, one chain x one temperature
    python examples/simple.py ensemble      # 1024 walkers x 8 temperatures, custom Python jump vectorised
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from ptmcmcsampler_b200 import PTMCMCSampler as ptmcmc  # reference: from PTMCMCSampler import PTMCMCSampler as ptmcmc
from ptmcmcsampler_b200.likelihoods import GaussianLikelihood, UniformPrior

ndim, pmin, pmax = 20, 0.0, 10.0
rng = np.random.default_rng(0)
mu = rng.uniform(pmin, pmax, ndim)
A = 0.5 - rng.random((ndim, ndim))
A = np.triu(A)
A += A.T - np.diag(A.diagonal())
cov = A @ A
icov = np.linalg.inv(cov)


def lnlikefn(x):          # ref examples/simple.py:34-36
    diff = x - mu
    return -np.dot(diff, np.dot(icov, diff)) / 2.0


def lnpriorfn(x):         # ref examples/simple.py:38-44
    return 0.0 if np.all(pmin <= x) and np.all(pmax >= x) else -np.inf


class UniformJump(object):  # ref tests/test_simple.py:44-62, one call for every chain
    vectorized = True
    __name__ = "uniform_jump"

    def __call__(self, X, it, beta):
        return rng.uniform(pmin, pmax, X.shape), np.zeros(len(X))


mode = sys.argv[1] if len(sys.argv) > 1 else "device"
p0 = rng.uniform(pmin, pmax, ndim)
cov0 = np.eye(ndim) * 0.1**2
kw = dict(burn=500, thin=1, covUpdate=500, SCAMweight=20, AMweight=20, DEweight=20)
if mode == "callables":
    sampler = ptmcmc.PTSampler(ndim, lnlikefn, lnpriorfn, np.copy(cov0), outDir="./chains_callables")
    sampler.sample(p0, 5000, **kw)
elif mode == "device":
    sampler = ptmcmc.PTSampler(ndim, GaussianLikelihood(mu, icov=icov), UniformPrior(pmin, pmax), np.copy(cov0),
                               outDir="./chains_device")
    sampler.sample(p0, 100000, **kw)
else:
    sampler = ptmcmc.PTSampler(ndim, lambda X: -0.5 * np.einsum("ni,ij,nj->n", X - mu, icov, X - mu),
                               lambda X: np.where(np.all((X >= pmin) & (X <= pmax), axis=1), 0.0, -np.inf),
                               np.copy(cov0), outDir="./chains_ensemble", ntemps=8, nwalkers=1024, vectorized=True)
    sampler.addProposalToCycle(UniformJump(), 5)
    sampler.sample(p0, 2000, isave=500, **kw)
chain = sampler._chain_all
print("\nrecorded", chain.shape, "posterior mean (first 4):", chain[len(chain) // 4:].reshape(-1, ndim).mean(0)[:4])
print("true mean (first 4):           ", mu[:4], " acceptance", sampler.naccepted / sampler._engine.iteration)
