"""Where does the end-to-end time of PTSampler.sample() go (C2, 1000 iterations)?"""
import cProfile
import pstats
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from ptmcmcsampler_b200 import PTMCMCSampler, _cabi  # noqa: E402
from ptmcmcsampler_b200.likelihoods import GaussianLikelihood, UniformPrior  # noqa: E402

D, W, T = bench.D, bench.W, bench.T
mu, cov, ladder = bench.problem()
lk, pr = GaussianLikelihood(mu, icov=np.linalg.inv(cov)), UniformPrior(-50.0, 60.0)
p0 = _cabi.pinned_empty((T, W, D))
p0[...] = np.random.default_rng(7).uniform(0, 10, (T, W, D))
outdir = tempfile.mkdtemp(prefix="ptmcmc_probe_")


def step(seed):
    s = PTMCMCSampler.PTSampler(D, lk, pr, 0.01 * np.eye(D), outDir=outdir, verbose=False, seed=seed, ntemps=T, nwalkers=W)
    s.sample(p0, 1000, burn=1000, covUpdate=1000, Tskip=100, thin=10, isave=1000, SCAMweight=20, AMweight=20, DEweight=20)
    v = float(s._lnlike_all[-1].mean())
    s.engine.close()
    return v


step(1)
for i in range(6):
    pr_ = cProfile.Profile()
    t0 = time.perf_counter()
    pr_.enable()
    step(2 + i)
    pr_.disable()
    print("e2e step %.1f ms" % (1e3 * (time.perf_counter() - t0)))
    st = pstats.Stats(pr_)
    rows = sorted(st.stats.items(), key=lambda kv: -kv[1][2])[:7]   # by tottime
    print("   " + "; ".join("%s:%s %.0fms" % (k[0].split("/")[-1], k[2], 1e3 * v[2]) for k, v in rows))
