"""Times the UNMODIFIED reference (nanograv/PTMCMCSampler, installed by scripts/install_reference.sh into the
git-ignored baseline/_ref/) on host cores.  Used only by `bench.py --impl reference`; nothing here touches the
engine.  Two workloads:

* ``config1``: examples/simple.py verbatim settings (ref examples/simple.py:52-122; SURVEY 8d "C1"): 20-dim
  Gaussian in a [0, 10] box, one chain, Niter = 10 000, burn = covUpdate = 500, thin = 1, SCAM/AM/DE 20/20/20.
* ``c2``: the bench target (C2: untruncated 20-dim Gaussian, burn = covUpdate = 1000, thin = 10, isave = Niter so
  that file output stays out of the timing), one chain per process at one rung of the C2 ladder.  mpi4py is not
  in the image, so the reference's PTswap (one MPI rank per temperature) cannot run: the processes are
  independent chains and the figure is the reference's MH-step rate on all cores.
"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "PTMCMCSampler", "PTMCMCSampler.py"))


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from PTMCMCSampler import PTMCMCSampler as ptmcmc  # the reference package, not ptmcmcsampler_b200

    assert os.path.abspath(ptmcmc.__file__).startswith(REF)
    return ptmcmc


class _Gaussian(object):  # the fixture of ref examples/simple.py:12-44
    def __init__(self, mu, cov, pmin, pmax):
        self.mu, self.icov, self.a, self.b = mu, np.linalg.inv(cov), pmin, pmax

    def lnlikefn(self, x):
        diff = x - self.mu
        return -np.dot(diff, np.dot(self.icov, diff)) / 2.0

    def lnpriorfn(self, x):
        if np.all(self.a <= x) and np.all(self.b >= x):
            return 0.0
        return -np.inf


def run_config1(seed=0):
    """examples/simple.py:52-122 as written (np.random.seed problem, cov0 = 0.01 I)."""
    ptmcmc = _import_reference()
    np.random.seed(seed)
    ndim = 20
    pmin, pmax = 0.0, 10.0
    means = np.random.rand(ndim) * (pmax - pmin) + pmin
    cov = 0.5 - np.random.rand(ndim**2).reshape((ndim, ndim))
    cov = np.triu(cov)
    cov += cov.T - np.diag(cov.diagonal())
    cov = np.dot(cov, cov)
    g = _Gaussian(means, cov, pmin * np.ones(ndim), pmax * np.ones(ndim))
    p0 = np.random.uniform(pmin, pmax, ndim)
    out = tempfile.mkdtemp(prefix="ref_c1_")
    niter = 10000
    try:
        s = ptmcmc.PTSampler(ndim, g.lnlikefn, g.lnpriorfn, np.diag(np.ones(ndim) * 0.01), outDir=out, verbose=False)
        t0 = time.perf_counter()
        s.sample(p0, niter, burn=500, thin=1, covUpdate=500, SCAMweight=20, AMweight=20, DEweight=20)
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return niter / dt, dt


def run_c2(args):
    """One reference chain on the C2 target at rung `rung` of the C2 ladder for `niter` iterations."""
    rung, niter = args
    ptmcmc = _import_reference()
    d = 20
    rng = np.random.default_rng(20)
    A = 0.5 - rng.random(d * d).reshape(d, d)
    A = np.triu(A)
    A += A.T - np.diag(A.diagonal())
    cov = A @ A + 0.1 * np.eye(d)
    g = _Gaussian(5.0 * np.ones(d), cov, -50.0 * np.ones(d), 60.0 * np.ones(d))
    p0 = np.random.default_rng(1 + rung).uniform(0, 10, d)
    out = tempfile.mkdtemp(prefix="ref_c2_")
    try:
        s = ptmcmc.PTSampler(d, g.lnlikefn, g.lnpriorfn, 0.01 * np.eye(d), outDir=out, verbose=False, seed=42 + rung)
        t0 = time.perf_counter()
        s.sample(p0, niter, burn=1000, covUpdate=1000, thin=10, isave=niter, SCAMweight=20, AMweight=20, DEweight=20,
                 ladder=np.array([float((1 + np.sqrt(2.0 / d)) ** rung)]))
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return niter / dt, dt


def run_c2_all_cores(cores, niter):
    """`cores` independent reference processes; returns (aggregate chain-steps/s, wall seconds)."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(run_c2, [(r % 32, niter) for r in range(cores)])
    wall = time.perf_counter() - t0
    return float(sum(r[0] for r in res)), wall


if __name__ == "__main__":
    print("config1: %.1f chain-steps/s (%.2f s)" % run_config1())
    print("c2 x1: %.1f chain-steps/s (%.2f s)" % run_c2((0, 20000)))
