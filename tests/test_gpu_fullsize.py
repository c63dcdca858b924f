"""GPU parity at the BASELINE configurations' full sizes (C2, C3, C4) against the OpenMP oracle: every jump id, accept
flag, swap map and counter bit-exact, floating-point state to tolerance.  Each case runs long enough that the pooled
covariance update, the DE-history append and DE joining the cycle all happen at full grid size (hundreds of blocks per
rung, several waves, GB-sized AM / DE rings)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

from test_gpu_parity import gaussian_target, make_pair

pytestmark = pytest.mark.gpu


def compare_full(o, g, x0, niter, tskip, T, ftol):
    nsw = niter // tskip if T > 1 else 0
    otrace, oswap = o.set_trace(niter, nsw)
    o.set_state(x0)
    g.set_state(x0)
    o.run(niter)
    g.run(niter)
    gtrace, gswap = g.trace(niter, nsw)
    njump = int(np.count_nonzero((gtrace & 0x7F) != (otrace & 0x7F)))
    nacc = int(np.count_nonzero((gtrace >> 7) != (otrace >> 7)))
    assert njump == 0, "%d of %d jump ids differ" % (njump, gtrace.size)
    assert nacc == 0, "%d of %d accept flags differ" % (nacc, gtrace.size)
    del gtrace, otrace
    if nsw:
        assert np.array_equal(gswap, oswap[:nsw]), "swap maps differ"
    for a, b in zip(o.counters(), g.counters()):
        assert np.array_equal(a, b)
    for a, b in zip(o.state(), g.state()):
        assert np.allclose(a, b, rtol=ftol, atol=ftol, equal_nan=True)
    for a, b in zip(o.chain(), g.chain()):
        assert a.shape == b.shape and np.allclose(a, b, rtol=ftol, atol=ftol, equal_nan=True)
    oc, omu, om2, onn = o.adapt()
    gc, gmu, gm2, gnn = g.adapt()
    assert onn == gnn and onn > 0
    assert np.allclose(oc, gc, rtol=1e-7, atol=1e-10 * np.abs(oc).max()) and np.allclose(omu, gmu, rtol=1e-7, atol=1e-10)
    oU, oS = o.factor()
    gU, gS = g.factor()
    assert np.allclose(oS, gS, rtol=1e-7, atol=1e-12)


def c2_target(d=20):
    rng = np.random.default_rng(20)
    A = 0.5 - rng.random(d * d).reshape(d, d)
    A = np.triu(A)
    A += A.T - np.diag(A.diagonal())
    cov = A @ A + 0.1 * np.eye(d)
    return (orc.LOGL_GAUSSIAN, orc.gaussian_params(5.0 * np.ones(d), np.linalg.inv(cov)), orc.LOGP_UNIFORM,
            orc.uniform_params(-50 * np.ones(d), 60 * np.ones(d)))


def test_c2_full_size_matches_oracle():
    """BASELINE config 2 as benchmarked: 8192 walkers x 32 temperatures, covUpdate = burn = 1000, Tskip = 100, thin = 10,
    1100 iterations (pooled covariance update over 8.2 M samples, DE append of a 1.3 GB ring, DE in the cycle for the last 100)."""
    d, W, T, niter = 20, 8192, 32, 1100
    o, g = make_pair(d, W, T, 0.01 * np.eye(d), seed=42, target=c2_target(d), cov_update=1000, burn=1000, tskip=100, thin=10,
                     niter=niter, record_hot=False, nthreads=os.cpu_count() or 4)
    assert "mh_sorted_kernel" in g.mh_kernel_name
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    compare_full(o, g, x0, niter, 100, T, 1e-9)


def test_c3_full_size_matches_oracle():
    """BASELINE config 3: 100-dim dense Gaussian, 4096 walkers x 64 temperatures on the tensor-core kernel; covUpdate = burn =
    100 so that 220 iterations cover two covariance updates (100 x 100 Jacobi), the DE append and DE steps."""
    d, W, T, niter = 100, 4096, 64, 220
    s = np.logspace(-1, 1, d)
    idx = np.arange(d)
    cov = 0.9 ** np.abs(idx[:, None] - idx[None, :]) * s[:, None] * s[None, :]
    tgt = (orc.LOGL_GAUSSIAN, orc.gaussian_params(np.zeros(d), np.linalg.inv(cov)), orc.LOGP_UNIFORM,
           orc.uniform_params(-500 * np.ones(d), 500 * np.ones(d)))
    ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(T), 1e30)
    o, g = make_pair(d, W, T, np.diag(0.01 * s * s), seed=7, target=tgt, cov_update=100, burn=100, tskip=50, thin=10,
                     niter=niter, record_hot=False, ladder=ladder, nthreads=os.cpu_count() or 4)
    assert "mh_mma_kernel" in g.mh_kernel_name
    x0 = np.random.default_rng(1).standard_normal((T, W, d)) * s
    compare_full(o, g, x0, niter, 50, T, 1e-8)


def test_c4_full_size_matches_oracle():
    """BASELINE config 4: curved 10-dim density, 16384 walkers x 128 temperatures, SCAM/AM/DE = 10/10/60."""
    d, W, T, niter = 10, 16384, 128, 220
    tgt = (orc.LOGL_CURVED, None, orc.LOGP_UNIFORM, orc.uniform_params(-10 * np.ones(d), 10 * np.ones(d), 0.0, False))
    ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(T), 1e30)
    o, g = make_pair(d, W, T, 0.1 * np.eye(d), seed=9, target=tgt, cycle=((0, 10), (1, 10)), de_weight=60, cov_update=100,
                     burn=100, tskip=50, thin=10, niter=niter, record_hot=False, ladder=ladder, nthreads=os.cpu_count() or 4)
    assert "mh_sorted_kernel" in g.mh_kernel_name
    x0 = np.random.default_rng(2).uniform(-1, 1, (T, W, d))
    compare_full(o, g, x0, niter, 50, T, 1e-8)
