"""Quick throughput probe of the MH path on one GPU (development aid, not the contract bench)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ptmcmcsampler_b200 import _cabi  # noqa: E402


def main(d=20, W=8192, T=32, niter=1000, reps=3):
    rng = np.random.default_rng(20)
    A = 0.5 - rng.random((d, d))
    A = np.triu(A)
    A += A.T - np.diag(A.diagonal())
    cov = A @ A + 0.1 * np.eye(d)
    mu = 5.0 * np.ones(d)
    lpar = np.concatenate([mu, np.linalg.inv(cov).ravel(), [0.0]])
    ppar = np.concatenate([-50 * np.ones(d), 60 * np.ones(d), [0.0, 1.0]])
    ladder = (1 + np.sqrt(2.0 / d)) ** np.arange(T)
    e = _cabi.Engine(d, W, T, 0.01 * np.eye(d), ladder, seed=1, cov_update=1000, burn=1000, tskip=100, thin=10,
                     logl_params=lpar, logp_params=ppar, record_rows=niter * (reps + 2) // 10 + 2, timing=False)
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    e.set_state(x0)
    print("kernel:", e.mh_kernel_name, flush=True)
    e.run(niter)
    e.sync()
    for r in range(reps):
        t0 = time.time()
        e.run(niter)
        e.sync()
        dt = time.time() - t0
        print("rep %d: %.3f s  %.3e chain-steps/s" % (r, dt, W * T * niter / dt), flush=True)
    prop, acc, sw, nsw = e.counters()
    print("acceptance per jump (cold):", acc[0].sum(0) / np.maximum(1, prop[0].sum(0)))
    print("swap acc (first rungs):", sw[:4].mean(1) / max(1, nsw))
    e.set_timing(True)
    e.reset_timing()
    e.run(niter)
    e.sync()
    tm = e.timing()
    print("mh kernel: %.3f ms / launch over %d launches -> %.3e chain-steps/s in-kernel; class ms %s" % (
        tm["ms"]["mh"] / tm["launches"]["mh"], tm["launches"]["mh"], W * T * niter / (tm["ms"]["mh"] * 1e-3),
        {k: round(v, 2) for k, v in tm["ms"].items() if v}))
    x, lnl, lp, lnp = e.state()
    print("cold mean", x[0].mean(0)[:4], "cold var/target", (x[0].var(0) / np.diag(cov))[:4])


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
