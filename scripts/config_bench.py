"""Throughput of the BASELINE.json configurations other than the bench.py workload (C2), one GPU:

    python scripts/config_bench.py c2|c3|c4 [iters] [reps] [W] [T]

C3: 100-dim correlated Gaussian (dense cov 0.9^|i-j| s_i s_j, s log-spaced 0.1..10), 4096 walkers x 64 temps.
C4: 10-dim = five copies of the reference's 2-D curved density, 16384 walkers x 128 temps, SCAM/AM/DE 10/10/60.
Prints chain-steps/s per repetition, acceptance rates and the per-class device times.
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ptmcmcsampler_b200 import _cabi  # noqa: E402


def config(name):
    if name == "c2":
        d, W, T = 20, 8192, 32
        rng = np.random.default_rng(20)
        A = 0.5 - rng.random((d, d))
        A = np.triu(A)
        A += A.T - np.diag(A.diagonal())
        cov = A @ A + 0.1 * np.eye(d)
        kw = dict(logl_kind=_cabi.LOGL_GAUSSIAN, logl_params=np.concatenate([5.0 * np.ones(d), np.linalg.inv(cov).ravel(), [0.0]]),
                  logp_params=np.concatenate([-50 * np.ones(d), 60 * np.ones(d), [0.0, 1.0]]), cycle=((0, 20), (1, 20)),
                  de_weight=20)
        x0 = lambda rng: rng.uniform(0, 10, (T, W, d))  # noqa: E731
        return d, W, T, kw, x0, np.diag(cov), 0.01 * np.eye(d)
    if name == "c3":
        d, W, T = 100, 4096, 64
        s = np.logspace(-1, 1, d)
        idx = np.arange(d)
        cov = 0.9 ** np.abs(idx[:, None] - idx[None, :]) * s[:, None] * s[None, :]
        kw = dict(logl_kind=_cabi.LOGL_GAUSSIAN, logl_params=np.concatenate([np.zeros(d), np.linalg.inv(cov).ravel(), [0.0]]),
                  logp_params=np.concatenate([-500 * np.ones(d), 500 * np.ones(d), [0.0, 1.0]]), cycle=((0, 20), (1, 20)),
                  de_weight=20)
        x0 = lambda rng: rng.standard_normal((T, W, d)) * s  # noqa: E731
        return d, W, T, kw, x0, np.diag(cov), np.diag(0.01 * s * s)
    if name == "c4":
        d, W, T = 10, 16384, 128
        kw = dict(logl_kind=_cabi.LOGL_CURVED, logl_params=None,
                  logp_params=np.concatenate([-10 * np.ones(d), 10 * np.ones(d), [0.0, 0.0]]), cycle=((0, 10), (1, 10)),
                  de_weight=60)
        x0 = lambda rng: rng.uniform(-1, 1, (T, W, d))  # noqa: E731
        return d, W, T, kw, x0, None, np.eye(d) * 0.1
    raise SystemExit("unknown config " + name)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    d, W, T, kw, x0, var, cov0 = config(name)
    if len(sys.argv) > 4:
        W = int(sys.argv[4])
    if len(sys.argv) > 5:
        T = int(sys.argv[5])
    d, W0, T0, kw, x0, var, cov0 = config(name)
    ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(T), 1e30)
    e = _cabi.Engine(d, W, T, cov0, ladder, seed=1, cov_update=1000, burn=1000, tskip=100, thin=10,
                     record_rows=(1100 + iters * reps) // 10 + 2, timing=False, **kw)
    full = x0(np.random.default_rng(1))
    e.set_state(np.ascontiguousarray(full[:T, :W]) if full.shape[0] >= T else np.resize(full, (T, W, d)))
    e.run(1000 + 100)   # covariance adapted once, DE in the cycle
    e.sync()
    e.set_timing(True)
    e.reset_timing()
    for r in range(reps):
        t0 = time.time()
        e.run(iters)
        e.sync()
        dt = time.time() - t0
        print("%s rep %d: %.3f s  %.3e chain-steps/s (d=%d W=%d T=%d)" % (name, r, dt, W * T * iters / dt, d, W, T), flush=True)
    tm = e.timing()
    print("class ms:", {k: round(v, 2) for k, v in tm["ms"].items() if v}, "launches", {k: v for k, v in tm["launches"].items() if v})
    prop, acc, sw, nsw = e.counters()
    print("acceptance per jump (cold):", acc[0].sum(0) / np.maximum(1, prop[0].sum(0)))
    print("swap acc (first rungs):", sw[:4].mean(1) / max(1, nsw))
    x = e.state()[0]
    if var is not None:
        print("cold var/target (first 4, last):", (x[0].var(0) / var)[[0, 1, 2, 3, -1]])


if __name__ == "__main__":
    main()
