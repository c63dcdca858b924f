"""Exploration for the d=100 known-answer test: cold-rung variance / target over time (engine only)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from ptmcmcsampler_b200 import _cabi  # noqa: E402

d, W, T = 100, 256, 4
s = np.logspace(-1, 1, d)
idx = np.arange(d)
cov = 0.9 ** np.abs(idx[:, None] - idx[None, :]) * s[:, None] * s[None, :]
ladder = (1 + np.sqrt(2.0 / d)) ** np.arange(T)
e = _cabi.Engine(d, W, T, np.diag(0.01 * s * s), ladder, seed=3, cov_update=1000, burn=1000, tskip=100, thin=10,
                 logl_params=np.concatenate([np.zeros(d), np.linalg.inv(cov).ravel(), [0.0]]),
                 logp_params=np.concatenate([-500 * np.ones(d), 500 * np.ones(d), [0.0, 1.0]]), record_rows=4096,
                 record_hot=True)
print(e.mh_kernel_name)
e.set_state(np.random.default_rng(1).standard_normal((T, W, d)) * s)
L = np.linalg.cholesky(np.linalg.inv(cov))
done = 0
for target in (2000, 5000, 10000, 20000, 40000, 60000):
    e.run(target - done)
    done = target
    ch = e.chain()[0]  # rows since the last release
    e.release_rows(e.rows)
    half = ch[len(ch) // 2:]
    out = []
    for t in range(T):
        x = half[:, t].reshape(-1, d)
        w = x @ L  # whitened: unit covariance at T=1, T at rung t
        out.append((x.var(0) / np.diag(cov)).mean() / ladder[t])
        out.append(np.abs(np.cov(w.T) / ladder[t] - np.eye(d)).max())
    prop, acc, sw, nsw = e.counters()
    print(target, "var/(T*target) and max|whitened cov/T - I| per rung:", np.round(out, 3), "acc(cold)", np.round(acc[0].sum(0) / prop[0].sum(0), 3))
