import sys, os
import numpy as np
sys.path.insert(0, ".")
from ptmcmcsampler_b200 import _cabi
d, W, T = 100, 256, 4
s = np.logspace(-1, 1, d)
idx = np.arange(d)
cov = 0.9 ** np.abs(idx[:, None] - idx[None, :]) * s[:, None] * s[None, :]
ladder = (1 + np.sqrt(2.0 / d)) ** np.arange(T)
for chunk in (1000, 20000):
    e = _cabi.Engine(d, W, T, np.diag(0.01 * s * s), ladder, seed=3, cov_update=1000, burn=1000, tskip=100, thin=10,
                     logl_params=np.concatenate([np.zeros(d), np.linalg.inv(cov).ravel(), [0.0]]),
                     logp_params=np.concatenate([-500 * np.ones(d), 500 * np.ones(d), [0.0, 1.0]]),
                     record_rows=2002, record_hot=True)
    print(e.mh_kernel_name, "chunk", chunk, flush=True)
    e.set_state(np.random.default_rng(1).standard_normal((T, W, d)) * s)
    for k in range(20000 // chunk):
        e.run(chunk)
        x, lnl, lp = e.state()[:3]
        prop, acc, sw, nsw = e.counters()
        print(e.iteration, "max|x|/s", np.abs(x / s).max(), "lnl min", lnl.min(), "acc", (acc.sum((0, 1)) / np.maximum(1, prop.sum((0, 1)))).round(3), flush=True)
    ch = e.chain()[0]
    print("chain", ch.shape, "row std of last rows (col 0, 50, 99):", ch[-500:, 0].reshape(-1, d).std(0)[[0, 50, 99]], "target", np.sqrt(np.diag(cov))[[0, 50, 99]])
    print("first bad row:", next((i for i in range(len(ch)) if np.abs(ch[i] / s).max() > 50), None))
