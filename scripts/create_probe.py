import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from ptmcmcsampler_b200 import _cabi
D, W, T = bench.D, bench.W, bench.T
mu, cov, ladder = bench.problem()
lpar = np.concatenate([mu, np.linalg.inv(cov).ravel(), [0.0]])
ppar = np.concatenate([-50 * np.ones(D), 60 * np.ones(D), [0.0, 1.0]])
x0 = np.random.default_rng(1).uniform(0, 10, (T, W, D))
for i in range(12):
    t0 = time.perf_counter()
    e = _cabi.Engine(D, W, T, 0.01 * np.eye(D), ladder, seed=42, cov_update=1000, burn=1000, tskip=100, thin=10,
                     logl_params=lpar, logp_params=ppar, record_rows=102)
    t1 = time.perf_counter()
    e.set_state(x0)
    e.sync()
    t2 = time.perf_counter()
    if i % 2:
        e.run(200); e.sync()
    t3 = time.perf_counter()
    e.close()
    t4 = time.perf_counter()
    print("create %.1f ms  set_state %.1f ms  run %.1f ms  close %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
