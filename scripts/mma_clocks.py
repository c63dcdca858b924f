"""Per-phase clocks of the tensor-core MH kernel (needs a build with PTMCMC_NVCC_EXTRA=-DPTMCMC_MMA_CLOCKS):

    PTMCMC_NVCC_EXTRA=-DPTMCMC_MMA_CLOCKS python -m ptmcmcsampler_b200.build --force
    python scripts/mma_clocks.py [iters]
"""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from ptmcmcsampler_b200 import _cabi  # noqa: E402
from config_bench import config  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
d, W, T, kw, x0, var, cov0 = config("c3")
ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(T), 1e30)
eng = _cabi.Engine(d, W, T, cov0, ladder, seed=1, cov_update=1000, burn=1000, tskip=100, thin=10,
                   record_rows=(1100 + iters) // 10 + 2, timing=False, **kw)
eng.set_state(x0(np.random.default_rng(1)))
eng.run(1100)
eng.sync()
fn = _cabi._lib.ptmcmc_debug_mma_clocks
fn.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
out = (ctypes.c_uint64 * 32)()
fn(out, 1)
eng.run(iters - 1100 if iters > 1100 else 100)
eng.sync()
fn(out, 0)
v = np.array(list(out), dtype=float)
names = ["R (gathers, AM normals)", "P (AM mat-vec by DMMA)", "L (quadratic form by DMMA)",
         "epilogue (Hastings, update; draws of the next iteration)", "-"]
if "split" not in eng.mh_kernel_name:
    names = ["A (bookkeeping, jump pick, scalar draws, lists)", "R (gathers + AM normals)", "P mma", "P write-back", "L"]
nc = int(eng.mh_kernel_name.split("(")[1].split()[0])
n = (iters - 1100 if iters > 1100 else 100) * ((T * W + nc - 1) // nc)
print("kernel", eng.mh_kernel_name)
for k, nm in enumerate(names):
    print("%-52s %8.0f clk / block-iteration  %5.1f %%" % (nm, v[k] / n, 100 * v[k] / v[:5].sum()))
print("total %.0f" % (v[:5].sum() / n))
if "split" in eng.mh_kernel_name:
    sub = {8: "R thread 0, first task: gather loads issued", 9: "R thread 0: AM task", 10: "R thread 0: loads consumed, stored",
           24: "R thread 0: nothing (two clock reads)", 11: "R thread 0: bookkeeping", 5: "R thread 0: setup", 12: "P thread 0: mat-vec before the barrier", 13: "L thread 0: quadratic form before the barrier",
           14: "epilogue warp 0: Hastings test, update", 15: "epilogue warp 7: jump kinds + lists (+ named barrier)",
           7: "epilogue warp 7: DE draws", 6: "epilogue warp 6: wait for the lists + SCAM draws"}
    for k, nm in sub.items():
        print("  %-58s %8.0f clk" % (nm, v[k] / n))
    print("  R: arrival of warps 0..7 at the closing barrier:", " ".join("%.0f" % (x / n) for x in v[16:24]))
