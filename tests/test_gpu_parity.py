"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.

Integer bookkeeping (jump ids, accept flags, swap maps, counters) must be bit exact; floating point
state must agree to FTOL (the two sides use different libm / FMA contraction and a parallel vs
sequential covariance reduction).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from ptmcmcsampler_b200 import _cabi

from _helpers import fixture_cycle, fixture_groups, fixture_mh_temp, fixture_target, load

pytestmark = pytest.mark.gpu
FTOL = 1e-9


def make_pair(d, W, T, cov0, seed, target, groups=None, cycle=((0, 20), (1, 20)), de_weight=20, cov_update=50,
              burn=100, tskip=10, thin=5, niter=300, ladder=None, record_hot=True, mh_temp=None, variant=None,
              sort_cfg=None, nthreads=4):
    lk, lpar, pk, ppar = target
    ladder = orc.temperature_ladder(d, T) if ladder is None else np.asarray(ladder, float)
    rows = niter // thin + 1
    o = orc.Oracle(d, W, T, cov0, seed=seed, ladder=ladder, mh_temp=mh_temp, groups=groups, cycle=cycle,
                   de_weight=de_weight, cov_update=cov_update, burn=burn, tskip=tskip, thin=thin, logl_kind=lk,
                   logl_params=lpar, logp_kind=pk, logp_params=ppar, record_hot=record_hot, max_rows=rows, nthreads=nthreads)
    env = {"PTMCMC_MH_VARIANT": variant, "PTMCMC_SORT_CFG": sort_cfg}  # read when the engine is created
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        if v is not None:
            os.environ[k] = str(v)
    try:
        g = _cabi.Engine(d, W, T, cov0, ladder, mh_temp=mh_temp, seed=seed, groups=groups, cycle=cycle,
                         de_weight=de_weight, cov_update=cov_update, burn=burn, tskip=tskip, thin=thin, logl_kind=lk,
                         logl_params=lpar, logp_kind=pk, logp_params=ppar, record_hot=record_hot, record_rows=rows,
                         trace_iters=niter)
    finally:
        for k, v in env.items():
            if v is not None:
                if old[k] is None:
                    del os.environ[k]
                else:
                    os.environ[k] = old[k]
    return o, g


def compare(o, g, x0, niter, tskip, T, chunks=(1.0,), ftol=FTOL):
    otrace, oswap = o.set_trace(niter, niter // tskip if T > 1 else 0)
    o.set_state(x0)
    g.set_state(x0)
    done = 0
    for frac in chunks:
        n = int(round(niter * frac)) - done
        o.run(n)
        g.run(n)
        done += n
    assert g.iteration == o.iteration == niter
    nsw = niter // tskip if T > 1 else 0
    gtrace, gswap = g.trace(niter, nsw)
    # --- integers: bit exact
    assert np.array_equal(gtrace & 0x7F, otrace & 0x7F), "jump ids differ"
    assert np.array_equal(gtrace >> 7, otrace >> 7), "accept flags differ"
    if nsw:
        assert np.array_equal(gswap, oswap[:nsw]), "swap maps differ"
    op, oa, osw, on = o.counters()
    gp, ga, gsw, gn = g.counters()
    assert np.array_equal(op, gp) and np.array_equal(oa, ga)
    assert np.array_equal(osw, gsw) and on == gn
    # --- floats
    for a, b in zip(o.state(), g.state()):
        assert np.allclose(a, b, rtol=ftol, atol=ftol, equal_nan=True)
    for a, b in zip(o.chain(), g.chain()):
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=ftol, atol=ftol, equal_nan=True)
    oc, omu, om2, onn = o.adapt()
    gc, gmu, gm2, gnn = g.adapt()
    assert onn == gnn
    assert np.allclose(oc, gc, rtol=1e-8, atol=1e-11 * np.abs(oc).max()) and np.allclose(omu, gmu, rtol=1e-8, atol=1e-12)
    assert np.allclose(om2, gm2, rtol=1e-8, atol=max(1e-9, 1e-11 * np.abs(om2).max()))
    oU, oS = o.factor()
    gU, gS = g.factor()
    assert np.allclose(oS, gS, rtol=1e-8, atol=1e-14) and np.allclose(oU, gU, rtol=0, atol=1e-7)
    for a, b in zip(o.buffers(), g.buffers()):
        assert np.allclose(a, b, rtol=ftol, atol=ftol)


def gaussian_target(d, seed, lo=-50.0, hi=60.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    mu = rng.uniform(2, 8, d)
    return (orc.LOGL_GAUSSIAN, orc.gaussian_params(mu, np.linalg.inv(cov)), orc.LOGP_UNIFORM,
            orc.uniform_params(lo * np.ones(d), hi * np.ones(d)))


def test_device_normals_equal_oracle_bits():
    """word_to_normals on the device (round-toward-zero conversion + fma, integer split of the float, sign-bit
    quadrant logic) returns the oracle's doubles bit for bit: 10^7 random words plus the edges of every branch."""
    rng = np.random.default_rng(5)
    words = rng.integers(0, 2**64, 10_000_000, dtype=np.uint64)
    his = np.array([0, 1, 2, 3, 2**23 - 1, 2**23, 2**23 + 1, 2**24 - 1, 2**24, 2**24 + 1, 2**25 + 3, 2**31, 2**32 - 129,
                    2**32 - 128, 2**32 - 1, 0xB504F300, 0xB504F3FF, 0xB504F400, 0x5A827980, 0x5A827A00], dtype=np.uint64)
    los = np.array([0, 1, 63, 64, 2**29, 2**29 + 64, 2**30 - 1, 2**30, 2**31 - 64, 2**31, 2**31 + 2**29, 3 * 2**30,
                    2**32 - 1, 0x20000000, 0x20000040, 0x1FFFFFC0], dtype=np.uint64)
    edge = ((his[:, None] << np.uint64(32)) | los[None, :]).ravel()
    words = np.concatenate([words, edge])
    d0, d1 = _cabi.device_normals(words)
    o0, o1 = orc.word_to_normals_many(words)
    assert np.array_equal(d0.view(np.uint64), o0.view(np.uint64))
    assert np.array_equal(d1.view(np.uint64), o1.view(np.uint64))


@pytest.mark.parametrize("sort_cfg", [None, 0])
@pytest.mark.parametrize("d,W,T", [(20, 64, 4), (5, 33, 3), (8, 128, 1), (12, 7, 5), (32, 16, 2), (3, 1, 1), (1, 5, 2),
                                   (2, 9, 1), (20, 300, 2), (18, 33, 3), (17, 5, 1), (16, 200, 2), (24, 150, 2), (10, 400, 1)])
def test_sorted_kernel_matches_oracle(d, W, T, sort_cfg):
    """mh_sorted_kernel, default geometry (AM chains on lane pairs) and one task per chain (cfg 0): same draws,
    same decisions; widths with odd and even numbers of normal pairs per lane, padded and unpadded ndim."""
    niter, tskip = 320, 10
    cov0 = np.diag(0.01 * (1.0 + np.arange(d)))
    o, g = make_pair(d, W, T, cov0, seed=11 + d, target=gaussian_target(d, d), niter=niter, tskip=tskip, sort_cfg=sort_cfg,
                     variant=6)
    assert "mh_sorted_kernel" in g.mh_kernel_name
    x0 = np.random.default_rng(d).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, tskip, T, chunks=(0.33, 0.7, 1.0))


@pytest.mark.parametrize("cycle,dew", [(((0, 5), (1, 40)), 5), (((1, 20),), 0), (((0, 20),), 0)])
def test_sorted_kernel_am_heavy_cycle_overflows_task_list(cycle, dew):
    """More AM chains in a block than its spare threads cover: the task loop takes a second pass."""
    d, W, T, niter, tskip = 20, 256, 2, 240, 10
    cov0 = np.diag(0.01 * (1.0 + np.arange(d)))
    o, g = make_pair(d, W, T, cov0, seed=3, target=gaussian_target(d, d), niter=niter, tskip=tskip, cycle=cycle, de_weight=dew)
    x0 = np.random.default_rng(d).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, tskip, T, chunks=(0.5, 1.0))


@pytest.mark.parametrize("d,W,T", [(20, 64, 4), (5, 33, 3), (8, 128, 1), (12, 7, 5), (32, 16, 2), (3, 1, 1),
                                   (24, 300, 2), (48, 40, 3), (100, 24, 2), (128, 9, 2), (1, 5, 2), (2, 9, 1)])
def test_tensor_core_kernel_matches_oracle(d, W, T):
    """mh_mma_kernel (fp64 DMMA): same draws, same decisions; the quadratic form differs in summation order."""
    niter, tskip = 320, 10
    cov0 = np.diag(0.01 * (1.0 + np.arange(d)))
    o, g = make_pair(d, W, T, cov0, seed=11 + d, target=gaussian_target(d, d), niter=niter, tskip=tskip, variant=3)
    x0 = np.random.default_rng(d).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, tskip, T, chunks=(0.33, 0.7, 1.0), ftol=1e-8 if d > 32 else FTOL)


@pytest.mark.parametrize("d,W,T", [(100, 24, 2), (40, 50, 1), (128, 9, 2)])
def test_tensor_core_split_kernel_symmetric_form_and_trace(d, W, T, monkeypatch):
    """ndim > 32 (mh_mma_split_kernel) with the form kept symmetric (what a matrix that is not positive definite gets)
    instead of the packed Cholesky factor, the jump trace on, and the one-block-per-SM kernel it replaced (A/B switch)."""
    niter, tskip = 120, 10
    cov0 = np.diag(0.01 * (1.0 + np.arange(d)))
    x0 = np.random.default_rng(d).uniform(0, 10, (T, W, d))
    for env in ({"PTMCMC_MMA_FULL": "1"}, {"PTMCMC_MMA_SPLIT": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        o, g = make_pair(d, W, T, cov0, seed=11 + d, target=gaussian_target(d, d), niter=niter, tskip=tskip, variant=3)
        assert ("mh_mma_split_kernel" in g.mh_kernel_name) == ("PTMCMC_MMA_FULL" in env)
        compare(o, g, x0, niter, tskip, T, chunks=(0.5, 1.0), ftol=1e-8)
        for k in env:
            monkeypatch.delenv(k)


def test_tensor_core_kernel_truncated_box_and_outside_start():
    d, W, T, niter = 6, 40, 3, 250
    tgt = gaussian_target(d, 3, lo=3.0, hi=7.0)
    o, g = make_pair(d, W, T, np.eye(d) * 0.05, seed=5, target=tgt, niter=niter, variant=3)
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T)


def test_sorted_kernel_widest_instance_matches_oracle():
    """ndim in (24, 32] defaults to the tensor-core kernel; variant 6 keeps the sorted kernel's DP=32 instance covered."""
    d, W, T, niter = 30, 140, 2, 200
    o, g = make_pair(d, W, T, np.diag(0.01 * (1.0 + np.arange(d))), seed=41, target=gaussian_target(d, d), niter=niter,
                     variant=6)
    x0 = np.random.default_rng(d).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T)


def test_prior_draw_jump_matches_oracle():
    """The device-side prior-draw jump (cycle id 3) in the generic kernel, W walkers, against the oracle."""
    d, W, T, niter = 7, 50, 3, 260
    tgt = gaussian_target(d, 5, lo=0.0, hi=10.0)
    o, g = make_pair(d, W, T, np.eye(d) * 0.05, seed=8, target=tgt, niter=niter,
                     cycle=((orc.JUMP_PRIOR, 5), (orc.JUMP_SCAM, 20), (orc.JUMP_AM, 20)))
    x0 = np.random.default_rng(6).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T)
    assert g.counters()[0][..., orc.JUMP_PRIOR].sum() > 0


def test_truncated_box_and_outside_start():
    """Box prior tighter than the likelihood (ref examples/simple.py), with walkers that start
    outside the prior: lnprob0 = -inf, first in-prior proposal always accepted (ref :481-483)."""
    d, W, T, niter = 6, 40, 3, 250
    tgt = gaussian_target(d, 3, lo=3.0, hi=7.0)
    o, g = make_pair(d, W, T, np.eye(d) * 0.05, seed=5, target=tgt, niter=niter)
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T)


def test_curved_likelihood_de_dominant():
    d, W, T, niter = 10, 48, 6, 300
    tgt = (orc.LOGL_CURVED, None, orc.LOGP_UNIFORM, orc.uniform_params(-10 * np.ones(d), 10 * np.ones(d), 0.0, False))
    o, g = make_pair(d, W, T, np.eye(d), seed=9, target=tgt, cycle=((0, 10), (1, 10)), de_weight=60, niter=niter,
                     cov_update=40, burn=80, tskip=5)
    x0 = np.random.default_rng(2).uniform(-1, 1, (T, W, d))
    compare(o, g, x0, niter, 5, T, ftol=1e-8)


def test_generic_kernel_groups():
    gfx = load("traj_t4_groups_d6")
    d, T, W, niter = 6, 4, 24, 300
    o, g = make_pair(d, W, T, gfx["cov0"], seed=77, target=fixture_target(gfx), groups=fixture_groups(gfx),
                     cycle=fixture_cycle(gfx), de_weight=50, cov_update=50, burn=200, tskip=7, thin=5, niter=niter,
                     ladder=gfx["ladder"])
    x0 = np.random.default_rng(3).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 7, T)


def test_generic_kernel_large_dim():
    d, W, T, niter = 48, 16, 3, 130
    tgt = gaussian_target(d, 8)
    o, g = make_pair(d, W, T, np.eye(d) * 0.01, seed=21, target=tgt, niter=niter, cov_update=40, burn=80)
    x0 = np.random.default_rng(4).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T, ftol=1e-8)


def test_hot_chain_temperature_override():
    """hotChain=True: the last rung samples at temp=1e80 but swaps with ladder[-1] (ref :281-282, :658)."""
    d, W, T, niter = 5, 32, 3, 200
    ladder = orc.temperature_ladder(d, T)
    mh = ladder.copy()
    mh[-1] = 1e80
    o, g = make_pair(d, W, T, np.eye(d) * 0.02, seed=2, target=gaussian_target(d, 1, lo=0.0, hi=10.0), niter=niter,
                     ladder=ladder, mh_temp=mh)
    x0 = np.random.default_rng(5).uniform(0, 10, (T, W, d))
    compare(o, g, x0, niter, 10, T)


@pytest.mark.parametrize("name", ["traj_t1_d5", "traj_t4_groups_d6", "traj_t1_d20", "traj_t3_hot_tmax_d4", "traj_t2_prior_d4"])
def test_engine_reproduces_reference_trajectory(name):
    """W=1: the engine fed the reference's own eigen-factors walks the reference's trajectory."""
    gfx = load(name)
    d, T, N = int(gfx["d"]), int(gfx["T"]), int(gfx["N"])
    cu, burn, tskip = int(gfx["kw_covUpdate"]), int(gfx["kw_burn"]), int(gfx["kw_Tskip"])
    lk, lpar, pk, ppar = fixture_target(gfx)
    g = _cabi.Engine(d, 1, T, gfx["cov0"], gfx["ladder"], mh_temp=fixture_mh_temp(gfx), seed=int(gfx["seed"]),
                     groups=fixture_groups(gfx),
                     cycle=fixture_cycle(gfx), de_weight=int(gfx["kw_DEweight"]), cov_update=cu, burn=burn,
                     tskip=tskip, thin=1, logl_kind=lk, logl_params=lpar, logp_kind=pk, logp_params=ppar,
                     record_hot=True, record_rows=N + 1, trace_iters=N)
    g.set_state(gfx["p0"][:, None, :])
    done, k = 0, 0
    while done < N:  # stop at every covariance boundary to inject LAPACK's factor (sign ambiguity)
        nxt = min(N, (done // cu + 1) * cu)
        g.run(nxt - done)
        done = nxt
        if done < N and done % cu == 0:
            g.run(0)
            # the update itself happens at the start of the next iteration: run 1 step after injecting
            # is not possible, so apply the engine's update first and then overwrite its factor
            batch = g.adapt_begin()
            g.adapt_finish(batch)
            g.set_factor(gfx["U"][k], gfx["S"][k])
            k += 1
    tr, _ = g.trace(N, 0)
    assert np.array_equal(tr[:, :, 0] & 0x7F, gfx["jump"])
    assert np.array_equal(tr[:, :, 0] >> 7, gfx["acc"])
    ch, lnl, lnp = g.chain()
    assert np.allclose(ch[1:, :, 0], gfx["x"], rtol=1e-9, atol=1e-9)
    assert np.allclose(lnl[1:, :, 0], gfx["lnl"], rtol=1e-9, atol=1e-9)
    prop, acc, sw, nsw = g.counters()
    assert np.array_equal(sw[:, 0], gfx["swap_acc"][-1]) and nsw == int(gfx["swap_proposed"])
    cov, mu, m2, n = g.adapt()
    assert np.allclose(cov, gfx["cov"], rtol=1e-8, atol=1e-12)


def test_engines_on_two_devices_in_one_process():
    """One host thread driving engines on cuda:0 and cuda:1 (every ABI call makes its device current)."""
    if _cabi.load().ptmcmc_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    d, W, T, niter = 20, 64, 3, 120
    tgt = gaussian_target(d, 2)
    lk, lpar, pk, ppar = tgt
    ladder = orc.temperature_ladder(d, T)
    x0 = np.random.default_rng(0).uniform(0, 10, (T, W, d))
    engs = [_cabi.Engine(d, W, T, np.eye(d) * 0.01, ladder, seed=3, cov_update=50, burn=100, tskip=10, thin=5,
                         logl_params=lpar, logp_params=ppar, record_rows=niter // 5 + 1, device=dev) for dev in (0, 1)]
    for e in engs:
        e.set_state(x0)
    for _ in range(3):          # interleaved calls
        for e in engs:
            e.run(niter // 3)
    a, b = engs[0].state(), engs[1].state()
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert np.array_equal(engs[0].chain()[0], engs[1].chain()[0])


def test_run_refuses_iterations_beyond_the_random_streams_range():
    """The Philox counter holds the iteration in 32 bits (round-1 advisor finding): a run that would pass 2^32 - 1 is
    refused instead of repeating its draws."""
    d, W, T = 3, 4, 2
    e = _cabi.Engine(d, W, T, np.eye(d) * 0.1, orc.temperature_ladder(d, T), seed=1, logl_kind=orc.LOGL_GAUSSIAN,
                     logl_params=orc.gaussian_params(np.zeros(d), np.eye(d)), logp_kind=orc.LOGP_UNIFORM,
                     logp_params=orc.uniform_params(-5 * np.ones(d), 5 * np.ones(d)))
    e.set_state(np.zeros((T, W, d)))
    with pytest.raises(_cabi.EngineError):
        e.run(2 ** 32)
    e.run(10)
    e.close()
