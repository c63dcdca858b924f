"""Ladder sharding (BASELINE config 5): the swap sweep cut at shard boundaries must reproduce the
unsharded sweep bit for bit.  CPU part: the exchange protocol of ptmcmcsampler_b200.distributed driven
against the oracle (host pointers), in one process and over a world-size-2/3 gloo group.  GPU part:
several CUDA engine shards on one device against the unsharded CUDA engine and the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as orc
from ptmcmcsampler_b200 import distributed as dist_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def problem(d=5, W=7, Tg=6, seed=0, cov_update=50, burn=100, tskip=10):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    kw = dict(seed=3 + seed, cov_update=cov_update, burn=burn, tskip=tskip, thin=5,
              logl_params=orc.gaussian_params(5 * np.ones(d), np.linalg.inv(cov)),
              logp_params=orc.uniform_params(-50 * np.ones(d), 60 * np.ones(d)), record_hot=True)
    return kw, orc.temperature_ladder(d, Tg), rng.uniform(0, 10, (Tg, W, d))


def full_oracle(d, W, Tg, N, kw, ladder, x0):
    o = orc.Oracle(d, W, Tg, 0.01 * np.eye(d), ladder=ladder, max_rows=N // kw["thin"] + 1, **kw)
    o.set_trace(N, N // kw["tskip"])
    o.set_state(x0)
    o.run(N)
    return o


def oracle_shard(d, W, N, kw, ladder, x0, G, g):
    sk = dist_mod.ladder_shard_kwargs(ladder, G, g)
    T = sk.pop("ntemps")
    o = orc.Oracle(d, W, T, 0.01 * np.eye(d), max_rows=N // kw["thin"] + 1, **sk, **kw)
    o.set_trace(N, N // kw["tskip"])
    o.set_state(x0[sk["temp_offset"]:sk["temp_offset"] + T])
    return o


def assert_shards_equal_full(shards, full):
    assert np.array_equal(np.concatenate([s.state()[0] for s in shards]), full.state()[0])
    assert np.array_equal(np.concatenate([s.state()[1] for s in shards]), full.state()[1])
    assert np.array_equal(np.concatenate([s.trace for s in shards], axis=1), full.trace)
    assert np.array_equal(np.concatenate([s.swapmaps for s in shards], axis=2), full.swapmaps)
    assert np.array_equal(np.concatenate([s.counters()[2] for s in shards]), full.counters()[2])
    assert np.array_equal(np.concatenate([s.chain()[0] for s in shards], axis=1), full.chain()[0])
    assert all(s.counters()[3] == full.counters()[3] for s in shards)


@pytest.mark.parametrize("Tg,G", [(6, 3), (6, 2), (4, 4)])
def test_protocol_in_one_process_matches_unsharded_oracle(Tg, G):
    d, W, N = 5, 7, 300
    kw, ladder, x0 = problem(d, W, Tg)
    full = full_oracle(d, W, Tg, N, kw, ladder, x0)
    shards = [oracle_shard(d, W, N, kw, ladder, x0, G, g) for g in range(G)]
    dist_mod.run_ladder_local(shards, 130, kw["tskip"], dist_mod.HostMem())
    dist_mod.run_ladder_local(shards, N - 130, kw["tskip"], dist_mod.HostMem())
    assert_shards_equal_full(shards, full)
    assert np.allclose(shards[0].adapt()[0], full.adapt()[0], rtol=0, atol=0)
    for s in shards[1:]:  # the cold shard's factor and DE history reached every shard
        assert np.array_equal(s.factor()[0], full.factor()[0])
        assert np.array_equal(s.buffers()[1], full.buffers()[1])


@pytest.mark.parametrize("cov_update,burn,tskip,first", [(25, 75, 10, 130), (30, 90, 7, 64), (25, 75, 10, 75)])
def test_covariance_and_de_boundaries_off_the_swap_grid(cov_update, burn, tskip, first):
    """covUpdate / burn that are not multiples of Tskip (nor of the caller's chunking): the shards must still stop at every
    boundary for the factor broadcast and the AM-ring hand-over, or the hot shards keep a stale factor and DE history."""
    d, W, Tg, G, N = 5, 7, 6, 3, 300
    kw, ladder, x0 = problem(d, W, Tg, cov_update=cov_update, burn=burn, tskip=tskip)
    full = full_oracle(d, W, Tg, N, kw, ladder, x0)
    shards = [oracle_shard(d, W, N, kw, ladder, x0, G, g) for g in range(G)]
    dist_mod.run_ladder_local(shards, first, tskip, dist_mod.HostMem())
    dist_mod.run_ladder_local(shards, N - first, tskip, dist_mod.HostMem())
    assert_shards_equal_full(shards, full)
    for s in shards[1:]:
        assert np.array_equal(s.factor()[0], full.factor()[0])
        assert np.array_equal(s.buffers()[1], full.buffers()[1])


def test_shard_refuses_to_cross_a_swap_iteration():
    d, W, N = 5, 3, 50
    kw, ladder, x0 = problem(d, W, 4)
    s = oracle_shard(d, W, N, kw, ladder, x0, 2, 0)
    with pytest.raises(ValueError):
        s.run(15)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, Tg, N, out, sched=(50, 100, 10)):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, W = 5, 7
        kw, ladder, x0 = problem(d, W, Tg, cov_update=sched[0], burn=sched[1], tskip=sched[2])
        shard = oracle_shard(d, W, N, kw, ladder, x0, world, rank)
        comm = dist_mod.LadderComm(shard, device="cpu")
        dist_mod.run_ladder(shard, 130, comm, kw["tskip"])
        dist_mod.run_ladder(shard, N - 130, comm, kw["tskip"])
        np.savez(out % rank, x=shard.state()[0], lnl=shard.state()[1], trace=shard.trace, swapmaps=shard.swapmaps,
                 swap_acc=shard.counters()[2], chain=shard.chain()[0], U=shard.factor()[0], de=shard.buffers()[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Tg,G,sched", [(6, 2, (50, 100, 10)), (6, 3, (50, 100, 10)), (6, 3, (25, 75, 10))])
def test_gloo_neighbour_exchange_matches_unsharded_oracle(Tg, G, sched, tmp_path):
    import torch.multiprocessing as mp

    d, W, N = 5, 7, 300
    kw, ladder, x0 = problem(d, W, Tg, cov_update=sched[0], burn=sched[1], tskip=sched[2])
    full = full_oracle(d, W, Tg, N, kw, ladder, x0)
    out = str(tmp_path / "shard%d.npz")
    mp.spawn(_gloo_worker, args=(G, _free_port(), Tg, N, out, sched), nprocs=G, join=True)
    r = [np.load(out % g) for g in range(G)]
    assert np.array_equal(np.concatenate([a["x"] for a in r]), full.state()[0])
    assert np.array_equal(np.concatenate([a["lnl"] for a in r]), full.state()[1])
    assert np.array_equal(np.concatenate([a["trace"] for a in r], axis=1), full.trace)
    assert np.array_equal(np.concatenate([a["swapmaps"] for a in r], axis=2), full.swapmaps)
    assert np.array_equal(np.concatenate([a["swap_acc"] for a in r]), full.counters()[2])
    assert np.array_equal(np.concatenate([a["chain"] for a in r], axis=1), full.chain()[0])
    for a in r[1:]:
        assert np.array_equal(a["U"], full.factor()[0]) and np.array_equal(a["de"], full.buffers()[1])


def _gloo_walker_worker(rank, world, port, N, out):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, Wtot, T = 5, 8, 3
        kw, ladder, x0 = problem(d, Wtot, T)
        W = Wtot // world
        o = orc.Oracle(d, W, T, 0.01 * np.eye(d), ladder=ladder, max_rows=N // kw["thin"] + 1, walker_offset=rank * W, **kw)
        o.set_trace(N, N // kw["tskip"])
        o.set_state(x0[:, rank * W:(rank + 1) * W])
        dist_mod.run(o, 70, None)
        dist_mod.run(o, N - 70, None)
        np.savez(out % rank, x=o.state()[0], trace=o.trace, cov=o.adapt()[0], n=o.adapt()[3])
    finally:
        dist.destroy_process_group()


def test_gloo_walker_sharding_pools_the_covariance(tmp_path):
    """Walker sharding (bench.py's N > 1 path): two ranks with half the walkers each and an all-gather of
    the batch moments at every covariance boundary reproduce the single-process run.  The pooled moments
    are merged (Chan) instead of accumulated sample by sample, so covariances agree to rounding.  The run
    stops at `burn`: from then on DE draws from the shard-local history, which by design differs from the
    single-process pooled history."""
    import torch.multiprocessing as mp

    d, Wtot, T, N = 5, 8, 3, 100
    kw, ladder, x0 = problem(d, Wtot, T)
    full = orc.Oracle(d, Wtot, T, 0.01 * np.eye(d), ladder=ladder, max_rows=N // kw["thin"] + 1, **kw)
    full.set_trace(N, N // kw["tskip"])
    full.set_state(x0)
    full.run(N)
    out = str(tmp_path / "w%d.npz")
    mp.spawn(_gloo_walker_worker, args=(2, _free_port(), N, out), nprocs=2, join=True)
    r = [np.load(out % g) for g in range(2)]
    assert r[0]["n"] == r[1]["n"] == full.adapt()[3]
    assert np.allclose(r[0]["cov"], full.adapt()[0], rtol=1e-9, atol=1e-12) and np.array_equal(r[0]["cov"], r[1]["cov"])
    tr = np.concatenate([a["trace"] for a in r], axis=2)
    assert np.array_equal(tr, full.trace)
    assert np.allclose(np.concatenate([a["x"] for a in r], axis=1), full.state()[0], rtol=1e-9, atol=1e-9)


def test_merge_batches_is_chan_merge():
    rng = np.random.default_rng(1)
    d = 4
    xs = [rng.standard_normal((n, d)) + 3 for n in (50, 70, 31)]
    batches = []
    for x in xs:
        m = x.mean(0)
        batches.append(np.concatenate([[len(x)], m, ((x - m).T @ (x - m)).ravel()]))
    merged = dist_mod.merge_batches(batches)
    allx = np.concatenate(xs)
    assert merged[0] == len(allx)
    assert np.allclose(merged[1:1 + d], allx.mean(0))
    assert np.allclose(merged[1 + d:].reshape(d, d) / (len(allx) - 1), np.cov(allx.T))


# ------------------------------------------------------------------------------------ GPU -----
@pytest.mark.gpu
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("d,W,Tg,G", [(5, 33, 6, 3), (20, 64, 8, 2), (8, 16, 4, 4), (20, 48, 256, 8)])
def test_cuda_shards_on_one_device_match_unsharded_engine_and_oracle(d, W, Tg, G, p2p):
    """p2p: the swap messages are stored into the neighbour's mailbox by the kernels themselves (ptmcmc_swap_p2p) instead
    of being handed over between the three steps.  The last case is BASELINE config 5's ladder: 256 rungs as 8 shards
    of 32."""
    from ptmcmcsampler_b200 import _cabi

    N = 300
    kw, ladder, x0 = problem(d, W, Tg, seed=d)
    full_o = full_oracle(d, W, Tg, N, kw, ladder, x0)
    rows = N // kw["thin"] + 1
    ekw = dict(kw)
    ekw.update(record_rows=rows, trace_iters=N)
    full_g = _cabi.Engine(d, W, Tg, 0.01 * np.eye(d), ladder, **ekw)
    full_g.set_state(x0)
    full_g.run(N)
    shards = []
    for g in range(G):
        sk = dist_mod.ladder_shard_kwargs(ladder, G, g)
        T = sk.pop("ntemps")
        lad = sk.pop("ladder")
        e = _cabi.Engine(d, W, T, 0.01 * np.eye(d), lad, **sk, **ekw)
        e.set_state(x0[sk["temp_offset"]:sk["temp_offset"] + T])
        shards.append(e)
    with pytest.raises(_cabi.EngineError):
        shards[0].run(kw["tskip"] + 1)
    if p2p:
        dist_mod.connect_local_p2p(shards)
    dist_mod.run_ladder_local(shards, 130, kw["tskip"], dist_mod.CudaMem(0), p2p=p2p)
    dist_mod.run_ladder_local(shards, N - 130, kw["tskip"], dist_mod.CudaMem(0), p2p=p2p)
    nsw = N // kw["tskip"]
    tr = np.concatenate([s.trace(N, nsw)[0] for s in shards], axis=1)
    sm = np.concatenate([s.trace(N, nsw)[1] for s in shards], axis=2)
    ftr, fsm = full_g.trace(N, nsw)
    # sharded CUDA == unsharded CUDA, bit for bit (same kernels' arithmetic, same draws)
    assert np.array_equal(tr, ftr) and np.array_equal(sm, fsm)
    xs = np.concatenate([s.state()[0] for s in shards])
    assert np.array_equal(xs, full_g.state()[0])
    assert np.array_equal(np.concatenate([s.counters()[2] for s in shards]), full_g.counters()[2])
    assert np.array_equal(np.concatenate([s.chain()[0] for s in shards], axis=1), full_g.chain()[0])
    # and the oracle: integers exact, floats to 1e-9
    assert np.array_equal(tr, full_o.trace) and np.array_equal(sm, full_o.swapmaps[:nsw])
    assert np.allclose(xs, full_o.state()[0], rtol=1e-9, atol=1e-9)
    for s in shards:
        s.close()
    full_g.close()
