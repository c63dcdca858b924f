"""Ladder sharding over NCCL (run under torchrun, one rank per GPU): (1) bit-exact check of the sharded
run against the unsharded engine, (2) throughput of a C5-like shard (32 rungs x 8192 walkers per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/ladder_nccl_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptmcmcsampler_b200 import _cabi, distributed as dm  # noqa: E402


def target(d):
    rng = np.random.default_rng(20)
    A = 0.5 - rng.random((d, d))
    A = np.triu(A)
    A += A.T - np.diag(A.diagonal())
    cov = A @ A + 0.1 * np.eye(d)
    lpar = np.concatenate([5.0 * np.ones(d), np.linalg.inv(cov).ravel(), [0.0]])
    ppar = np.concatenate([-50 * np.ones(d), 60 * np.ones(d), [0.0, 1.0]])
    return lpar, ppar


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- (1) parity
    d, W, Tg, N = 20, 256, 4 * world, 400
    lpar, ppar = target(d)
    ladder = (1 + np.sqrt(2.0 / d)) ** np.arange(Tg)
    kw = dict(seed=5, cov_update=100, burn=100, tskip=10, thin=10, logl_params=lpar, logp_params=ppar, record_hot=True,
              record_rows=N // 10 + 1, trace_iters=N, device=local)
    x0 = np.random.default_rng(1).uniform(0, 10, (Tg, W, d))
    sk = dm.ladder_shard_kwargs(ladder, world, rank)
    T = sk.pop("ntemps")
    lad = sk.pop("ladder")
    e = _cabi.Engine(d, W, T, 0.01 * np.eye(d), lad, **sk, **kw)
    e.set_state(x0[sk["temp_offset"]:sk["temp_offset"] + T])
    comm = dm.LadderComm(e)
    dm.run_ladder(e, N, comm, 10)
    comm.check(e)
    e.sync()
    if rank == 0:
        print("swap messages through", "peer memory (ptmcmc_swap_p2p)" if comm.p2p else "NCCL send / receive", flush=True)
    mine = dict(x=e.state()[0], tr=e.trace(N, N // 10)[0], sm=e.trace(N, N // 10)[1], sw=e.counters()[2], ch=e.chain()[0])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        f = _cabi.Engine(d, W, Tg, 0.01 * np.eye(d), ladder, **kw)
        f.set_state(x0)
        f.run(N)
        ok = (np.array_equal(np.concatenate([p["x"] for p in parts]), f.state()[0])
              and np.array_equal(np.concatenate([p["tr"] for p in parts], axis=1), f.trace(N, N // 10)[0])
              and np.array_equal(np.concatenate([p["sm"] for p in parts], axis=2), f.trace(N, N // 10)[1])
              and np.array_equal(np.concatenate([p["sw"] for p in parts]), f.counters()[2])
              and np.array_equal(np.concatenate([p["ch"] for p in parts], axis=1), f.chain()[0]))
        print("LADDER NCCL PARITY", "OK" if ok else "FAILED", "world", world, flush=True)
        f.close()
    e.close()
    # ---- (2) throughput: 32 rungs x 8192 walkers per GPU, Tskip=100 (BASELINE config 5 shape)
    d, W, Tl, N = 20, 8192, 32, 2000
    Tg = Tl * world
    ladder = (1 + np.sqrt(2.0 / d)) ** np.arange(Tg)
    ladder = np.minimum(ladder, 1e30)
    kw = dict(seed=9, cov_update=1000, burn=1000, tskip=100, thin=10, logl_params=lpar, logp_params=ppar,
              record_rows=(3 * N) // 10 + 2, device=local)
    sk = dm.ladder_shard_kwargs(ladder, world, rank)
    T = sk.pop("ntemps")
    lad = sk.pop("ladder")
    e = _cabi.Engine(d, W, T, 0.01 * np.eye(d), lad, **sk, **kw)
    e.set_state(np.random.default_rng(2 + rank).uniform(0, 10, (T, W, d)))
    comm = dm.LadderComm(e)
    dm.run_ladder(e, 1100, comm, 100)   # warm-up past the first covariance / DE update
    e.sync()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dm.run_ladder(e, N, comm, 100)
    e.sync()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        rate = world * W * T * N / float(dt.item())
        print("LADDER NCCL THROUGHPUT world %d: %.3e chain-steps/s (%d rungs x %d walkers, %.3f s for %d iterations)"
              % (world, rate, Tg, W, float(dt.item()), N), flush=True)
    e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
