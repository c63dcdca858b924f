"""Walker sharding across GPUs (one process per GPU, ``torch.distributed``).

Every rank owns ``nwalkers`` complete ladders (``walker_offset = rank * nwalkers`` keys the RNG, so a
walker's draws do not depend on the sharding).  The MH step and the swap need no communication.  The
only coupling is the pooled proposal covariance: at every covariance boundary the ranks exchange
their batch moments {n, mean[d], M2c[d*d]} (d*d+d+1 doubles), merge them with Chan's formula in rank
order, and each applies the merged batch -- the multi-device form of the reference's rank-0
``send(cov)`` broadcast (ref PTMCMCSampler.py:545-560).  The DE history stays shard-local.
"""
import numpy as np


def merge_batches(batches):
    """Chan merge of [{n, mean[d], M2c[d*d]}, ...] in list order (deterministic)."""
    batches = [np.asarray(b, dtype=np.float64) for b in batches]
    d = int(round((-1 + np.sqrt(1 + 4 * (len(batches[0]) - 1))) / 2))
    n, mean, m2 = 0.0, np.zeros(d), np.zeros((d, d))
    for b in batches:
        nb, mb, m2b = b[0], b[1:1 + d], b[1 + d:].reshape(d, d)
        if nb == 0:
            continue
        delta = mb - mean
        tot = n + nb
        m2 = m2 + m2b + np.outer(delta, delta) * (n * nb / tot)
        mean = mean + delta * (nb / tot)
        n = tot
    return np.concatenate([[n], mean, m2.ravel()])


def pooled_adapt(engine, group=None):
    """If a covariance update is due, pool the batch moments over the ranks of ``group`` and apply.
    Collective: every rank must call it at the same iteration.  Returns True if an update ran."""
    import torch
    import torch.distributed as dist

    batch = engine.adapt_begin()
    if batch is None:
        return False
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(batch).to(dev)
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine, group=group)
        batch = merge_batches([p.cpu().numpy() for p in parts])
    engine.adapt_finish(batch)
    return True


def run(engine, niter, group=None):
    """``engine.run(niter)`` with pooled covariance updates at every boundary."""
    cu = engine.cov_update
    done = 0
    while done < niter:
        it = engine.iteration
        pooled_adapt(engine, group)          # boundary reached by a previous call
        step = min(niter - done, cu - it % cu)
        engine.run(step)
        done += step
    return done
