#!/usr/bin/env python
"""Benchmark of the PT-MCMC hot path (BASELINE.json metric and config).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): 20-dim correlated Gaussian target,
8192 walkers x 32 temperatures per GPU, SCAM/AM/DE = 20/20/20, covUpdate = burn = 1000, Tskip = 100,
thin = 10, default geometric ladder.  One bench "step" = 1000 MH iterations of all 262 144 chains
(10 swap sweeps, one pooled covariance update + eigen-factorisation and one DE-history update
included).  N > 1: every rank runs its own 8192 x 32 shard (weak scaling); the only collective is the
pooled-covariance all-gather at every covariance boundary.

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes
through the public PTSampler.sample() call with host buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, W, T = 20, 8192, 32
ITERS = 1000          # MH iterations per bench step
COV_UPDATE = BURN = 1000
TSKIP, THIN = 100, 10
WEIGHTS = (20, 20, 20)
METRIC = "walker-steps/sec (20-dim Gaussian, 8192 walkers x 32 temps)"
UNIT = "chain-steps/s"


def algorithmic_bytes_per_chain_step(d=D, t=T, thin=THIN, p_de=1.0 / 3.0):
    """state read+write (x, lnL, lnP, 16-B RNG counter), two DE rows on DE
    steps, the thinned record and the cold rung's AM-ring write."""
    return 2 * (8 * d + 16 + 16) + p_de * 2 * 8 * d + (8 * d + 16) / thin + 8 * d / t


def algorithmic_flops_per_chain_step(d=D, p_am=1.0 / 3.0, p_scam=1.0 / 3.0, p_de=1.0 / 3.0):
    """dense Gaussian logl 2d^2+3d, box prior 2d, AM mat-vecs 2d^2 (the engine's
    x + U delta form), SCAM / DE 2d, ~30 for the Hastings test."""
    return (2 * d * d + 3 * d) + 2 * d + p_am * 2 * d * d + (p_scam + p_de) * 2 * d + 30


def problem():
    """C2 target: mu = 5, Sigma = A.A + 0.1 I with A as in examples/simple.py:27-30 from default_rng(20)."""
    rng = np.random.default_rng(20)
    A = 0.5 - rng.random(D * D).reshape(D, D)
    A = np.triu(A)
    A += A.T - np.diag(A.diagonal())
    cov = A @ A + 0.1 * np.eye(D)
    mu = 5.0 * np.ones(D)
    ladder = (1 + np.sqrt(2.0 / D)) ** np.arange(T)
    return mu, cov, ladder


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The sampler is started before the warm-up
    (the tool takes a few hundred ms to come up) and only the samples stamped inside the window given to
    stop() are summarised."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in open(self.path):
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v == "Active"]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        if inside:
            out["window"] = "timed region"
            rows = inside
        elif rows:
            out["window"] = "whole run (no sample fell inside the timed region)"
        sm, mx, reasons = [r[1] for r in rows], [r[2] for r in rows], set(n for r in rows for n in r[3])
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def cpu_port_rate(walkers, iters, threads, seed=3):
    """Time the CPU oracle (plain-C port of the reference algorithm) on a bounded sample of the
    same workload: `walkers` x 32 temperatures x `iters` iterations."""
    from oracle import oracle as orc

    mu, cov, ladder = problem()
    o = orc.Oracle(D, walkers, T, 0.01 * np.eye(D), seed=seed, ladder=ladder,
                   cycle=((orc.JUMP_SCAM, WEIGHTS[0]), (orc.JUMP_AM, WEIGHTS[1])), de_weight=WEIGHTS[2],
                   cov_update=COV_UPDATE, burn=BURN, tskip=TSKIP, thin=THIN,
                   logl_params=orc.gaussian_params(mu, np.linalg.inv(cov)),
                   logp_params=orc.uniform_params(-50 * np.ones(D), 60 * np.ones(D)),
                   max_rows=(2 * iters) // THIN + 2, nthreads=threads)
    o.set_state(np.random.default_rng(1).uniform(0, 10, (T, walkers, D)))
    o.run(BURN + 1)          # DE joins the cycle, covariance adapted once: the steady-state mix
    t0 = time.perf_counter()
    o.run(iters)
    dt = time.perf_counter() - t0
    return walkers * T * iters / dt, dt


def run_reference(args, rank, world):
    """Reference arm: the reference is pure Python and cannot travel to the GPU box, so this times
    its plain-C restatement (oracle/, kind "port") on all host cores, same config and metric."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    walkers, iters = 128 * threads, 1000
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(walkers, 100, threads)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_port_rate(walkers, iters, threads)
        rates.append(r)
        times.append(dt)
    value = float(walkers * T * iters * len(times) / sum(times))
    sample = "%d walkers x %d temps x %d iterations per step, OpenMP over walkers" % (walkers, T, iters)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 20-dim Gaussian, SCAM/AM/DE 20/20/20, covUpdate=burn=1000, Tskip=100, thin=10; "
                               "CPU sample " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ladder_steps_per_s": value / T,
    }
    print(json.dumps(line), flush=True)


def run_engine(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ptmcmcsampler_b200 import PTMCMCSampler, _cabi, distributed
    from ptmcmcsampler_b200.likelihoods import GaussianLikelihood, UniformPrior

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    group = None

    mu, cov, ladder = problem()
    icov = np.linalg.inv(cov)
    lpar = np.concatenate([mu, icov.ravel(), [0.0]])
    ppar = np.concatenate([-50 * np.ones(D), 60 * np.ones(D), [0.0, 1.0]])
    total_iters = ITERS * (2 * args.steps + args.warmup)
    ladder_mode = args.shard == "ladder" and world > 1
    shard_kw, Tg = dict(ntemps=T, ladder=ladder, walker_offset=rank * W), T
    if ladder_mode:  # BASELINE config 5: T rungs per GPU of one ladder of T * N rungs, same walkers everywhere
        Tg = T * world
        ladder = np.minimum((1 + np.sqrt(2.0 / D)) ** np.arange(Tg), 1e30)
        shard_kw = distributed.ladder_shard_kwargs(ladder, world, rank)
    eng = _cabi.Engine(D, W, shard_kw.pop("ntemps"), 0.01 * np.eye(D), shard_kw.pop("ladder"), seed=42,
                       cycle=((0, WEIGHTS[0]), (1, WEIGHTS[1])), de_weight=WEIGHTS[2], cov_update=COV_UPDATE, burn=BURN,
                       tskip=TSKIP, thin=THIN, logl_params=lpar, logp_params=ppar, record_rows=total_iters // THIN + 2,
                       device=local_rank, timing=False, **shard_kw)
    x0 = np.random.default_rng(1 + rank).uniform(0, 10, (T, W, D))
    eng.set_state(x0)
    comm = distributed.LadderComm(eng) if ladder_mode else None

    def step():
        if ladder_mode:
            distributed.run_ladder(eng, ITERS, comm, TSKIP)
        elif world > 1:
            distributed.run(eng, ITERS, group)
        else:
            eng.run(ITERS)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    eng.sync()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.reset_timing()
    t_wall0 = time.time()
    # timed region: exactly K steps, nothing but the engine's own launches on its stream
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    gpu_launches = int(sum(eng.timing()["launches"].values()))
    # same K steps again with every launch bracketed by CUDA events on the engine's stream: the
    # per-kernel-class durations behind `roofline` (the bracketing serialises host and device, so this
    # pass is not the one `value` is taken from)
    eng.reset_timing()
    eng.set_timing(True)
    for _ in range(args.steps):
        step()
    eng.sync()
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    tm = eng.timing()
    eng.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    chain_steps_rank = W * T * ITERS * args.steps
    value = world * chain_steps_rank / (ms * 1e-3)

    # roofline of the dominant kernel (fused MH segment): algorithmic bytes per launch / mean duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpcs = algorithmic_bytes_per_chain_step()
    mh_launches, mh_ms = tm["launches"]["mh"], tm["ms"]["mh"]
    steps_per_launch = W * T * ITERS * args.steps / max(1, mh_launches)
    achieved = bpcs * steps_per_launch / (mh_ms / max(1, mh_launches) * 1e-3) / 1e9
    fp64_peak = 37.1e12  # measured on this pool's B200: scripts/micro/dmma_rate.cu (DFMA and DMMA alike)
    flops = algorithmic_flops_per_chain_step()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "mh_sorted_kernel<20,256,2>", "peak_source": peak_src,
                "fp64": {"algorithmic_flops_per_chain_step": flops, "peak_tflops": fp64_peak / 1e12,
                         "achieved_tflops": flops * steps_per_launch / (mh_ms / max(1, mh_launches) * 1e-3) / 1e12,
                         "frac": flops * steps_per_launch / (mh_ms / max(1, mh_launches) * 1e-3) / fp64_peak,
                         "note": "the launch keeps chain state on chip for Tskip iterations, so fp64 issue, not HBM, "
                                 "is the binding resource (SURVEY 8d)"},
                "algorithmic_bytes_per_chain_step": bpcs, "launches": mh_launches,
                "avg_launch_ms": mh_ms / max(1, mh_launches),
                "kernel_share_of_step": mh_ms / sum(tm["ms"].values()) if sum(tm["ms"].values()) > 0 else None,
                "class_ms": tm["ms"]}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("mh_kernel_dram_bytes_per_launch")
        except Exception:
            pass
    eng.close()

    # end to end through the public API: host p0 in, recorded chain out, every step
    lk, pr = GaussianLikelihood(mu, icov=icov), UniformPrior(-50.0, 60.0)
    p0 = _cabi.pinned_empty((T, W, D))
    p0[...] = np.random.default_rng(7 + rank).uniform(0, 10, (T, W, D))
    outdir = tempfile.mkdtemp(prefix="ptmcmc_bench_")

    def e2e_step(seed):
        s = PTMCMCSampler.PTSampler(D, lk, pr, 0.01 * np.eye(D), outDir=outdir, verbose=False, seed=seed, ntemps=Tg,
                                    nwalkers=W, device=local_rank, walker_offset=0 if ladder_mode else rank * W,
                                    dist_group=True if world > 1 else None, shard="ladder" if ladder_mode else "walkers")
        s.sample(p0, ITERS, burn=BURN, covUpdate=COV_UPDATE, Tskip=TSKIP, thin=THIN, isave=ITERS,
                 SCAMweight=WEIGHTS[0], AMweight=WEIGHTS[1], DEweight=WEIGHTS[2],
                 ladder=ladder if ladder_mode else None)
        loss = float(s._lnlike_all[-1].mean())   # the step's result read on the host
        d2h = s._chain_all.nbytes + s._lnlike_all.nbytes + s._lnprob_all.nbytes
        s.engine.close()
        return loss, d2h

    e2e_step(100)
    nrep = max(1, min(args.steps, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(nrep):
        _, d2h = e2e_step(101 + r)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * W * T * ITERS * nrep / dt, "unit": UNIT, "h2d_bytes_per_step": int(p0.nbytes),
           "d2h_bytes_per_step": int(d2h), "steps": nrep,
           "note": "PTSampler(...).sample(p0_host, 1000): engine build, H2D of p0, 1000 iterations, D2H of the "
                   "thinned T=1 record of all walkers, chain file of walker 0"}

    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        walkers = 128 * threads
        rate, dtc = cpu_port_rate(walkers, 1000, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d walkers x %d temps x 1000 iterations (%.1f s)" % (walkers, T, dtc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 20-dim Gaussian, 8192 walkers x 32 temps per GPU, SCAM/AM/DE 20/20/20, "
                                   "covUpdate=burn=1000, Tskip=100, thin=10; step = 1000 MH iterations of all chains",
                       "l2": "no flush needed: each step streams the 1.3 GB AM ring and gathers from the 1.3 GB DE "
                             "history (inputs >> 126 MB L2)",
                       "parallelism": ("ladder-sharded x%d (%d rungs, neighbour exchange of the boundary rung)" % (world, Tg)
                                       if ladder_mode else "walker-sharded x%d" % world)},
            "ladder_steps_per_s": value / T,
            "clocks": clk, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--shard", default="walkers", choices=["walkers", "ladder"],
                    help="N > 1: independent walkers per GPU (default, the contract's weak scaling) or one ladder of "
                         "32 N rungs split over the GPUs (BASELINE config 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
