#!/usr/bin/env python
"""Benchmark of the PT-MCMC hot path (BASELINE.json metric and configs).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--shard walkers|ladder]

Headline workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): 20-dim correlated Gaussian target,
8192 walkers x 32 temperatures per GPU, SCAM/AM/DE = 20/20/20, covUpdate = burn = 1000, Tskip = 100,
thin = 10, default geometric ladder.  One bench "step" = 1000 MH iterations of all 262 144 chains
(10 swap sweeps, one pooled covariance update + eigen-factorisation and one DE-history update
included).  N > 1: every rank runs its own 8192 x 32 shard (weak scaling); the only collective is the
pooled-covariance all-gather at every covariance boundary.

The same JSON line also carries the other BASELINE configurations, each measured in this run with its own
`value`, `ms_per_step`, `roofline` and `e2e`: `configs.C3` (100-dim dense Gaussian, 4096 x 64, tensor-core
kernel) and `configs.C4` (curved 10-dim, 16384 x 128, DE-dominant) at N = 1, and `c5_ladder` (one ladder of
32 N rungs split 32 per GPU, neighbour exchange of the boundary rung over NCCL) at N > 1.

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes
through the public PTSampler.sample() call with host buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 1000          # MH iterations per bench step
E2E_ITERS = 2000      # iterations of one end-to-end sample() call
COV_UPDATE = BURN = 1000
TSKIP, THIN = 100, 10
METRIC = "walker-steps/sec (20-dim Gaussian, 8192 walkers x 32 temps)"
UNIT = "chain-steps/s"


# ------------------------------------------------------------------------------------------ workloads
def algorithmic_bytes_per_chain_step(d, t, thin=THIN, p_de=1.0 / 3.0):
    """state read+write (x, lnL, lnP, 16-B RNG counter), two DE rows on DE
    steps, the thinned record and the cold rung's AM-ring write."""
    return 2 * (8 * d + 16 + 16) + p_de * 2 * 8 * d + (8 * d + 16) / thin + 8 * d / t


def algorithmic_flops_per_chain_step(d, f_logl, p_am, p_scam, p_de):
    """logl, box prior 2d, AM mat-vecs 2d^2 (the engine's x + U delta form), SCAM / DE 2d,
    ~30 for the Hastings test."""
    return f_logl + 2 * d + p_am * 2 * d * d + (p_scam + p_de) * 2 * d + 30


class Workload(object):
    """One BASELINE configuration: sizes, target, proposal mix, and what the engine / sampler need."""

    def __init__(self, name):
        from ptmcmcsampler_b200 import _cabi

        self.name = name
        if name == "C2":
            d, W, T = 20, 8192, 32
            rng = np.random.default_rng(20)  # Sigma = A.A + 0.1 I, A as in ref examples/simple.py:27-30
            A = 0.5 - rng.random(d * d).reshape(d, d)
            A = np.triu(A)
            A += A.T - np.diag(A.diagonal())
            self.cov = A @ A + 0.1 * np.eye(d)
            self.mu = 5.0 * np.ones(d)
            self.box, self.inclusive = (-50.0, 60.0), True
            self.weights = (20, 20, 20)
            self.cov0 = 0.01 * np.eye(d)
            self.x0 = lambda seed, T_, W_: np.random.default_rng(seed).uniform(0, 10, (T_, W_, d))
            self.logl_kind = _cabi.LOGL_GAUSSIAN
            self.desc = "C2: 20-dim Gaussian, 8192 walkers x 32 temps per GPU, SCAM/AM/DE 20/20/20"
        elif name == "C3":
            d, W, T = 100, 4096, 64
            s = np.logspace(-1, 1, d)
            idx = np.arange(d)
            self.cov = 0.9 ** np.abs(idx[:, None] - idx[None, :]) * s[:, None] * s[None, :]
            self.mu = np.zeros(d)
            self.box, self.inclusive = (-500.0, 500.0), True
            self.weights = (20, 20, 20)
            self.cov0 = np.diag(0.01 * s * s)
            self.x0 = lambda seed, T_, W_: np.random.default_rng(seed).standard_normal((T_, W_, d)) * s
            self.logl_kind = _cabi.LOGL_GAUSSIAN
            self.desc = "C3: 100-dim correlated Gaussian (dense cov 0.9^|i-j| s_i s_j), 4096 walkers x 64 temps, SCAM/AM/DE 20/20/20"
        elif name == "C4":
            d, W, T = 10, 16384, 128
            self.cov, self.mu = None, None
            self.box, self.inclusive = (-10.0, 10.0), False
            self.weights = (10, 10, 60)
            self.cov0 = 0.1 * np.eye(d)
            self.x0 = lambda seed, T_, W_: np.random.default_rng(seed).uniform(-1, 1, (T_, W_, d))
            self.logl_kind = _cabi.LOGL_CURVED
            self.desc = ("C4: curved 10-dim (five copies of the reference's 2-D curved density), 16384 walkers x 128 temps, "
                         "SCAM/AM/DE 10/10/60")
        else:
            raise ValueError(name)
        self.d, self.W, self.T = d, W, T
        self.ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(T), 1e30)
        ws = float(sum(self.weights))
        self.p_scam, self.p_am, self.p_de = (w / ws for w in self.weights)
        self.icov = np.linalg.inv(self.cov) if self.cov is not None else None
        # flops of one log-likelihood: dense Gaussian 2d^2+3d; curved ~ 2 exp + 1 log + 20 flops per pair
        self.f_logl = (2 * d * d + 3 * d) if self.cov is not None else (d // 2) * 80
        self.bytes_per_step = algorithmic_bytes_per_chain_step(d, T, THIN, self.p_de)
        self.flops_per_step = algorithmic_flops_per_chain_step(d, self.f_logl, self.p_am, self.p_scam, self.p_de)

    def engine_kwargs(self):
        from ptmcmcsampler_b200 import _cabi

        d = self.d
        kw = dict(cycle=((_cabi.JUMP_SCAM, self.weights[0]), (_cabi.JUMP_AM, self.weights[1])), de_weight=self.weights[2],
                  cov_update=COV_UPDATE, burn=BURN, tskip=TSKIP, thin=THIN, logl_kind=self.logl_kind,
                  logp_params=np.concatenate([self.box[0] * np.ones(d), self.box[1] * np.ones(d),
                                              [0.0, 1.0 if self.inclusive else 0.0]]))
        if self.icov is not None:
            kw["logl_params"] = np.concatenate([self.mu, self.icov.ravel(), [0.0]])
        return kw

    def targets(self):
        from ptmcmcsampler_b200.likelihoods import CurvedLikelihood, GaussianLikelihood, UniformPrior

        lk = GaussianLikelihood(self.mu, icov=self.icov) if self.icov is not None else CurvedLikelihood()
        return lk, UniformPrior(self.box[0], self.box[1], inclusive=self.inclusive)

    def sample_kwargs(self):
        return dict(burn=BURN, covUpdate=COV_UPDATE, Tskip=TSKIP, thin=THIN, isave=E2E_ITERS, SCAMweight=self.weights[0],
                    AMweight=self.weights[1], DEweight=self.weights[2])


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


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_port_rate(walkers, iters, threads, seed=3):
    """Time the CPU oracle (plain-C port of the reference algorithm) on a bounded sample of the
    C2 workload: `walkers` x 32 temperatures x `iters` iterations."""
    from oracle import oracle as orc

    wl = Workload("C2")
    D, T = wl.d, wl.T
    o = orc.Oracle(D, walkers, T, wl.cov0, seed=seed, ladder=wl.ladder,
                   cycle=((orc.JUMP_SCAM, wl.weights[0]), (orc.JUMP_AM, wl.weights[1])), de_weight=wl.weights[2],
                   cov_update=COV_UPDATE, burn=BURN, tskip=TSKIP, thin=THIN,
                   logl_params=orc.gaussian_params(wl.mu, wl.icov),
                   logp_params=orc.uniform_params(-50 * np.ones(D), 60 * np.ones(D)),
                   max_rows=(2 * iters) // THIN + 2, nthreads=threads)
    o.set_state(wl.x0(1, T, walkers))
    o.run(BURN + 1)          # DE joins the cycle, covariance adapted once: the steady-state mix
    t0 = time.perf_counter()
    o.run(iters)
    dt = time.perf_counter() - t0
    return walkers * T * iters / dt, dt


def reference_python_rates(threads):
    """The unmodified reference on this host (baseline/_ref, scripts/install_reference.sh): config 1 verbatim on
    one core, and one process per core on the C2 target.  None when the install is absent."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_worker
    except Exception as exc:  # pragma: no cover
        return {"unavailable": "baseline/ref_worker.py failed to import: %s" % exc}
    if not ref_worker.available():
        return {"unavailable": "baseline/_ref is absent (run scripts/install_reference.sh where /root/reference exists); "
                               "the CPU arm is the C port alone"}
    c1_rate, c1_dt = ref_worker.run_config1()
    niter = 30000
    all_rate, all_wall = ref_worker.run_c2_all_cores(threads, niter)
    return {"config1": {"value": c1_rate, "unit": UNIT, "cores": 1, "seconds": c1_dt,
                        "workload": "examples/simple.py verbatim: 20-dim Gaussian in a [0,10] box, 1 chain, 10000 iterations, "
                                    "burn=covUpdate=500, thin=1"},
            "value": all_rate, "unit": UNIT, "cores": threads, "seconds": all_wall,
            "workload": "C2 target, one reference process per core, %d iterations each, thin=10, isave=Niter; independent "
                        "chains (mpi4py is not in the image, so the reference's PTswap cannot run)" % niter,
            "per_core": all_rate / threads}


def run_reference(args, rank, world):
    """Reference arm: the path's CPU implementation on all host cores, same config and metric.  The ratio uses the
    plain-C restatement of the reference algorithm (oracle/, kind "port": OpenMP over walkers, the strongest CPU
    baseline available); the unmodified Python reference is timed beside it (`reference_python`)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    walkers, iters = 128 * threads, 1000
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(walkers, 100, threads)
    times = []
    for _ in range(args.steps):
        r, dt = cpu_port_rate(walkers, iters, threads)
        times.append(dt)
    T = 32
    value = float(walkers * T * iters * len(times) / sum(times))
    sample = "%d walkers x %d temps x %d iterations per step, OpenMP over walkers" % (walkers, T, iters)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 20-dim Gaussian, SCAM/AM/DE 20/20/20, covUpdate=burn=1000, Tskip=100, thin=10; "
                               "CPU sample " + sample,
                   "cpu_sample_walkers": walkers,
                   "note": "per-chain-step rate of a bounded sample (the full 8192 walkers would take minutes); the C port, "
                           "not the Python reference, is the denominator of the driver's ratio"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ladder_steps_per_s": value / T,
        "reference_python": reference_python_rates(threads),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ engine arm
def peaks(local_rank):
    from ptmcmcsampler_b200 import _cabi

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    fp64 = _cabi.measure_fp64_peak(local_rank)
    return hbm, src, fp64


def traffic_of(kernel_name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json); null when the
    capture is of another kernel."""
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(prof))
        for rec in t.get("kernels", []):
            if rec["kernel"].split("(")[0].strip() == kernel_name.split("(")[0].strip():
                return rec["dram_bytes_per_launch"], rec.get("source")
    except Exception:
        pass
    return None, None


def device_run(wl, args_steps, args_warmup, rank, world, local_rank, group=None, ladder_mode=False):
    """Device-timed K steps of one workload with the state resident in HBM, then the same K steps with every
    launch bracketed by CUDA events (per-class durations).  Returns a dict of raw measurements."""
    import torch
    import torch.distributed as dist

    from ptmcmcsampler_b200 import _cabi, distributed

    d, W, T = wl.d, wl.W, wl.T
    kw = wl.engine_kwargs()
    total_iters = ITERS * (2 * args_steps + args_warmup) + BURN + 100
    ladder = wl.ladder
    shard_kw, Tg = dict(ntemps=T, ladder=ladder, walker_offset=rank * W), T
    if ladder_mode:  # BASELINE config 5: T rungs per GPU of one ladder of T * N rungs, same walkers everywhere
        Tg = T * world
        ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(Tg), 1e30)
        shard_kw = distributed.ladder_shard_kwargs(ladder, world, rank)
    eng = _cabi.Engine(d, W, shard_kw.pop("ntemps"), wl.cov0, shard_kw.pop("ladder"), seed=42,
                       record_rows=total_iters // THIN + 2, device=local_rank, timing=False, **kw, **shard_kw)
    eng.set_state(wl.x0(1 + rank, T, W))
    comm = distributed.LadderComm(eng) if ladder_mode else None

    def step(n=ITERS):
        if ladder_mode:
            distributed.run_ladder(eng, n, comm, TSKIP)
        elif world > 1:
            distributed.run(eng, n, group)
        else:
            eng.run(n)

    for _ in range(args_warmup):
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
    for _ in range(args_steps):
        step()
    ev1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    if comm is not None:
        comm.check(eng)  # the peer-memory swap exchange: no message timed out
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = int(sum(eng.timing()["launches"].values()))
    # same K steps again with every launch bracketed by CUDA events on the engine's stream: the
    # per-kernel-class durations behind `roofline` (the bracketing serialises host and device, so this
    # pass is not the one `value` is taken from)
    eng.reset_timing()
    eng.set_timing(True)
    for _ in range(args_steps):
        step()
    eng.sync()
    tm = eng.timing()
    eng.set_timing(False)
    kernel = eng.mh_kernel_name
    eng.close()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return dict(ms=ms, steps=args_steps, launches=launches, tm=tm, kernel=kernel, wall=(t_wall0, t_wall1), Tg=Tg,
                p2p=bool(comm is not None and comm.p2p),
                value=world * W * T * ITERS * args_steps / (ms * 1e-3))


def roofline_of(wl, run, hbm_peak, hbm_src, fp64_peak):
    """Roofline of the dominant kernel (the fused MH segment): algorithmic bytes and flops per launch over the
    kernel's mean launch duration (CUDA events around every launch, measured in this run)."""
    tm = run["tm"]
    n_launch, mh_ms = max(1, tm["launches"]["mh"]), tm["ms"]["mh"]
    steps_per_launch = wl.W * wl.T * ITERS * run["steps"] / n_launch
    sec = mh_ms / n_launch * 1e-3
    gbs = wl.bytes_per_step * steps_per_launch / sec / 1e9
    tfs = wl.flops_per_step * steps_per_launch / sec / 1e12
    traffic, traffic_src = traffic_of(run["kernel"])
    total_ms = sum(tm["ms"].values())
    common = {"kernel": run["kernel"], "launches": n_launch, "avg_launch_ms": mh_ms / n_launch,
              "kernel_share_of_step": mh_ms / total_ms if total_ms > 0 else None, "class_ms": tm["ms"],
              "algorithmic_bytes_per_chain_step": wl.bytes_per_step, "algorithmic_flops_per_chain_step": wl.flops_per_step,
              "traffic": traffic, "traffic_source": traffic_src}
    hbm = {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}
    fp64 = {"achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfs / fp64_peak,
            "peak_source": "measured in this run (ptmcmc_measure_fp64_peak: DFMA kernel, 8 chains x 512 threads x 2 blocks per SM)"}
    # the binding bound is the one with the larger fraction of its peak (HBM for C2 / C4, fp64 for C3)
    if fp64["frac"] > hbm["frac"]:
        out = dict(bound="tensor", **fp64)
        out["note"] = ("fp64 is the binding bound (dense quadratic form and AM mat-vec run as fp64 DMMA on the tensor pipe; "
                       "tcgen05 has no fp64 kind)")
        out["hbm"] = hbm
    else:
        out = dict(bound="hbm", **hbm)
        out["note"] = ("algorithmic stream-in/stream-out bytes per chain-step over the launch time; a launch keeps chain state "
                       "on chip for Tskip iterations, so real DRAM traffic (`traffic`) is far below the algorithmic bytes")
        out["fp64"] = fp64
    out.update(common)
    return out


def e2e_run(wl, rank, world, local_rank, nrep, ladder_mode=False, Tg=None):
    """End to end through the public API: host p0 in, recorded chain out, every step."""
    import torch
    import torch.distributed as dist

    from ptmcmcsampler_b200 import PTMCMCSampler, _cabi

    d, W, T = wl.d, wl.W, wl.T
    Tg = Tg or T
    ladder = np.minimum((1 + np.sqrt(2.0 / d)) ** np.arange(Tg), 1e30) if ladder_mode else None
    lk, pr = wl.targets()
    p0 = _cabi.pinned_empty((T, W, d))
    p0[...] = wl.x0(7 + rank, T, W)
    # chain / jump files go to memory-backed storage when there is one (a shared box's overlay file system stalls for
    # tens of ms now and then, which would be timed as if it were the sampler)
    outdir = tempfile.mkdtemp(prefix="ptmcmc_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    debug = bool(os.environ.get("PTMCMC_E2E_DEBUG"))

    def once(seed):
        t_a = time.perf_counter()
        s = PTMCMCSampler.PTSampler(d, lk, pr, wl.cov0.copy(), outDir=outdir, verbose=False, seed=seed, ntemps=Tg,
                                    nwalkers=W, device=local_rank, walker_offset=0 if ladder_mode else rank * W,
                                    dist_group=True if world > 1 else None, shard="ladder" if ladder_mode else "walkers")
        t_b = time.perf_counter()
        s.sample(p0, E2E_ITERS, ladder=ladder, **wl.sample_kwargs())
        t_c = time.perf_counter()
        loss = float(s._lnlike_all[-1].mean())   # the step's result read on the host
        d2h = s._chain_all.nbytes + s._lnlike_all.nbytes + s._lnprob_all.nbytes
        s.close()
        if debug:
            sys.stderr.write("e2e rank %d: construct %.1f ms, sample %.1f ms, close %.1f ms\n" % (
                rank, 1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (time.perf_counter() - t_c)))
        return loss, d2h

    once(100)
    once(99)   # (two warm-up calls: page-locked result arrays and device memory then come from the pools)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(nrep):
        _, d2h = once(101 + r)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    import shutil

    shutil.rmtree(outdir, ignore_errors=True)
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"value": world * W * T * E2E_ITERS * nrep / dt, "unit": UNIT, "h2d_bytes_per_step": int(p0.nbytes),
            "d2h_bytes_per_step": int(d2h), "steps": nrep, "iterations_per_step": E2E_ITERS,
            "note": "one step = a fresh PTSampler(...).sample(p0_host, 2000): engine build, H2D of p0, 2000 iterations (DE joins "
                    "the cycle at burn + 1 = 1001, so the second half runs the steady-state mix), D2H of the thinned T=1 "
                    "record of all walkers streamed while the engine runs, covariance / counters snapshot, chain file of "
                    "walker 0"}


def side_config(name, local_rank, hbm_peak, hbm_src, fp64_peak, steps=3, warmup=2):
    """One of the other BASELINE configurations on this GPU (bounded: `steps` timed steps)."""
    wl = Workload(name)
    run = device_run(wl, steps, warmup, 0, 1, local_rank)
    out = {"workload": wl.desc + ", covUpdate=burn=1000, Tskip=100, thin=10; step = 1000 MH iterations of all chains",
           "value": run["value"], "unit": UNIT, "ms_per_step": run["ms"] / steps, "steps": steps, "warmup": warmup,
           "ladder_steps_per_s": run["value"] / wl.T, "gpu_launches": run["launches"],
           "roofline": roofline_of(wl, run, hbm_peak, hbm_src, fp64_peak)}
    out["e2e"] = e2e_run(wl, 0, 1, local_rank, 2)
    return out


GAUSS_SOURCE = """
// par = mu[ndim], icov[ndim * ndim] (row-major): the C2 likelihood written the way a user would write it
double user_logl(const double *x, int ndim, const double *par) {
    const double *mu = par, *icov = par + ndim;
    double acc = 0.0;
    for (int i = 0; i < ndim; ++i) {
        double row = 0.0;
        for (int j = 0; j < ndim; ++j) row += icov[i * ndim + j] * (x[j] - mu[j]);
        acc += (x[i] - mu[i]) * row;
    }
    return -0.5 * acc;
}
"""


def user_target_run(local_rank):
    """The C2 workload with the likelihood supplied as CUDA source (compiled with NVRTC into the thread-per-chain
    kernel): what an arbitrary user likelihood costs on the device."""
    from ptmcmcsampler_b200 import _cabi

    wl = Workload("C2")
    kw = wl.engine_kwargs()
    kw.pop("logl_params")
    kw["logl_kind"] = _cabi.LOGL_USER
    t0 = time.perf_counter()
    eng = _cabi.Engine(wl.d, wl.W, wl.T, wl.cov0, wl.ladder, seed=42, record_rows=(BURN + 100 + 3 * ITERS) // THIN + 2,
                       device=local_rank, logl_source=GAUSS_SOURCE,
                       logl_user_params=np.concatenate([wl.mu, wl.icov.ravel()]), **kw)
    t_create = time.perf_counter() - t0
    eng.set_state(wl.x0(1, wl.T, wl.W))
    eng.run(BURN + 100)
    eng.sync()
    t0 = time.perf_counter()
    eng.run(2 * ITERS)
    eng.sync()
    dt = time.perf_counter() - t0
    name = eng.mh_kernel_name
    eng.close()
    return {"workload": wl.desc + "; log-likelihood given as CUDA source (SourceLikelihood), box prior built in",
            "value": wl.W * wl.T * 2 * ITERS / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / 2, "kernel": name,
            "create_s_including_nvrtc": t_create}


def host_callback_run(local_rank):
    """Python callables in the loop (the reference's own calling convention, vectorised over chains): one host round trip
    per iteration through the engine's page-locked buffers."""
    from ptmcmcsampler_b200 import PTMCMCSampler

    wl = Workload("C2")
    d, W, T, n = wl.d, 1024, 8, 300
    mu, icov = wl.mu, wl.icov

    def logl(X):
        D = X - mu
        return -0.5 * np.einsum("ni,ij,nj->n", D, icov, D)

    def logp(X):
        return np.where(np.all((X >= -50.0) & (X <= 60.0), axis=1), 0.0, -np.inf)

    outdir = tempfile.mkdtemp(prefix="ptmcmc_bench_cb_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    s = PTMCMCSampler.PTSampler(d, logl, logp, wl.cov0.copy(), outDir=outdir, verbose=False, seed=5, ntemps=T, nwalkers=W,
                                device=local_rank, vectorized=True)
    p0 = wl.x0(3, T, W)
    t0 = time.perf_counter()
    s.sample(p0, n, burn=100, covUpdate=100, Tskip=10, thin=10, isave=n)
    dt = time.perf_counter() - t0
    s.close()
    import shutil

    shutil.rmtree(outdir, ignore_errors=True)
    return {"workload": "C2 target as vectorised numpy callables, %d walkers x %d temps, %d iterations" % (W, T, n),
            "value": W * T * n / dt, "unit": UNIT, "ms_per_iteration": 1e3 * dt / n}


def run_engine(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = Workload("C2")
    ladder_mode = args.shard == "ladder" and world > 1
    hbm_peak, hbm_src, fp64_peak = peaks(local_rank)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    run = device_run(wl, args.steps, args.warmup, rank, world, local_rank, ladder_mode=ladder_mode)
    clk = clocks.stop(*run["wall"]) if rank == 0 else None
    roofline = roofline_of(wl, run, hbm_peak, hbm_src, fp64_peak)
    e2e = e2e_run(wl, rank, world, local_rank, max(1, min(args.steps, 3)), ladder_mode=ladder_mode, Tg=run["Tg"])

    extra = {}
    if world == 1 and not args.no_side:
        extra["configs"] = {name: side_config(name, local_rank, hbm_peak, hbm_src, fp64_peak) for name in ("C3", "C4")}
        extra["user_target"] = user_target_run(local_rank)
        extra["host_callback"] = host_callback_run(local_rank)
    if world > 1 and not ladder_mode and not args.no_side:
        # BASELINE config 5: the same GPUs as ONE ladder of 32 N rungs, 32 per GPU, nearest-neighbour swap exchange
        k = max(2, min(args.steps, 5))
        try:  # (a failure of this leg is collective, see LadderComm.check, and must not cost the headline line)
            r5 = device_run(wl, k, 2, rank, world, local_rank, ladder_mode=True)
            extra["c5_ladder"] = {
                "workload": "C5: %d-rung ladder as 32 rungs per GPU x %d GPUs, 8192 walkers, 20-dim Gaussian, nearest-neighbour "
                            "exchange of the boundary rung at every swap; factor and AM ring broadcast from the T=1 shard (NCCL)" % (run["Tg"] * world, world),
                "swap_exchange": "peer memory: the swap kernels store into the neighbour's mailbox over NVLink and spin on a flag"
                                 if r5["p2p"] else "NCCL send / receive",
                "value": r5["value"], "unit": UNIT, "ms_per_step": r5["ms"] / k, "steps": k, "warmup": 2,
                "fraction_of_walker_sharded": r5["value"] / run["value"], "gpu_launches": r5["launches"],
                "roofline": roofline_of(wl, r5, hbm_peak, hbm_src, fp64_peak),
                "e2e": e2e_run(wl, rank, world, local_rank, 2, ladder_mode=True, Tg=r5["Tg"])}
        except Exception as exc:  # noqa: BLE001 -- whatever it is, the headline line is still printed
            extra["c5_ladder"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        walkers = 128 * threads
        rate, dtc = cpu_port_rate(walkers, 1000, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "cpu_sample_walkers": walkers,
               "sample": "%d walkers x %d temps x 1000 iterations (%.1f s)" % (walkers, wl.T, dtc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": run["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": run["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "dtype_note": "state, proposals, likelihood and the Hastings test are f64; the Box-Muller normals of the AM / SCAM "
                          "proposals are generated in f32 with individually rounded operations (24-bit mantissa, |z| <= 6.76, "
                          "bit-identical on the CPU oracle; KS / tail tests in tests/test_oracle_golden.py)",
            "data": "synthetic",
            "config": {"workload": wl.desc + ", covUpdate=burn=1000, Tskip=100, thin=10; step = 1000 MH iterations of all chains",
                       "l2": "no flush needed: each step streams the 1.3 GB AM ring and gathers from the 1.3 GB DE "
                             "history (inputs >> 126 MB L2)",
                       "parallelism": ("ladder-sharded x%d (%d rungs, neighbour exchange of the boundary rung)" % (world, run["Tg"])
                                       if ladder_mode else "walker-sharded x%d" % world)},
            "ladder_steps_per_s": run["value"] / wl.T,
            "clocks": clk, "e2e": e2e, "gpu_launches": run["launches"], "roofline": roofline, "cpu_baseline": cpu,
        }
        line.update(extra)
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
                         "32 N rungs split over the GPUs (BASELINE config 5; also measured as `c5_ladder` in the default run)")
    ap.add_argument("--no-side", action="store_true", help="skip the C3 / C4 / C5 side measurements")
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
