#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--traj-only]

Trajectory fixtures (``traj_*.npz``) run the reference with its ``stream`` replaced by the
oracle's counter-based draws (oracle/ref_harness.py), so every jump choice, accept flag, swap
and state is a known answer for the oracle.  Statistical fixtures (``stats_*.npz``) run the
reference with its own NumPy PCG64 stream and store posterior moments / acceptance rates with
Monte-Carlo errors from independent repeats; they define the 3-sigma band for the engine.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh  # noqa: E402


def gaussian_problem(d, seed, pmin, pmax, mu_lo=2.0, mu_hi=8.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T + 0.5 * np.eye(d)
    mu = rng.uniform(mu_lo, mu_hi, d)
    return rh.GaussianProblem(mu, cov, pmin, pmax)


def save(name, **kw):
    kw = {k: v for k, v in kw.items() if not k.startswith("_")}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


def traj_case(name, d, T, N, seed, sample_kwargs, groups=None, curved=False, pmin=-50.0, pmax=60.0,
              ext=False, prior_weight=0):
    rng = np.random.default_rng(seed + 1000)
    if curved:
        pb = rh.CurvedProblem(d)
        meta = dict(kind="curved", pb_lo=pb.pmin, pb_hi=pb.pmax, inclusive=0)
        p0 = rng.uniform(-1.0, 1.0, (T, d))
        cov0 = np.diag(0.5 * (1.0 + 0.1 * np.arange(d)))
    else:
        pb = gaussian_problem(d, seed, pmin, pmax)
        meta = dict(kind="gaussian", pb_mu=pb.mu, pb_icov=pb.icov, pb_lo=pb.a, pb_hi=pb.b, inclusive=1)
        p0 = rng.uniform(0.0, 10.0, (T, d))
        cov0 = np.diag(0.01 * (1.0 + np.arange(d)))
    ext_jumps = []
    if ext:
        lo, hi = (pb.pmin, pb.pmax) if curved else (pb.a, pb.b)

        def golden_ext_jump(x, it, beta):
            # deterministic stand-in for the reference's UniformJump (tests/test_simple.py:44-62)
            frac = np.modf(np.abs(np.sin(np.arange(1, len(x) + 1) * (it + 1.0) + 13.0 * x)) * 1e4)[0]
            return lo + (hi - lo) * (0.45 + 0.1 * frac), 0.05 * np.sin(it) * beta

        ext_jumps = [(golden_ext_jump, 7)]
    if prior_weight:
        lo, hi = (pb.pmin, pb.pmax) if curved else (pb.a, pb.b)

        def bind(sampler):
            # the reference's UniformJump (tests/test_simple.py:44-62) drawing from the sampler's own stream,
            # one random() per parameter: what the engine's built-in prior-draw jump restates
            def priorDrawJump(x, it, beta):
                return lo + (hi - lo) * np.array([sampler.stream.random() for _ in range(len(x))]), 0.0
            return priorDrawJump

        bind.bind_sampler = True
        bind.jump_name = "priorDrawJump"
        ext_jumps = ext_jumps + [(bind, prior_weight, rh.orc.JUMP_PRIOR)]
    r = rh.run_reference(d, pb.lnlikefn, pb.lnpriorfn, cov0, p0, N, seed=seed, ntemps=T,
                         sample_kwargs=sample_kwargs, groups=groups, ext_jumps=ext_jumps)
    gflat = np.concatenate(groups) if groups is not None else np.arange(d)
    goff = np.cumsum([0] + [len(g) for g in groups]) if groups is not None else np.array([0, d])
    save(name, d=d, T=T, N=N, seed=seed, cov0=cov0, p0=p0, group_offsets=goff, group_indices=gflat,
         has_groups=int(groups is not None), ext=int(ext), prior_weight=int(prior_weight),
         **{"kw_" + k: v for k, v in sample_kwargs.items()}, **meta, **r, **output_files(r["_outdir"]))


def output_files(outdir):
    """The files the reference wrote (chain_*.txt, jumps.txt, *_jump.txt as bytes; cov.npy as an array), keyed
    ``file_<name with . -> _>`` plus the list of names."""
    files = rh.read_output_files(outdir)
    out = {"file_names": np.array(sorted(files))}
    for name, val in files.items():
        out["file_" + name.replace(".", "_")] = val if isinstance(val, np.ndarray) else np.frombuffer(val, dtype=np.uint8)
    return out


def resume_case(name, d, N, seed, sample_kwargs):
    """Reference run of N iterations, then a second sampler with resume=True on the same outDir to 2N (ref :290-319,
    :474-476, :591-599), both under the shim stream.  The proposal cycle is a deterministic plugin jump plus DE only
    (SCAMweight = AMweight = 0), so the trajectory does not depend on LAPACK's eigenvector signs."""
    import shutil

    pb = gaussian_problem(d, seed, 0.0, 10.0)
    rng = np.random.default_rng(seed + 1000)
    p0 = rng.uniform(0.0, 10.0, (1, d))
    cov0 = np.diag(0.01 * (1.0 + np.arange(d)))

    def golden_ext_jump(x, it, beta):
        frac = np.modf(np.abs(np.sin(np.arange(1, len(x) + 1) * (it + 1.0) + 13.0 * x)) * 1e4)[0]
        return pb.a + (pb.b - pb.a) * (0.45 + 0.1 * frac), 0.05 * np.sin(it) * beta

    ext = [(golden_ext_jump, 7)]
    first = rh.run_reference(d, pb.lnlikefn, pb.lnpriorfn, cov0, p0, N, seed=seed, sample_kwargs=sample_kwargs, ext_jumps=ext)
    files_first = output_files(first["_outdir"])
    second = rh.run_reference(d, pb.lnlikefn, pb.lnpriorfn, cov0, p0, 2 * N, seed=seed, sample_kwargs=sample_kwargs,
                              ext_jumps=ext, outdir=first["_outdir"], resume=True)
    s2 = second["_samplers"][0]
    save(name, d=d, T=1, N=N, seed=seed, cov0=cov0, p0=p0, ext=1, kind="gaussian", pb_mu=pb.mu, pb_icov=pb.icov, pb_lo=pb.a,
         pb_hi=pb.b, inclusive=1, **{"kw_" + k: v for k, v in sample_kwargs.items()},
         **{"first_" + k: v for k, v in files_first.items()},
         **{"second_" + k: v for k, v in output_files(first["_outdir"]).items()},
         resume_length=int(s2.resumeLength), naccepted=float(s2.naccepted), jump=second["jump"], acc=second["acc"],
         x=second["x"], lnl=second["lnl"], chain=second["chain"], chain_lnl=second["chain_lnl"],
         chain_lnp=second["chain_lnp"], am=second["am"], de=second["de"], cov=second["cov"], mu=second["mu"],
         m2=second["m2"], jump_prop=second["jump_prop"], jump_acc=second["jump_acc"], ladder=second["ladder"])
    shutil.rmtree(first["_outdir"], ignore_errors=True)


def stats_case(name, d, T, N, nrep, sample_kwargs, pmin, pmax, seed0, burn_frac=0.25):
    """Reference with its own RNG; moments of the post-burn-in T=1 chain (and hot chains)."""
    pb = gaussian_problem(d, 77, pmin, pmax)
    means, vars_, accs, swaps, jacc = [], [], [], [], []
    t0 = time.time()
    for rep in range(nrep):
        rng = np.random.default_rng(seed0 + rep)
        p0 = rng.uniform(max(pmin, 0.0), min(pmax, 10.0), (T, d))
        r = rh.run_reference(d, pb.lnlikefn, pb.lnpriorfn, np.eye(d) * 0.01, p0, N, seed=seed0 + rep,
                             ntemps=T, shim=False, sample_kwargs=sample_kwargs)
        x = r["x"][int(N * burn_frac):]                       # [n][T][d]
        means.append(x.mean(axis=0))
        vars_.append(x.var(axis=0))
        accs.append(r["naccepted"] / float(N))
        swaps.append(r["swap_acc"][-1] / max(1.0, float(r["swap_proposed"])))
        jacc.append(r["jump_acc"] / np.maximum(1, r["jump_prop"]))
    print("%s: %d reps in %.1f s" % (name, nrep, time.time() - t0))
    save(name, d=d, T=T, N=N, nrep=nrep, burn_frac=burn_frac, pb_mu=pb.mu, pb_icov=pb.icov, pb_cov=pb.cov,
         pb_lo=pb.a, pb_hi=pb.b, ladder=r["ladder"], means=np.array(means), vars=np.array(vars_),
         acc=np.array(accs), swap=np.array(swaps), jump_acc=np.array(jacc),
         **{"kw_" + k: v for k, v in sample_kwargs.items()})


def main():
    kw1 = dict(burn=200, thin=1, covUpdate=100, SCAMweight=20, AMweight=20, DEweight=20, isave=1000, Tskip=10)
    traj_case("traj_t1_d5", 5, 1, 700, 1234, kw1)
    kw2 = dict(burn=200, thin=5, covUpdate=50, SCAMweight=30, AMweight=15, DEweight=50, isave=1000, Tskip=7)
    traj_case("traj_t4_groups_d6", 6, 4, 450, 99, kw2, pmin=0.0, pmax=10.0,
              groups=[np.array([0, 1, 2, 3, 4, 5]), np.array([1, 3]), np.array([5, 0, 2])])
    kw3 = dict(burn=150, thin=3, covUpdate=50, SCAMweight=10, AMweight=10, DEweight=60, isave=3000, Tskip=5)
    traj_case("traj_t3_curved_ext_d4", 4, 3, 400, 4242, kw3, curved=True, ext=True)
    kw4 = dict(burn=100, thin=10, covUpdate=100, SCAMweight=20, AMweight=20, DEweight=20, isave=1000, Tskip=100)
    traj_case("traj_t1_d20", 20, 1, 350, 7, kw4, pmin=0.0, pmax=10.0)
    # hotChain=True with an explicit Tmax: the last rung samples at temp = 1e80 (ref :281-282) but swaps with
    # ladder[-1] (ref :658, :673-676); the ladder comes from Tmax (ref :709-718)
    kw5 = dict(burn=100, thin=2, covUpdate=50, SCAMweight=20, AMweight=20, DEweight=30, isave=1000, Tskip=5,
               Tmax=20.0, hotChain=True)
    traj_case("traj_t3_hot_tmax_d4", 4, 3, 300, 31, kw5, pmin=0.0, pmax=10.0)
    # the device-side prior-draw jump against the reference running the equivalent plugin
    kw6 = dict(burn=100, thin=1, covUpdate=50, SCAMweight=15, AMweight=15, DEweight=20, isave=1000, Tskip=10)
    traj_case("traj_t2_prior_d4", 4, 2, 300, 77, kw6, pmin=0.0, pmax=10.0, prior_weight=10)
    # reference-style resume: 300 iterations, then resume=True on the same directory to 600
    kw7 = dict(burn=100, thin=2, covUpdate=50, SCAMweight=0, AMweight=0, DEweight=40, isave=100, Tskip=100)
    resume_case("resume_t1_d4", 4, 300, 55, kw7)
    if "--traj-only" in sys.argv:  # the statistical bands do not depend on the oracle's draw functions
        return
    # statistical bands (reference's own PCG64 stream)
    kws = dict(burn=1000, thin=1, covUpdate=1000, SCAMweight=20, AMweight=20, DEweight=20, isave=10**9, Tskip=10)
    stats_case("stats_t1_d8", 8, 1, 40000, 12, kws, -50.0, 60.0, 500)
    stats_case("stats_t1_d8_box", 8, 1, 40000, 12, kws, 3.0, 7.0, 600)
    stats_case("stats_t4_d8", 8, 4, 12000, 8, kws, -50.0, 60.0, 700)


if __name__ == "__main__":
    main()
