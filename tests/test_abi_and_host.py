"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares (no compute
calls without a GPU), and the host-side mirror of the reference API behaves like the reference
where no device is involved."""
import ctypes
import os
import re

import numpy as np
import pytest

from ptmcmcsampler_b200 import PTMCMCSampler, _cabi, nompi4py
from ptmcmcsampler_b200.likelihoods import CurvedLikelihood, GaussianLikelihood, UniformPrior

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptmcmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptmcmc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_cabi.SYMBOLS) == names
    assert lib.ptmcmc_abi_version() == _cabi.ABI_VERSION


def test_config_struct_layout_matches_header():
    # the header's struct, compiled by gcc, must have the size ctypes computes
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "ptmcmc_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(ptmcmc_config), sizeof(ptmcmc_timing));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "sz.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "sz")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b = subprocess.check_output([exe]).split()
    assert int(a) == ctypes.sizeof(_cabi.Config)
    assert int(b) == ctypes.sizeof(_cabi.Timing)


def test_engine_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.EngineError) as ei:
        _cabi.Engine(3, 2, 2, np.eye(3), np.array([1.0, 2.0]), logl_params=np.zeros(13), logp_params=np.zeros(8))
    assert "no CUDA device" in str(ei.value)


def test_sampler_host_surface(tmp_path):
    out = str(tmp_path / "chains")
    lk, pr = GaussianLikelihood(np.zeros(4), cov=np.eye(4)), UniformPrior(-5, 5)
    s = PTMCMCSampler.PTSampler(4, lk, pr, np.eye(4), outDir=out, verbose=False, seed=3, ntemps=5)
    assert os.path.isdir(out) and s.nchain == 5 and s.MPIrank == 0
    # ladder formula, ref :709-718
    assert np.allclose(s.temperatureLadder(1), (1 + np.sqrt(2 / 4.0)) ** np.arange(5))
    assert np.allclose(s.temperatureLadder(2.0, Tmax=32.0), 2.0 * 2.0 ** np.arange(5))
    # plugin registry, ref :988-1014
    def myjump(x, it, beta):
        return x, 0.0
    s.addProposalToCycle(myjump, 0)
    assert s.propCycle == [] and "myjump" not in s.jumpDict
    s.addProposalToCycle(myjump, 3)
    s.addProposalToCycle(s.covarianceJumpProposalSCAM, 2)
    assert len(s.propCycle) == 5 and s.jumpDict["myjump"] == [0, 0]
    assert os.path.isfile(os.path.join(out, "myjump_jump.txt"))
    assert s._cycle_segments() == [(_cabi.JUMP_EXT0, 3), (_cabi.JUMP_SCAM, 2)]
    s.addAuxilaryJump(lambda x, q, it, beta: (q, 0.0))
    assert len(s.aux) == 1 and s._external
    one = PTMCMCSampler.PTSampler(4, lk, pr, np.eye(4), outDir=out, verbose=False)
    assert np.array_equal(one.temperatureLadder(1), np.array([1]))


def test_sampler_rejects_bad_arguments_before_touching_the_device(tmp_path):
    out = str(tmp_path / "chains")
    lk, pr = GaussianLikelihood(np.zeros(3), cov=np.eye(3)), UniformPrior(-5, 5)
    s = PTMCMCSampler.PTSampler(3, lk, pr, np.eye(3), outDir=out, verbose=False)
    with pytest.raises(ValueError):
        s.sample(np.zeros(3), 100, isave=15, thin=10)
    with pytest.raises(ValueError):
        PTMCMCSampler.PTSampler(3, lk, pr, np.eye(4), outDir=out)

    class FakeWorld(nompi4py.MPIDummy):
        def Get_size(self):
            return 4

    with pytest.raises(NotImplementedError):
        PTMCMCSampler.PTSampler(3, lk, pr, np.eye(3), outDir=out, comm=FakeWorld())


def test_target_descriptors():
    lk = GaussianLikelihood(np.arange(3.0), cov=2 * np.eye(3), offset=1.5)
    p = lk.params(3)
    assert p.shape == (13,) and np.allclose(p[3:12].reshape(3, 3), 0.5 * np.eye(3)) and p[-1] == 1.5
    with pytest.raises(ValueError):
        lk.params(4)
    with pytest.raises(ValueError):
        GaussianLikelihood(np.zeros(2))
    pr = UniformPrior(-1.0, [1.0, 2.0], inclusive=False, value=-3.0)
    assert np.array_equal(pr.params(2), [-1, -1, 1, 2, -3, 0])
    assert CurvedLikelihood().kind == _cabi.LOGL_CURVED and CurvedLikelihood().params(4) is None


def test_shift_array_semantics():
    a = np.arange(12.0).reshape(6, 2)
    assert np.array_equal(PTMCMCSampler.shift_array(a, -2)[:4], a[2:])
    assert np.all(PTMCMCSampler.shift_array(a, -2)[4:] == 0)
    assert np.array_equal(PTMCMCSampler.shift_array(a, 2)[2:], a[:4])
    assert np.array_equal(PTMCMCSampler.shift_array(a, 0), a)


def test_bench_accounting_and_clock_parsing(tmp_path):
    """bench.py's roofline inputs are SURVEY section 8d's figures; the clock sampler keeps the samples
    stamped inside the timed window and reports throttle reasons."""
    import datetime
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c2, c3, c4 = bench.Workload("C2"), bench.Workload("C3"), bench.Workload("C4")
    assert abs(c2.bytes_per_step - 513.2667) < 1e-3 and abs(c3.bytes_per_step - 2291.2) < 0.5   # 513 B, 2 291 B
    assert abs(c4.bytes_per_step - 354.2) < 0.5                                                    # 354 B
    assert abs(c2.flops_per_step - (860 + 40 + 800 / 3 + 80 / 3 + 30)) < 1e-9
    assert 2.6e4 < c3.flops_per_step < 2.8e4                                                       # ~2.7e4
    assert c2.cov.shape == (20, 20) and np.all(np.linalg.eigvalsh(c2.cov) > 0) and len(c2.ladder) == 32
    assert (c3.d, c3.W, c3.T) == (100, 4096, 64) and (c4.d, c4.W, c4.T, c4.weights) == (10, 16384, 128, (10, 10, 60))
    assert set(c3.engine_kwargs()) >= {"cycle", "de_weight", "logl_params", "logp_params"}
    cs = bench.ClockSampler(0)
    cs.proc = type("P", (), {"terminate": lambda s: None, "wait": lambda s, timeout=None: 0, "kill": lambda s: None})()
    cs.path = str(tmp_path / "clk.csv")
    t0 = datetime.datetime(2026, 1, 1, 12, 0, 0)
    rows = [(t0 + datetime.timedelta(milliseconds=50 * i), 1965 if i != 5 else 1200) for i in range(20)]
    with open(cs.path, "w") as fh:
        for ts, mhz in rows:
            fh.write("%s, %d, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, %s\n"
                     % (ts.strftime("%Y/%m/%d %H:%M:%S.%f")[:-3], mhz, "Active" if mhz == 1200 else "Not Active"))
    out = cs.stop(rows[4][0].timestamp(), rows[9][0].timestamp())
    assert out["window"] == "timed region" and 6 <= out["samples"] <= 8
    assert out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


def test_log_comparison_decides_like_the_exponential_outside_the_band():
    """The swap sweep on the device decides ``u <= exp(lar)`` (ref :679, the oracle's swap_accept) from ``lar - log(u)``
    unless the two agree to 1e-12, where it evaluates the exponential (csrc/swap_kernels.cuh swap_accept_prep).  The
    restated rule must agree with the reference's comparison everywhere, ties and non-finite values included."""
    rng = np.random.default_rng(5)
    u = rng.random(200000)
    lar = np.concatenate([rng.normal(0, 3, 100000), np.log(u[100000:]) * (1 + rng.normal(0, 1e-13, 100000))])
    u[::5000] = 0.0
    lar[1::5000] = -np.inf
    lar[2::5000] = np.inf
    lar[3::5000] = np.nan
    with np.errstate(all="ignore"):
        ref = u <= np.exp(lar)
        logu = np.log(u)
        gap, band = lar - logu, 1e-12 * np.maximum(1.0, np.abs(lar))
        rule = np.where(gap > band, True, np.where(gap < -band, False, ref))
    assert np.array_equal(rule, ref)
    assert ((np.abs(gap) <= band) & np.isfinite(gap)).sum() > 1000  # the band was exercised
