"""The on-disk outputs and the reference-style resume, pinned to files the UNMODIFIED reference wrote (tests/golden/
make_golden.py keeps the reference run's outDir in the fixtures): chain file layout and values, jumps.txt,
<jump>_jump.txt, cov.npy (ref writeOutput / _writeToFile :341-372, :722-766), and a run resumed from the reference's
own chain file (ref :290-319, :474-476, :591-599)."""
import os
import re

import numpy as np
import pytest

from ptmcmcsampler_b200 import PTMCMCSampler
from ptmcmcsampler_b200.likelihoods import CurvedLikelihood, GaussianLikelihood, UniformPrior

from _helpers import fixture_groups, golden_ext_jump_factory, load

pytestmark = pytest.mark.gpu

ROW = re.compile(r"^(-?\d+\.\d{22}\t)+(-?(\d+\.\d{6}|inf|nan)\t){3}-?\d+\.\d{6}\n$")


class InjectedFactors(PTMCMCSampler.PTSampler):
    """Stops at every covariance boundary and overwrites the engine's eigen-factor with the one LAPACK gave the
    reference (eigenvector signs are a convention; with them the W=1 engine walks the reference's path)."""

    factors = None

    def _advance(self, n, iter0):
        cu, done, eng = self.covUpdate, 0, self._engine
        while done < n:
            it = iter0 + done
            step = min(n - done, cu - it % cu)
            super()._advance(step, it)
            done += step
            k = (it + step) // cu - 1
            if (it + step) % cu == 0 and k < len(self.factors[0]):
                eng.adapt_finish(eng.adapt_begin())
                eng.set_factor(self.factors[0][k], self.factors[1][k])


def sampler_from_fixture(g, outdir, cls=PTMCMCSampler.PTSampler, **extra):
    d, T = int(g["d"]), int(g["T"])
    inclusive = bool(int(g["inclusive"]))
    if str(g["kind"]) == "gaussian":
        lk = GaussianLikelihood(g["pb_mu"], icov=g["pb_icov"])
    else:
        lk = CurvedLikelihood()
    pr = UniformPrior(g["pb_lo"], g["pb_hi"], inclusive=inclusive)
    s = cls(d, lk, pr, np.array(g["cov0"]), groups=fixture_groups(g), outDir=outdir, verbose=False, seed=int(g["seed"]),
            ntemps=T, nwalkers=1, **extra)
    if int(g["ext"]):
        s.addProposalToCycle(golden_ext_jump_factory(g["pb_lo"], g["pb_hi"]), 7)
    if "prior_weight" in g and int(g["prior_weight"]):
        s.addProposalToCycle(s.priorDrawJump, int(g["prior_weight"]))
    kw = {k[3:]: g[k].item() for k in g if k.startswith("kw_")}
    return s, kw


def check_chain_file(ours, theirs, d):
    lines_o, lines_t = ours.decode().splitlines(True), theirs.decode().splitlines(True)
    assert len(lines_o) == len(lines_t)
    for ln in lines_o:
        assert ROW.match(ln), ln
        assert ln.count("\t") == d + 3
    a = np.array([[float(v) for v in ln.split("\t")] for ln in lines_o])
    b = np.array([[float(v) for v in ln.split("\t")] for ln in lines_t])
    assert np.allclose(a[:, :d], b[:, :d], rtol=1e-9, atol=1e-9)
    assert np.allclose(a[:, d:d + 2], b[:, d:d + 2], rtol=0, atol=2e-6, equal_nan=True)   # printed with %f
    assert np.array_equal(a[:, d + 2:], b[:, d + 2:])                                      # rates of integer counters


@pytest.mark.parametrize("name", ["traj_t1_d5", "traj_t1_d20", "traj_t4_groups_d6", "traj_t2_prior_d4", "traj_t3_hot_tmax_d4"])
def test_output_files_match_the_reference_files(name, tmp_path):
    g = load(name)
    d, T, N = int(g["d"]), int(g["T"]), int(g["N"])
    out = str(tmp_path / "chains")
    s, kw = sampler_from_fixture(g, out, cls=InjectedFactors)
    s.factors = (g["U"], g["S"])
    s.sample(np.array(g["p0"])[:, None, :], N, **kw)
    names = [str(n) for n in g["file_names"]]
    cold = "chain_1.txt" if T == 1 else "chain_1.0.txt"
    assert cold in names
    check_chain_file(open(os.path.join(out, cold), "rb").read(), bytes(g["file_" + cold.replace(".", "_")]), d)
    for n in names:
        key = "file_" + n.replace(".", "_")
        if n == "jumps.txt":   # the reference iterates over a set of bound methods: the line order is arbitrary
            assert sorted(open(os.path.join(out, n), "rb").read().splitlines()) == sorted(bytes(g[key]).splitlines())
        elif n.endswith("_jump.txt"):
            assert open(os.path.join(out, n), "rb").read() == bytes(g[key]), n
        elif n == "cov.npy":
            assert np.allclose(np.load(os.path.join(out, n)), g[key], rtol=1e-8, atol=1e-12)
        elif n != cold:        # hot rungs: the reference creates no rows without writeHotChains
            assert len(g[key]) == 0 and (not os.path.exists(os.path.join(out, n)) or os.path.getsize(os.path.join(out, n)) == 0)


def test_resume_from_a_chain_file_the_reference_wrote(tmp_path):
    """The fixture holds a reference run of 300 iterations and its continuation with resume=True to 600.  Here the engine
    resumes from the REFERENCE's chain file (first half) and must arrive where the reference's resumed run arrived:
    replayed rows, every post-resume jump and accept flag, buffers, covariance, acceptance counter, files."""
    g = load("resume_t1_d4")
    d, N = int(g["d"]), int(g["N"])
    out = str(tmp_path / "chains")
    os.makedirs(out)
    for n in g["first_file_names"]:
        key = "first_file_" + str(n).replace(".", "_")
        if str(n).endswith(".txt"):
            open(os.path.join(out, str(n)), "wb").write(bytes(g[key]))
    s, kw = sampler_from_fixture(g, out, resume=True)
    s.sample(np.array(g["p0"])[:, None, :], 2 * N, **kw)
    assert s.resumeLength == int(g["resume_length"])
    check_chain_file(open(os.path.join(out, "chain_1.txt"), "rb").read(), bytes(g["second_file_chain_1_txt"]), d)
    assert np.allclose(s._chain, g["chain"], rtol=1e-9, atol=1e-9)
    assert np.allclose(s._lnlike, g["chain_lnl"], rtol=0, atol=2e-6) and np.allclose(s._lnprob, g["chain_lnp"], rtol=0, atol=2e-6)
    assert abs(float(s.naccepted) - float(g["naccepted"])) < 1e-6
    # counters: proposals of the second half only (the replay proposes nothing, ref :591-599)
    for jid, name in ((2, "DEJump"), (4, "golden_ext_jump")):
        assert s.jumpDict[name] == [int(g["jump_prop"][0, jid]), int(g["jump_acc"][0, jid])], name
    am, de = s._AMbuffer, s._DEbuffer
    assert np.allclose(am, g["am"], rtol=1e-9, atol=1e-9) and np.allclose(de, g["de"], rtol=1e-9, atol=1e-9)
    assert np.allclose(np.asarray(s.cov), g["cov"], rtol=1e-8, atol=1e-12)
    assert np.allclose(np.load(os.path.join(out, "cov.npy")), g["second_file_cov_npy"], rtol=1e-8, atol=1e-12)
    for n in ("DEJump_jump.txt", "golden_ext_jump_jump.txt"):
        assert open(os.path.join(out, n), "rb").read() == bytes(g["second_file_" + n.replace(".", "_")]), n
