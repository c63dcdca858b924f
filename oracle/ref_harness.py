"""Drive the UNMODIFIED reference (``/root/reference``) for golden-vector generation.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Runs only in the build container (the
reference is not present on the GPU box); its outputs are committed under
``tests/golden/`` by ``tests/golden/make_golden.py``.

Two things are substituted around the reference, neither touches its code:

* ``sampler.stream`` (a public attribute, ref ``PTMCMCSampler.py:99-105``) is
  replaced by :class:`ShimStream`, which answers ``integers/random/uniform/
  standard_normal/shuffle`` with the oracle's counter-based draws.  The
  reference then makes exactly the decisions the oracle must reproduce.
* ``comm`` is a thread-backed communicator implementing the eight methods the
  sampler calls so that ``PTswap`` and the covariance / DE
  ``send``/``recv`` exchange run in-process with one sampler per temperature.
"""
import importlib
import os
import queue
import sys
import tempfile
import threading
import types

import numpy as np

from . import oracle as orc

REF_ROOT = os.environ.get("PTMCMC_REFERENCE", "/root/reference")


def import_reference():
    """Import ``PTMCMCSampler.PTMCMCSampler`` straight from the read-only checkout.

    The package ``__init__`` imports a setuptools_scm-generated ``version.py`` that is
    absent from the checkout, so a stub package object is registered first.
    """
    name = "PTMCMCSampler"
    if name + ".PTMCMCSampler" in sys.modules:
        return sys.modules[name + ".PTMCMCSampler"]
    pkg_dir = os.path.join(REF_ROOT, "PTMCMCSampler")
    if not os.path.isdir(pkg_dir):
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    pkg = types.ModuleType(name)
    pkg.__path__ = [pkg_dir]
    sys.modules[name] = pkg
    ver = types.ModuleType(name + ".version")
    ver.version = "0+ref.dd837f9"
    sys.modules[name + ".version"] = ver
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # "Optional mpi4py package is not installed"
        mod = importlib.import_module(name + ".PTMCMCSampler")
    return mod


class ShimStream(object):
    """Duck-typed ``numpy.random.Generator`` backed by the oracle's Philox draws."""

    def __init__(self, seed, walker, temp):
        self.seed, self.walker, self.temp = seed, walker, temp
        self.iter = None
        self.j = 0
        self.jswap = 0

    def begin_step(self, it):
        self.iter, self.j, self.jswap = it, 0, 0

    def _word(self):
        w = orc.draw_word(self.seed, orc.PURPOSE_MH, self.iter, self.walker, self.temp, self.j)
        self.j += 1
        return w

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        if self.iter is None:
            raise RuntimeError("integers() outside a step")
        v = low + orc.word_to_int(self._word(), high - low)
        return np.array([v]) if size is not None else v

    def random(self):
        return orc.word_to_unit(self._word())

    def uniform(self):
        # only PTswap draws with uniform() (ref :679); those come from the swap stream of walker 0
        w = orc.draw_word(self.seed, orc.PURPOSE_SWAP, self.iter, self.walker, 0, self.jswap)
        self.jswap += 1
        return orc.word_to_unit(w)

    def standard_normal(self, size=None):
        if size is None:
            return orc.word_to_normals(self._word())[0]
        out = np.zeros(size)
        for j in range(0, size, 2):
            z0, z1 = orc.word_to_normals(self._word())
            out[j] = z0
            if j + 1 < size:
                out[j + 1] = z1
        return out

    def shuffle(self, arr):
        # randomizeProposalCycle's result is never read (ref :1045); consume nothing
        return None


class ThreadComm(object):
    """mpi4py-like communicator for ``size`` threads of one process."""

    class _Shared(object):
        def __init__(self, size):
            self.size = size
            self.barrier = threading.Barrier(size)
            self.slots = [None] * size
            self.bcast_slot = None
            self.queues = {}
            self.lock = threading.Lock()

    def __init__(self, shared, rank):
        self.sh, self.rank = shared, rank

    @classmethod
    def world(cls, size):
        sh = cls._Shared(size)
        return [cls(sh, r) for r in range(size)]

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.sh.size

    def barrier(self):
        self.sh.barrier.wait()

    def _q(self, src, dst, tag):
        with self.sh.lock:
            return self.sh.queues.setdefault((src, dst, tag), queue.Queue())

    def send(self, obj, dest=1, tag=55):
        self._q(self.rank, dest, tag).put(np.copy(obj) if isinstance(obj, np.ndarray) else obj)

    def recv(self, source=1, tag=55):
        return self._q(source, self.rank, tag).get()

    def gather(self, obj, root=0):
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        out = list(self.sh.slots) if self.rank == root else None
        self.sh.barrier.wait()
        return out

    def scatter(self, objs, root=0):
        if self.rank == root:
            self.sh.bcast_slot = objs
        self.sh.barrier.wait()
        out = self.sh.bcast_slot[self.rank]
        self.sh.barrier.wait()
        return out

    def bcast(self, obj, root=0):
        if self.rank == root:
            self.sh.bcast_slot = obj
        self.sh.barrier.wait()
        out = self.sh.bcast_slot
        self.sh.barrier.wait()
        return out


# ---- the fixture problems (formulas as in the reference's examples) -------

class GaussianProblem(object):
    """examples/simple.py:17-44 with explicit mean/covariance."""

    def __init__(self, mu, cov, pmin, pmax):
        self.mu = np.asarray(mu, dtype=float)
        self.cov = np.asarray(cov, dtype=float)
        self.icov = np.linalg.inv(self.cov)
        self.a = np.ones(len(self.mu)) * pmin
        self.b = np.ones(len(self.mu)) * pmax

    def lnlikefn(self, x):
        diff = x - self.mu
        return -np.dot(diff, np.dot(self.icov, diff)) / 2.0

    def lnpriorfn(self, x):
        if np.all(self.a <= x) and np.all(self.b >= x):
            return 0.0
        return -np.inf


class CurvedProblem(object):
    """examples/curved_likelihood.ipynb CurvedLikelihood, replicated over 2-D blocks."""

    def __init__(self, ndim=2):
        self.ndim = ndim
        self.pmin = -10.0 * np.ones(ndim)
        self.pmax = 10.0 * np.ones(ndim)

    def lnlikefn(self, x):
        tot = 0.0
        with np.errstate(divide="ignore"):
            for b in range(0, self.ndim, 2):
                a, y = x[b], x[b + 1]
                ll = np.exp(-a**2 - (9 + 4 * a**2 + 9 * y) ** 2) + 0.5 * np.exp(-8 * a**2 - 8 * (y - 2) ** 2)
                tot += np.log(ll)
        return tot

    def lnpriorfn(self, x):
        if np.all(self.pmin < x) and np.all(self.pmax > x):
            return 0.0
        return -np.inf


def read_output_files(outdir):
    """Everything the reference wrote into its outDir (ref writeOutput / _writeToFile :341-372, :722-766): text files
    as bytes, cov.npy as an array."""
    files = {}
    for name in sorted(os.listdir(outdir)):
        path = os.path.join(outdir, name)
        if name.endswith(".txt"):
            files[name] = open(path, "rb").read()
        elif name == "cov.npy":
            files[name] = np.load(path)
    return files


def run_reference(ndim, logl, logp, cov0, p0s, niter, seed, ntemps=1, shim=True, sample_kwargs=None,
                  groups=None, ext_jumps=(), ladder=None, outdir=None, resume=False):
    """Run the reference with one sampler per temperature and trace every iteration.

    p0s: [ntemps][ndim] initial points.  ext_jumps: list of (callable(x, iter, beta), weight) added
    with addProposalToCycle before sample(), as a user would.  Returns a dict of arrays.
    """
    ref = import_reference()
    sample_kwargs = dict(sample_kwargs or {})
    comms = ThreadComm.world(ntemps) if ntemps > 1 else [ref.MPI.COMM_WORLD]
    out = [None] * ntemps
    errors = []
    outdir = outdir or tempfile.mkdtemp(prefix="ptmcmc_ref_")

    def worker(rank):
        try:
            sampler = ref.PTSampler(ndim, logl, logp, np.copy(cov0), groups=groups, comm=comms[rank],
                                    outDir=outdir, verbose=False, seed=seed, resume=resume)
            if shim:
                sampler.stream = ShimStream(seed, 0, rank)
            for entry in ext_jumps:
                fn, wgt = entry[0], entry[1]
                if getattr(fn, "bind_sampler", False):  # a jump that draws from the sampler's own stream
                    fn = fn(sampler)
                sampler.addProposalToCycle(fn, wgt)
            tr = dict(x=[], lnl=[], lnp=[], jump=[], acc=[], U=[], S=[], swap_acc=[])
            orig = sampler.PTMCMCOneStep

            def traced(p0, lnlike0, lnprob0, it):
                if shim:
                    sampler.stream.begin_step(it)
                before = {k: list(v) for k, v in sampler.jumpDict.items()}
                res = orig(p0, lnlike0, lnprob0, it)
                name, accepted = None, 0
                for k, v in sampler.jumpDict.items():
                    b = before.get(k, [0, 0])
                    if v[0] != b[0]:
                        name, accepted = k, int(v[1] != b[1])
                tr["x"].append(np.array(res[0], dtype=float))
                tr["lnl"].append(float(res[1]))
                tr["lnp"].append(float(res[2]))
                tr["jump"].append(name)
                tr["acc"].append(accepted)
                tr["swap_acc"].append(sampler.nswap_accepted)
                cu = sampler.covUpdate
                if (it - 1) % cu == 0 and it - 1 != 0:
                    tr["U"].append(np.concatenate([np.asarray(u).ravel() for u in sampler.U]))
                    tr["S"].append(np.concatenate([np.asarray(s_).ravel() for s_ in sampler.S]))
                return res

            sampler.PTMCMCOneStep = traced
            kw = dict(sample_kwargs)
            if ladder is not None:
                kw["ladder"] = np.asarray(ladder, dtype=float)
            sampler.sample(np.array(p0s[rank], dtype=float), niter, **kw)
            tr["sampler"] = sampler
            out[rank] = tr
        except BaseException as e:  # pragma: no cover - surfaced below
            errors.append(e)
            try:
                comms[rank].sh.barrier.abort()
            except Exception:
                pass

    if ntemps == 1:
        worker(0)
    else:
        ths = [threading.Thread(target=worker, args=(r,)) for r in range(ntemps)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    if errors:
        raise errors[0]
    s0 = out[0]["sampler"]
    names = {"covarianceJumpProposalSCAM": orc.JUMP_SCAM, "covarianceJumpProposalAM": orc.JUMP_AM,
             "DEJump": orc.JUMP_DE}
    next_ext = orc.JUMP_EXT0
    for entry in ext_jumps:  # (fn, weight) -> next external id; (fn, weight, jump_id) -> that built-in id
        if len(entry) > 2:
            names[getattr(entry[0], "jump_name", entry[0].__name__)] = entry[2]
        else:
            names[entry[0].__name__] = next_ext
            next_ext += 1
    names[None] = -1  # iterations replayed from the chain file on resume (ref :591-599) propose nothing

    def _by_id(col):
        arr = np.zeros((ntemps, max(names.values()) + 1), dtype=np.int64)
        for nm, jid in names.items():
            if nm is None:
                continue
            for t in range(ntemps):
                arr[t, jid] = out[t]["sampler"].jumpDict.get(nm, [0, 0])[col]
        return arr

    res = dict(
        x=np.array([[out[t]["x"][i] for t in range(ntemps)] for i in range(niter)]),
        lnl=np.array([[out[t]["lnl"][i] for t in range(ntemps)] for i in range(niter)]),
        lnp=np.array([[out[t]["lnp"][i] for t in range(ntemps)] for i in range(niter)]),
        jump=np.array([[names[out[t]["jump"][i]] for t in range(ntemps)] for i in range(niter)], dtype=np.int8),
        acc=np.array([[out[t]["acc"][i] for t in range(ntemps)] for i in range(niter)], dtype=np.int8),
        swap_acc=np.array([[out[t]["swap_acc"][i] for t in range(ntemps)] for i in range(niter)], dtype=np.int64),
        U=np.array(out[0]["U"]), S=np.array(out[0]["S"]),
        cov=np.array(s0.cov), mu=np.array(s0.mu), m2=np.array(s0.M2),
        am=np.array(s0._AMbuffer), de=np.array(s0._DEbuffer),
        chain=np.array(s0._chain), chain_lnl=np.array(s0._lnlike), chain_lnp=np.array(s0._lnprob),
        ladder=np.array(s0.ladder, dtype=float),
        naccepted=np.array([out[t]["sampler"].naccepted for t in range(ntemps)]),
        swap_proposed=np.array(s0.swapProposed),
        jump_prop=_by_id(0), jump_acc=_by_id(1),  # [T][jump id]: proposed / accepted (ref jumpDict :602, :622)
    )
    res["_samplers"] = [o["sampler"] for o in out]
    res["_outdir"] = outdir
    return res
