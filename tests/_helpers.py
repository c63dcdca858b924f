"""Shared helpers for the parity tests: fixture loading and oracle construction."""
import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_ext_jump_factory(lo, hi):
    def golden_ext_jump(x, it, beta):
        frac = np.modf(np.abs(np.sin(np.arange(1, len(x) + 1) * (it + 1.0) + 13.0 * x)) * 1e4)[0]
        return lo + (hi - lo) * (0.45 + 0.1 * frac), 0.05 * np.sin(it) * beta

    return golden_ext_jump


def fixture_groups(g):
    if "has_groups" not in g or not int(g["has_groups"]):
        return None
    off, idx = g["group_offsets"], g["group_indices"]
    return [idx[off[i]:off[i + 1]].astype(np.int32) for i in range(len(off) - 1)]


def fixture_target(g):
    """(logl_kind, logl_params, logp_kind, logp_params) of a trajectory fixture."""
    if str(g["kind"]) == "gaussian":
        return (orc.LOGL_GAUSSIAN, orc.gaussian_params(g["pb_mu"], g["pb_icov"]), orc.LOGP_UNIFORM,
                orc.uniform_params(g["pb_lo"], g["pb_hi"], 0.0, bool(int(g["inclusive"]))))
    return (orc.LOGL_CURVED, None, orc.LOGP_UNIFORM,
            orc.uniform_params(g["pb_lo"], g["pb_hi"], 0.0, bool(int(g["inclusive"]))))


def fixture_cycle(g):
    """Proposal cycle in the reference's registration order: user jumps first (added before
    sample()), then SCAM, AM (ref PTMCMCSampler.py:261-264); DE joins at burn+1."""
    cyc = []
    if int(g["ext"]):
        cyc.append((orc.JUMP_EXT0, 7))
    if "prior_weight" in g and int(g["prior_weight"]):
        cyc.append((orc.JUMP_PRIOR, int(g["prior_weight"])))
    if int(g["kw_SCAMweight"]):
        cyc.append((orc.JUMP_SCAM, int(g["kw_SCAMweight"])))
    if int(g["kw_AMweight"]):
        cyc.append((orc.JUMP_AM, int(g["kw_AMweight"])))
    return tuple(cyc)


def fixture_mh_temp(g):
    """MH temperatures: the ladder, with 1e80 on the last rung when the run used hotChain (ref :281-282)."""
    if "kw_hotChain" in g and bool(g["kw_hotChain"]) and int(g["T"]) > 1:
        mh = np.array(g["ladder"], dtype=float)
        mh[-1] = 1e80
        return mh
    return None


def oracle_from_fixture(g, inject=True, thin=1, **over):
    d, T, N = int(g["d"]), int(g["T"]), int(g["N"])
    lk, lpar, pk, ppar = fixture_target(g)
    ext = None
    if int(g["ext"]):
        fn = golden_ext_jump_factory(g["pb_lo"], g["pb_hi"])
        ext = lambda k, x, it, beta, w, t: fn(x, it, beta)  # noqa: E731
    kw = dict(seed=int(g["seed"]), ladder=g["ladder"], mh_temp=fixture_mh_temp(g), groups=fixture_groups(g),
              cycle=fixture_cycle(g),
              de_weight=int(g["kw_DEweight"]), cov_update=int(g["kw_covUpdate"]), burn=int(g["kw_burn"]),
              tskip=int(g["kw_Tskip"]), thin=thin, logl_kind=lk, logl_params=lpar, logp_kind=pk,
              logp_params=ppar, record_hot=True, max_rows=N // thin + 1, ext_jump=ext)
    kw.update(over)
    o = orc.Oracle(d, 1, T, g["cov0"], **kw)
    if inject and len(g["U"]):
        o.inject_factors(g["U"], g["S"])
    return o
