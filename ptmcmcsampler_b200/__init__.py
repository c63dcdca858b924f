"""B200-native parallel-tempering MCMC engine behind the PTMCMCSampler API.

``PTSampler`` mirrors ``PTMCMCSampler.PTMCMCSampler.PTSampler`` of nanograv/PTMCMCSampler; the hot
path (MH step, adaptive covariance, DE history, temperature swap) runs in hand-written sm_100a
CUDA reached through the C ABI in ``include/ptmcmc_b200.h``.
"""
from . import _cabi  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("PTSampler", "PTMCMCSampler"):
        from . import PTMCMCSampler as _m

        return _m if name == "PTMCMCSampler" else _m.PTSampler
    raise AttributeError(name)
