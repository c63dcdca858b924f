"""B200-native parallel-tempering MCMC engine behind the PTMCMCSampler API.

``PTSampler`` mirrors ``PTMCMCSampler.PTMCMCSampler.PTSampler`` of nanograv/PTMCMCSampler; the hot
path (MH step, adaptive covariance, DE history, temperature swap) runs in hand-written sm_100a
CUDA reached through the C ABI in ``include/ptmcmc_b200.h``.
"""
from . import _cabi  # noqa: F401
from . import PTMCMCSampler  # noqa: F401
from .PTMCMCSampler import PTSampler  # noqa: F401

__version__ = "0.1.0"
