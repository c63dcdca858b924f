"""Device-resident targets for :class:`ptmcmcsampler_b200.PTSampler`.

The reference takes arbitrary Python callables ``logl(x)`` / ``logp(x)``
(ref PTMCMCSampler.py:108-109).  Passing one of the objects below instead selects a built-in
CUDA implementation (``PTMCMC_LOGL_*`` / ``PTMCMC_LOGP_*`` in include/ptmcmc_b200.h) so the whole
Metropolis-Hastings step stays on the GPU.  Plain callables are still accepted by ``PTSampler`` and
are evaluated on the host once per iteration (slow path).

These objects only describe the target; they carry no host implementation of it.
"""
import numpy as np

from . import _cabi


class DeviceLogLikelihood(object):
    kind = _cabi.LOGL_EXTERNAL
    source = None        # CUDA source of a user target (SourceLikelihood)
    user_params = None

    def params(self, ndim):
        return None


class DeviceLogPrior(object):
    kind = _cabi.LOGP_EXTERNAL
    source = None
    user_params = None

    def params(self, ndim):
        return None


class SourceLikelihood(DeviceLogLikelihood):
    """An arbitrary log-likelihood as CUDA C++ source, compiled with NVRTC into the engine's MH kernel when the sampler
    starts (the reference's ``logl(x, *args)`` Python callable, ref :108, :1072-1086, moved onto the device)::

        SourceLikelihood('''
            double user_logl(const double *x, int ndim, const double *par) {
                double s = 0.0;
                for (int i = 0; i < ndim; ++i) s += (x[i] - par[i]) * (x[i] - par[i]);
                return -0.5 * s;
            }''', params=mu)

    The source must define ``user_logl(const double *x, int ndim, const double *par)`` returning a double; ``params`` (the
    reference's ``loglargs``) arrive as ``par``.  Helper functions may be defined alongside; every function of the source
    is a device function.  A compile error raises ``ValueError`` with the NVRTC log."""

    kind = _cabi.LOGL_USER

    def __init__(self, source, params=None):
        self.source = source
        self.user_params = None if params is None else np.ascontiguousarray(params, dtype=np.float64).ravel()


class SourcePrior(DeviceLogPrior):
    """An arbitrary log-prior as CUDA C++ source defining ``user_logp(const double *x, int ndim, const double *par)``;
    return ``-INFINITY`` outside the support (the log-likelihood is then not evaluated, ref :607-608)."""

    kind = _cabi.LOGP_USER

    def __init__(self, source, params=None):
        self.source = source
        self.user_params = None if params is None else np.ascontiguousarray(params, dtype=np.float64).ravel()


class GaussianLikelihood(DeviceLogLikelihood):
    """``-0.5 (x-mu)^T icov (x-mu) + offset`` (ref examples/simple.py:34-36)."""

    kind = _cabi.LOGL_GAUSSIAN

    def __init__(self, mu, cov=None, icov=None, offset=0.0):
        self.mu = np.asarray(mu, dtype=np.float64)
        if (cov is None) == (icov is None):
            raise ValueError("give exactly one of cov / icov")
        self.icov = np.linalg.inv(np.asarray(cov, dtype=np.float64)) if icov is None else np.asarray(icov, np.float64)
        self.offset = float(offset)

    def params(self, ndim):
        if self.mu.shape != (ndim,) or self.icov.shape != (ndim, ndim):
            raise ValueError("GaussianLikelihood shape does not match ndim=%d" % ndim)
        return np.concatenate([self.mu, self.icov.ravel(), [self.offset]])


class CurvedLikelihood(DeviceLogLikelihood):
    """Sum over consecutive (x, y) pairs of the reference's curved bimodal density
    ``log(exp(-x^2-(9+4x^2+9y)^2) + 0.5 exp(-8x^2-8(y-2)^2))`` (ref examples/curved_likelihood.ipynb)."""

    kind = _cabi.LOGL_CURVED


class RosenbrockLikelihood(DeviceLogLikelihood):
    """``-sum(100 (x[i+1]-x[i]^2)^2 + (1-x[i])^2) / 20``."""

    kind = _cabi.LOGL_ROSENBROCK


class UniformPrior(DeviceLogPrior):
    """Box prior: ``value`` inside ``[pmin, pmax]`` (closed if ``inclusive``, as in
    ref examples/simple.py:38-44; open as in the curved example), ``-inf`` outside."""

    kind = _cabi.LOGP_UNIFORM

    def __init__(self, pmin, pmax, inclusive=True, value=0.0):
        self.pmin, self.pmax = pmin, pmax
        self.inclusive, self.value = bool(inclusive), float(value)

    def params(self, ndim):
        lo = np.broadcast_to(np.asarray(self.pmin, dtype=np.float64), (ndim,))
        hi = np.broadcast_to(np.asarray(self.pmax, dtype=np.float64), (ndim,))
        return np.concatenate([lo, hi, [self.value, 1.0 if self.inclusive else 0.0]])


class FlatPrior(DeviceLogPrior):
    """Improper flat prior: 0 everywhere."""

    kind = _cabi.LOGP_FLAT
