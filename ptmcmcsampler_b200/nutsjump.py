"""Gradient-based jump proposals: MALA, HMC and NUTS (the reference's ``PTMCMCSampler/nutsjump.py``).

Like the reference's, these are ordinary plugin proposals ``jump(x, iter, beta) -> (q, qxy)`` built from the user's
``logl_grad(x) -> (logl, grad)`` and ``logp_grad(x) -> (logp, grad)`` callables (ref PTMCMCSampler.py:110-115, :226-258)
and registered with ``addProposalToCycle``; they run on the host between the engine's propose / accept calls, because
the gradients are Python.  They work in coordinates whitened by the Cholesky factor of the proposal covariance
(ref nutsjump.py:50-54, :78-90).  Written from the algorithms, not from the reference's code:

* **MALA** -- one Langevin step along a random axis of the whitened space, with the exact Hastings ratio of its
  Gaussian proposal (ref :182-235).
* **HMC** -- leapfrog trajectory of a random length in ``[nminsteps, nmaxsteps)``; ``qxy`` is the change of kinetic
  energy, so that the sampler's own ``lnprob`` difference completes the Hamiltonian (ref :238-291).  With
  ``vectorized = True`` every chain of the ensemble integrates in lock step (one batched gradient call per leapfrog).
* **NUTS** -- the no-U-turn sampler with dual-averaging step-size adaptation (Hoffman & Gelman 2014, algorithm 6);
  the tree's draw is already a sample of the trajectory, so ``qxy`` cancels the sampler's ``lnprob`` difference and the
  move is always taken, as in the reference (ref :379-840).

Two deliberate differences from the reference, both switchable with ``compat=True``: its HMC stops a trajectory when
``joint1 - 1000 < joint0`` (true after the first leapfrog, ref :283-285) and returns ``qxy = joint1 - joint0``, which
counts the ``lnprob`` difference twice (ref :287); here the trajectory runs its drawn length unless the energy error
exceeds 1000, and ``qxy`` is the kinetic part only.  The reference draws from the global ``np.random`` state
(ref :94, :222, :277); these classes take a ``numpy.random.Generator``.
"""
import numpy as np

__all__ = ["GradientJump", "MALAJump", "HMCJump", "NUTSJump"]


class GradientJump(object):
    """Shared machinery: whitening by the Cholesky factor of ``mm_inv`` (the proposal covariance), tempered target."""

    name = "GradientJump"
    vectorized = False

    def __init__(self, loglik_grad, logprior_grad, mm_inv, nburn=100, rng=None, batched_gradients=False):
        self._loglik_grad, self._logprior_grad = loglik_grad, logprior_grad
        self.mm_inv = np.array(mm_inv, dtype=np.float64)
        self.ndim = len(self.mm_inv)
        self.nburn = int(nburn)
        self.rng = np.random.default_rng() if rng is None else rng
        self.batched_gradients = bool(batched_gradients)
        self.iter = 0
        self.set_cf()

    @property
    def __name__(self):
        return self.name

    def set_cf(self):
        """x = L^T q with L the lower Cholesky factor of the mass-matrix inverse (ref :50-54)."""
        self.cov_cf = np.linalg.cholesky(self.mm_inv)
        self.cov_cfi = np.linalg.inv(self.cov_cf)

    def forward(self, x):
        return x @ self.cov_cfi          # q = L^-T x, for one point [d] or a batch [n, d]

    def backward(self, q):
        return q @ self.cov_cf           # x = L^T q

    def _grad_batch(self, fn, X):
        if self.batched_gradients:
            v, g = fn(X)
            return np.asarray(v, dtype=np.float64), np.asarray(g, dtype=np.float64)
        out = [fn(x) for x in X]
        return np.array([o[0] for o in out], dtype=np.float64), np.array([o[1] for o in out], dtype=np.float64)

    def func_grad_white(self, Q, beta):
        """Tempered log-density ``beta logl + logp`` and its gradient in whitened coordinates; Q is [n, d], beta [n]."""
        X = self.backward(Q)
        ll, gl = self._grad_batch(self._loglik_grad, X)
        lp, gp = self._grad_batch(self._logprior_grad, X)
        val = beta * ll + lp
        grad = (beta[:, None] * gl + gp) @ self.cov_cf.T
        bad = ~np.isfinite(val)
        if bad.any():   # outside the prior support: a wall of -inf with no force
            val = np.where(bad, -np.inf, val)
            grad = np.where(bad[:, None], 0.0, grad)
        return val, grad

    def _batch(self, x, beta):
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 1
        X = x[None, :] if single else x
        b = np.broadcast_to(np.asarray(beta, dtype=np.float64), (len(X),)).copy()
        return X, b, single


class MALAJump(GradientJump):
    """Metropolis-adjusted Langevin step along one random axis of the whitened space (ref :182-235)."""

    name = "MALAJump"
    vectorized = True

    def __init__(self, *a, **kw):
        super(MALAJump, self).__init__(*a, **kw)
        self.cd = 2.4 / np.sqrt(self.ndim)

    def __call__(self, x, iter, beta):
        self.iter += 1
        X, b, single = self._batch(x, beta)
        n = len(X)
        q0 = self.forward(X)
        _, g0 = self.func_grad_white(q0, b)
        axis = self.rng.integers(0, self.ndim, n)
        z = self.rng.standard_normal(n)
        rows = np.arange(n)
        drift0 = 0.25 * self.cd**2 * g0[rows, axis]           # mean of the proposal: q0 + cd^2/4 * grad along the axis
        q1 = q0.copy()
        q1[rows, axis] += drift0 + self.cd * z
        _, g1 = self.func_grad_white(q1, b)
        drift1 = 0.25 * self.cd**2 * g1[rows, axis]
        fwd = q1[rows, axis] - (q0[rows, axis] + drift0)       # forward and reverse displacements from the proposal means
        rev = q0[rows, axis] - (q1[rows, axis] + drift1)
        qxy = 0.5 * (fwd**2 - rev**2) / self.cd**2             # log q(x | y) - log q(y | x)
        Q = self.backward(q1)
        return (Q[0], float(qxy[0])) if single else (Q, qxy)


class HMCJump(GradientJump):
    """Hamiltonian Monte Carlo with unit mass in the whitened space (ref :238-291)."""

    name = "HMCJump"
    vectorized = True

    def __init__(self, loglik_grad, logprior_grad, mm_inv, nburn=100, stepsize=0.1, nminsteps=10, nmaxsteps=300, compat=False,
                 **kw):
        super(HMCJump, self).__init__(loglik_grad, logprior_grad, mm_inv, nburn=nburn, **kw)
        self.epsilon, self.nminsteps, self.nmaxsteps, self.compat = float(stepsize), int(nminsteps), int(nmaxsteps), bool(compat)

    def __call__(self, x, iter, beta):
        self.iter += 1
        X, b, single = self._batch(x, beta)
        n, eps = len(X), self.epsilon
        q = self.forward(X)
        logp0, grad = self.func_grad_white(q, b)
        p = self.rng.standard_normal((n, self.ndim))
        kin0 = 0.5 * np.sum(p * p, axis=1)
        joint0 = logp0 - kin0
        nsteps = self.rng.integers(self.nminsteps, max(self.nminsteps + 1, self.nmaxsteps), n)
        live = np.ones(n, dtype=bool)
        logp = logp0.copy()
        for step in range(int(nsteps.max())):
            live &= step < nsteps
            if not live.any():
                break
            idx = np.nonzero(live)[0]
            ph = p[idx] + 0.5 * eps * grad[idx]                 # leapfrog (ref :159-165), live chains only
            qn = q[idx] + eps * ph
            lpn, gn = self.func_grad_white(qn, b[idx])
            pn = ph + 0.5 * eps * gn
            q[idx], p[idx], grad[idx], logp[idx] = qn, pn, gn, lpn
            joint = lpn - 0.5 * np.sum(pn * pn, axis=1)
            if self.compat:
                stop = (joint - 1000.0) < joint0[idx]           # the reference's test: true after the first step
            else:
                stop = ~(joint > joint0[idx] - 1000.0)          # hopelessly inaccurate (or NaN): give up on the trajectory
            live[idx[stop]] = False
        kin1 = 0.5 * np.sum(p * p, axis=1)
        qxy = (logp - kin1) - joint0 if self.compat else kin0 - kin1
        qxy = np.where(np.isfinite(qxy), qxy, -np.inf)
        Q = self.backward(q)
        return (Q[0], float(qxy[0])) if single else (Q, qxy)


class NUTSJump(GradientJump):
    """No-U-turn sampler (Hoffman & Gelman 2014, algorithm 6: slice variable, doubling, dual-averaging step size during
    the first ``nburn`` calls).  One chain per call."""

    name = "NUTSJump"

    def __init__(self, loglik_grad, logprior_grad, mm_inv, nburn=100, delta=0.6, max_depth=8, force_epsilon=None, **kw):
        super(NUTSJump, self).__init__(loglik_grad, logprior_grad, mm_inv, nburn=nburn, **kw)
        self.delta, self.max_depth = float(delta), int(max_depth)
        self.epsilon = force_epsilon
        self._adapt = force_epsilon is None
        self._mu = self._hbar = self._logeps_bar = None
        self._m = 0

    def _fg(self, q, beta):
        v, g = self.func_grad_white(q[None, :], np.array([beta]))
        return float(v[0]), g[0]

    def _leapfrog(self, q, p, g, eps, beta):
        ph = p + 0.5 * eps * g
        qn = q + eps * ph
        lp, gn = self._fg(qn, beta)
        return qn, ph + 0.5 * eps * gn, gn, lp

    def _find_epsilon(self, q, lp, g, beta):
        eps, p = 1.0, self.rng.standard_normal(self.ndim)
        h0 = lp - 0.5 * p @ p
        _, pn, _, lpn = self._leapfrog(q, p, g, eps, beta)
        a = 1.0 if (lpn - 0.5 * pn @ pn) - h0 > np.log(0.5) else -1.0
        for _ in range(60):
            _, pn, _, lpn = self._leapfrog(q, p, g, eps, beta)
            dh = (lpn - 0.5 * pn @ pn) - h0
            if not np.isfinite(dh):
                dh = -np.inf
            if a * dh <= -a * np.log(2.0):
                break
            eps *= 2.0**a
        return eps

    def _tree(self, q, p, g, logu, v, j, eps, h0, beta):
        if j == 0:
            qn, pn, gn, lp = self._leapfrog(q, p, g, v * eps, beta)
            h = lp - 0.5 * pn @ pn
            if not np.isfinite(h):
                h = -np.inf
            n = int(logu <= h)
            s = bool(logu < h + 1000.0)
            return qn, pn, gn, qn, pn, gn, qn, lp, n, s, min(1.0, np.exp(min(0.0, h - h0))), 1
        qm, pm, gm, qp, pp, gp, qc, lpc, n1, s1, a1, na1 = self._tree(q, p, g, logu, v, j - 1, eps, h0, beta)
        if s1:
            if v < 0:
                qm, pm, gm, _, _, _, qc2, lpc2, n2, s2, a2, na2 = self._tree(qm, pm, gm, logu, v, j - 1, eps, h0, beta)
            else:
                _, _, _, qp, pp, gp, qc2, lpc2, n2, s2, a2, na2 = self._tree(qp, pp, gp, logu, v, j - 1, eps, h0, beta)
            if n2 > 0 and self.rng.random() < n2 / max(1, n1 + n2):
                qc, lpc = qc2, lpc2
            a1, na1, n1 = a1 + a2, na1 + na2, n1 + n2
            s1 = s2 and (qp - qm) @ pm >= 0 and (qp - qm) @ pp >= 0
        return qm, pm, gm, qp, pp, gp, qc, lpc, n1, s1, a1, na1

    def __call__(self, x, iter, beta):
        self.iter += 1
        x = np.asarray(x, dtype=np.float64)
        beta = float(beta)
        q0 = self.forward(x)
        lp0, g0 = self._fg(q0, beta)
        if not np.isfinite(lp0):
            return x.copy(), 0.0
        if self.epsilon is None:
            self.epsilon = self._find_epsilon(q0, lp0, g0, beta)
            self._mu, self._hbar, self._logeps_bar = np.log(10.0 * self.epsilon), 0.0, 0.0
        eps = self.epsilon
        p0 = self.rng.standard_normal(self.ndim)
        h0 = lp0 - 0.5 * p0 @ p0
        logu = h0 + np.log(self.rng.random())
        qm = qp = qc = q0
        pm = pp = p0
        gm = gp = g0
        lpc, n, s, j, alpha, nalpha = lp0, 1, True, 0, 0.0, 1
        while s and j < self.max_depth:
            v = 1 if self.rng.random() < 0.5 else -1
            if v < 0:
                qm, pm, gm, _, _, _, qn, lpn, n2, s2, alpha, nalpha = self._tree(qm, pm, gm, logu, v, j, eps, h0, beta)
            else:
                _, _, _, qp, pp, gp, qn, lpn, n2, s2, alpha, nalpha = self._tree(qp, pp, gp, logu, v, j, eps, h0, beta)
            if s2 and self.rng.random() < min(1.0, n2 / n):
                qc, lpc = qn, lpn
            n += n2
            s = s2 and (qp - qm) @ pm >= 0 and (qp - qm) @ pp >= 0
            j += 1
        if self._adapt and self._m < self.nburn:   # dual averaging (gamma = 0.05, t0 = 10, kappa = 0.75)
            self._m += 1
            m = self._m
            self._hbar = (1.0 - 1.0 / (m + 10.0)) * self._hbar + (self.delta - alpha / max(1, nalpha)) / (m + 10.0)
            logeps = self._mu - np.sqrt(m) / 0.05 * self._hbar
            self._logeps_bar = m**-0.75 * logeps + (1.0 - m**-0.75) * self._logeps_bar
            self.epsilon = float(np.exp(logeps if m < self.nburn else self._logeps_bar))
        # the tree's draw is a sample of the trajectory: cancel the sampler's lnprob difference so it is always taken
        return self.backward(qc), float(lp0 - lpc)
