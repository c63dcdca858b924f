"""ctypes wrapper around the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this module.  See ``ptmcmc_oracle.h``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libptmcmc_oracle.so")

JUMP_SCAM, JUMP_AM, JUMP_DE, JUMP_PRIOR, JUMP_EXT0 = 0, 1, 2, 3, 4
LOGL_EXTERNAL, LOGL_GAUSSIAN, LOGL_CURVED, LOGL_ROSENBROCK = 0, 1, 2, 3
LOGP_EXTERNAL, LOGP_UNIFORM, LOGP_FLAT = 0, 1, 2
PURPOSE_MH, PURPOSE_SWAP = 0, 1

LOGFN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)
JUMPFN = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int64, C.c_double, C.c_int,
                     C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


class _Config(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nwalkers", C.c_int32), ("ntemps", C.c_int32),
        ("walker_offset", C.c_int32), ("temp_offset", C.c_int32),
        ("seed", C.c_uint64),
        ("ladder", C.POINTER(C.c_double)), ("mh_temp", C.POINTER(C.c_double)),
        ("cov", C.POINTER(C.c_double)),
        ("ngroups", C.c_int32),
        ("group_offsets", C.POINTER(C.c_int32)), ("group_indices", C.POINTER(C.c_int32)),
        ("ncycle", C.c_int32),
        ("cycle_jump", C.POINTER(C.c_int32)), ("cycle_weight", C.POINTER(C.c_int32)),
        ("de_weight", C.c_int32),
        ("cov_update", C.c_int64), ("burn", C.c_int64), ("tskip", C.c_int64), ("thin", C.c_int64),
        ("logl_kind", C.c_int32), ("logl_params", C.POINTER(C.c_double)),
        ("logp_kind", C.c_int32), ("logp_params", C.POINTER(C.c_double)),
        ("record_hot", C.c_int32), ("max_rows", C.c_int64), ("nthreads", C.c_int32),
        ("ext_logl", LOGFN), ("ext_logp", LOGFN), ("ext_jump", JUMPFN), ("user", C.c_void_p),
        ("ntemps_global", C.c_int32), ("ladder_above", C.c_double), ("ladder_below", C.c_double),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ptmcmc_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Config)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_state.argtypes = [C.c_void_p, dp]
        L.orc_run.argtypes = [C.c_void_p, C.c_int64]
        L.orc_iteration.restype = C.c_int64
        L.orc_iteration.argtypes = [C.c_void_p]
        L.orc_rows.restype = C.c_int64
        L.orc_rows.argtypes = [C.c_void_p]
        L.orc_njumps.restype = C.c_int32
        L.orc_njumps.argtypes = [C.c_void_p]
        L.orc_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_get_chain.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_get_adapt.argtypes = [C.c_void_p, dp, dp, dp, i64p]
        L.orc_get_factor.argtypes = [C.c_void_p, dp, dp]
        L.orc_set_factor.argtypes = [C.c_void_p, dp, dp]
        L.orc_inject_factors.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_get_buffers.argtypes = [C.c_void_p, dp, dp]
        L.orc_get_counters.argtypes = [C.c_void_p, i64p, i64p, i64p, i64p]
        L.orc_set_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int16), C.c_int64]
        L.orc_swap_msg_doubles.restype = C.c_int64
        L.orc_swap_msg_doubles.argtypes = [C.c_void_p]
        L.orc_swap_pending.argtypes = [C.c_void_p]
        L.orc_swap_pack_top.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_swap_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_swap_finish.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_maintain.argtypes = [C.c_void_p]
        L.orc_adapt_begin.argtypes = [C.c_void_p, dp]
        L.orc_adapt_finish.argtypes = [C.c_void_p, dp]
        L.orc_am_ring.restype = C.c_void_p
        L.orc_am_ring.argtypes = [C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_draw_word.restype = C.c_uint64
        L.orc_draw_word.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_word_to_int.restype = C.c_uint64
        L.orc_word_to_int.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_word_to_unit.restype = C.c_double
        L.orc_word_to_unit.argtypes = [C.c_uint64]
        L.orc_word_to_normals.argtypes = [C.c_uint64, dp, dp]
        L.orc_word_to_normals_many.argtypes = [C.POINTER(C.c_uint64), C.c_int64, dp, dp]
        L.orc_sym_factor.argtypes = [C.c_int, dp, dp, dp]
        L.orc_temperature_ladder.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def draw_word(seed, purpose, it, walker, temp, j):
    return int(lib().orc_draw_word(seed, purpose, it, walker, temp, j))


def word_to_int(word, n):
    return int(lib().orc_word_to_int(word, n))


def word_to_unit(word):
    return float(lib().orc_word_to_unit(word))


def word_to_normals(word):
    a, b = C.c_double(), C.c_double()
    lib().orc_word_to_normals(word, C.byref(a), C.byref(b))
    return a.value, b.value


def word_to_normals_many(words):
    words = np.ascontiguousarray(words, dtype=np.uint64)
    z0, z1 = np.empty(words.size), np.empty(words.size)
    lib().orc_word_to_normals_many(words.ctypes.data_as(C.POINTER(C.c_uint64)), words.size, _dp(z0), _dp(z1))
    return z0, z1


def sym_factor(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    U = np.zeros((n, n))
    S = np.zeros(n)
    lib().orc_sym_factor(n, _dp(a), _dp(U), _dp(S))
    return U, S


def temperature_ladder(ndim, ntemps, tmin=1.0, tmax=None):
    out = np.zeros(ntemps)
    lib().orc_temperature_ladder(ndim, ntemps, float(tmin), float(tmax) if tmax else -1.0, _dp(out))
    return out


def gaussian_params(mu, icov, offset=0.0):
    mu = np.asarray(mu, dtype=np.float64)
    return np.concatenate([mu, np.asarray(icov, dtype=np.float64).ravel(), [offset]])


def uniform_params(lo, hi, inside=0.0, inclusive=True):
    return np.concatenate([np.asarray(lo, float), np.asarray(hi, float), [inside, 1.0 if inclusive else 0.0]])


class Oracle(object):
    """W walkers x T temperatures of the reference algorithm on the CPU."""

    def __init__(self, ndim, nwalkers, ntemps, cov, seed=0, ladder=None, mh_temp=None, groups=None,
                 cycle=((JUMP_SCAM, 20), (JUMP_AM, 20)), de_weight=20, cov_update=1000, burn=10000,
                 tskip=100, thin=10, logl_kind=LOGL_GAUSSIAN, logl_params=None,
                 logp_kind=LOGP_UNIFORM, logp_params=None, record_hot=False, max_rows=1,
                 nthreads=1, walker_offset=0, temp_offset=0, ext_logl=None, ext_logp=None, ext_jump=None,
                 ntemps_global=0, ladder_above=0.0, ladder_below=0.0):
        L = lib()
        self.d, self.W, self.T = ndim, nwalkers, ntemps
        self.cov_update, self.burn = cov_update, burn
        self._keep = []
        cfg = _Config()
        cfg.ndim, cfg.nwalkers, cfg.ntemps = ndim, nwalkers, ntemps
        cfg.walker_offset, cfg.temp_offset, cfg.seed = walker_offset, temp_offset, seed
        if ladder is None:
            ladder = temperature_ladder(ndim, ntemps)
        ladder = np.ascontiguousarray(ladder, dtype=np.float64)
        mh_temp = ladder if mh_temp is None else np.ascontiguousarray(mh_temp, dtype=np.float64)
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        self._keep += [ladder, mh_temp, cov]
        cfg.ladder, cfg.mh_temp, cfg.cov = _dp(ladder), _dp(mh_temp), _dp(cov)
        if groups is None:
            cfg.ngroups = 0
            self.groups = [np.arange(ndim)]
        else:
            self.groups = [np.asarray(g, dtype=np.int32) for g in groups]
            offs = np.zeros(len(groups) + 1, dtype=np.int32)
            offs[1:] = np.cumsum([len(g) for g in self.groups])
            idx = np.concatenate(self.groups).astype(np.int32)
            self._keep += [offs, idx]
            cfg.ngroups, cfg.group_offsets, cfg.group_indices = len(groups), _ip(offs), _ip(idx)
        cj = np.array([c[0] for c in cycle], dtype=np.int32)
        cw = np.array([c[1] for c in cycle], dtype=np.int32)
        self._keep += [cj, cw]
        cfg.ncycle, cfg.cycle_jump, cfg.cycle_weight, cfg.de_weight = len(cj), _ip(cj), _ip(cw), de_weight
        cfg.cov_update, cfg.burn, cfg.tskip, cfg.thin = cov_update, burn, tskip, thin
        cfg.logl_kind, cfg.logp_kind = logl_kind, logp_kind
        if logl_params is not None:
            lp_ = np.ascontiguousarray(logl_params, dtype=np.float64)
            self._keep.append(lp_)
            cfg.logl_params = _dp(lp_)
        if logp_params is not None:
            pp_ = np.ascontiguousarray(logp_params, dtype=np.float64)
            self._keep.append(pp_)
            cfg.logp_params = _dp(pp_)
        cfg.record_hot, cfg.max_rows, cfg.nthreads = int(record_hot), max_rows, nthreads
        if ext_logl is not None:
            f = LOGFN(lambda x, n, u: float(ext_logl(np.ctypeslib.as_array(x, (n,)).copy())))
            self._keep.append(f)
            cfg.ext_logl = f
        if ext_logp is not None:
            f = LOGFN(lambda x, n, u: float(ext_logp(np.ctypeslib.as_array(x, (n,)).copy())))
            self._keep.append(f)
            cfg.ext_logp = f
        if ext_jump is not None:
            def _j(k, x, n, it, beta, w, t, q, qxy, u):
                qq, lq = ext_jump(k, np.ctypeslib.as_array(x, (n,)).copy(), it, beta, w, t)
                np.ctypeslib.as_array(q, (n,))[:] = qq
                qxy[0] = lq
            f = JUMPFN(_j)
            self._keep.append(f)
            cfg.ext_jump = f
        cfg.ntemps_global, cfg.ladder_above, cfg.ladder_below = int(ntemps_global), float(ladder_above), float(ladder_below)
        self.temp_offset, self.ntemps_global = int(temp_offset), int(ntemps_global) or int(ntemps)
        self.ntr = ntemps if record_hot else 1
        self.usize = sum(len(g) ** 2 for g in self.groups)
        self.ssize = sum(len(g) for g in self.groups)
        self._h = L.orc_create(C.byref(cfg))
        self.njumps = L.orc_njumps(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def set_state(self, x0):
        x0 = np.ascontiguousarray(np.broadcast_to(x0, (self.T, self.W, self.d)), dtype=np.float64)
        return lib().orc_set_state(self._h, _dp(x0))

    def run(self, niter):
        rc = lib().orc_run(self._h, niter)
        if rc:
            raise ValueError("oracle run failed rc=%d" % rc)

    @property
    def iteration(self):
        return lib().orc_iteration(self._h)

    def state(self):
        x = np.zeros((self.T, self.W, self.d))
        lnl, lp, lnp = np.zeros((self.T, self.W)), np.zeros((self.T, self.W)), np.zeros((self.T, self.W))
        lib().orc_get_state(self._h, _dp(x), _dp(lnl), _dp(lp), _dp(lnp))
        return x, lnl, lp, lnp

    def chain(self):
        rows = lib().orc_rows(self._h)
        ch = np.zeros((rows, self.ntr, self.W, self.d))
        lnl, lnp = np.zeros((rows, self.ntr, self.W)), np.zeros((rows, self.ntr, self.W))
        lib().orc_get_chain(self._h, _dp(ch), _dp(lnl), _dp(lnp))
        return ch, lnl, lnp

    def adapt(self):
        cov, mu, m2 = np.zeros((self.d, self.d)), np.zeros(self.d), np.zeros((self.d, self.d))
        n = C.c_int64()
        lib().orc_get_adapt(self._h, _dp(cov), _dp(mu), _dp(m2), C.byref(n))
        return cov, mu, m2, n.value

    def factor(self):
        U, S = np.zeros(self.usize), np.zeros(self.ssize)
        lib().orc_get_factor(self._h, _dp(U), _dp(S))
        return U, S

    def set_factor(self, U, S):
        U = np.ascontiguousarray(U, dtype=np.float64).ravel()
        S = np.ascontiguousarray(S, dtype=np.float64).ravel()
        lib().orc_set_factor(self._h, _dp(U), _dp(S))

    def inject_factors(self, Us, Ss):
        Us = np.ascontiguousarray(Us, dtype=np.float64).reshape(len(Us), -1)
        Ss = np.ascontiguousarray(Ss, dtype=np.float64).reshape(len(Ss), -1)
        assert Us.shape[1] == self.usize and Ss.shape[1] == self.ssize
        lib().orc_inject_factors(self._h, len(Us), _dp(Us), _dp(Ss))

    def buffers(self):
        am = np.zeros((self.cov_update, self.W, self.d))
        de = np.zeros((self.burn, self.W, self.d))
        lib().orc_get_buffers(self._h, _dp(am), _dp(de))
        return am, de

    def counters(self):
        shp = (self.T, self.W, self.njumps)
        prop, acc = np.zeros(shp, dtype=np.int64), np.zeros(shp, dtype=np.int64)
        sw = np.zeros((self.T, self.W), dtype=np.int64)
        n = C.c_int64()
        i64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))  # noqa: E731
        lib().orc_get_counters(self._h, i64(prop), i64(acc), i64(sw), C.byref(n))
        return prop, acc, sw, n.value

    # ---- ladder sharding (same surface as ptmcmcsampler_b200._cabi.Engine; pointers are host addresses)
    @property
    def swap_msg_doubles(self):
        return int(lib().orc_swap_msg_doubles(self._h))

    @property
    def swap_pending(self):
        return bool(lib().orc_swap_pending(self._h))

    def swap_pack_top(self, msg_ptr):
        lib().orc_swap_pack_top(self._h, msg_ptr)

    def swap_sweep(self, carry_in_ptr, carry_out_ptr):
        rc = lib().orc_swap_sweep(self._h, carry_in_ptr or None, carry_out_ptr or None)
        if rc:
            raise ValueError("oracle swap_sweep failed rc=%d" % rc)

    def swap_finish(self, below_ptr):
        rc = lib().orc_swap_finish(self._h, below_ptr or None)
        if rc:
            raise ValueError("oracle swap_finish failed rc=%d" % rc)

    def adapt_begin(self):
        batch = np.zeros(1 + self.d + self.d * self.d)
        return batch if lib().orc_adapt_begin(self._h, _dp(batch)) == 1 else None

    def adapt_finish(self, batch):
        batch = np.ascontiguousarray(batch, dtype=np.float64)
        if lib().orc_adapt_finish(self._h, _dp(batch)):
            raise ValueError("oracle adapt_finish: no covariance update is due")

    def maintain(self):
        rc = lib().orc_maintain(self._h)
        if rc:
            raise ValueError("oracle maintain failed rc=%d" % rc)

    def am_ring(self):
        """(host address, number of doubles) of the AM ring."""
        return int(lib().orc_am_ring(self._h)), self.cov_update * self.W * self.d

    def set_trace(self, niter, nswaps=0):
        self.trace = np.zeros((niter, self.T, self.W), dtype=np.uint8)
        self.swapmaps = np.full((max(nswaps, 1), self.W, self.T), -1, dtype=np.int16)
        lib().orc_set_trace(self._h, self.trace.ctypes.data_as(C.POINTER(C.c_uint8)), niter,
                            self.swapmaps.ctypes.data_as(C.POINTER(C.c_int16)), nswaps)
        return self.trace, self.swapmaps
