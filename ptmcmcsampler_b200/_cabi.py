"""ctypes binding of the engine's C ABI (include/ptmcmc_b200.h).

There is no CPU implementation behind this module: if the shared library is missing, or no CUDA
device is visible when an engine is created, the error is raised to the caller.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libptmcmc_b200.so")
ABI_VERSION = 4
MAX_CYCLE = 16

JUMP_SCAM, JUMP_AM, JUMP_DE, JUMP_PRIOR, JUMP_EXT0 = 0, 1, 2, 3, 4
LOGL_EXTERNAL, LOGL_GAUSSIAN, LOGL_CURVED, LOGL_ROSENBROCK, LOGL_USER = 0, 1, 2, 3, 4
LOGP_EXTERNAL, LOGP_UNIFORM, LOGP_FLAT, LOGP_USER = 0, 1, 2, 3
ERR_ARG, ERR_DE_SHAPE, ERR_CUDA, ERR_STATE, ERR_CAPACITY = -1, -2, -3, -4, -5
K_NAMES = ["mh", "swap", "adapt", "de", "init", "propose", "accept", "reserved"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("ndim", C.c_int32), ("nwalkers", C.c_int32),
        ("ntemps", C.c_int32), ("walker_offset", C.c_int32), ("temp_offset", C.c_int32), ("ntemps_global", C.c_int32),
        ("seed", C.c_uint64),
        ("ladder", _dp), ("mh_temp", _dp), ("cov", _dp),
        ("ngroups", C.c_int32), ("reserved1", C.c_int32),
        ("group_offsets", _ip), ("group_indices", _ip),
        ("ncycle", C.c_int32), ("de_weight", C.c_int32),
        ("cycle_jump", C.c_int32 * MAX_CYCLE), ("cycle_weight", C.c_int32 * MAX_CYCLE),
        ("cov_update", C.c_int64), ("burn", C.c_int64), ("tskip", C.c_int64), ("thin", C.c_int64),
        ("logl_kind", C.c_int32), ("logp_kind", C.c_int32),
        ("logl_params", _dp), ("logp_params", _dp),
        ("record_hot", C.c_int32), ("trace", C.c_int32),
        ("record_rows", C.c_int64), ("trace_iters", C.c_int64),
        ("timing", C.c_int32), ("reserved2", C.c_int32),
        ("ladder_above", C.c_double), ("ladder_below", C.c_double),
        ("logl_source", C.c_char_p), ("logp_source", C.c_char_p), ("user_params", _dp),
        ("n_logl_user_params", C.c_int32), ("n_logp_user_params", C.c_int32),
    ]


class Timing(C.Structure):
    _fields_ = [("launches", C.c_int64 * 8), ("ms", C.c_double * 8), ("chain_steps", C.c_int64)]


SYMBOLS = [
    "ptmcmc_abi_version", "ptmcmc_device_count", "ptmcmc_create_error", "ptmcmc_last_error", "ptmcmc_create",
    "ptmcmc_destroy", "ptmcmc_set_state", "ptmcmc_set_state_external", "ptmcmc_run", "ptmcmc_propose",
    "ptmcmc_accept", "ptmcmc_iteration", "ptmcmc_sync", "ptmcmc_get_state", "ptmcmc_rows", "ptmcmc_row_base",
    "ptmcmc_get_chain", "ptmcmc_release_rows", "ptmcmc_get_adapt", "ptmcmc_get_factor", "ptmcmc_set_factor",
    "ptmcmc_get_buffers", "ptmcmc_adapt_begin", "ptmcmc_adapt_finish", "ptmcmc_njumps", "ptmcmc_get_counters",
    "ptmcmc_get_trace", "ptmcmc_get_timing", "ptmcmc_reset_timing", "ptmcmc_stream", "ptmcmc_set_timing",
    "ptmcmc_host_alloc", "ptmcmc_host_free", "ptmcmc_swap_msg_doubles", "ptmcmc_swap_pending",
    "ptmcmc_swap_pack_top", "ptmcmc_swap_sweep", "ptmcmc_swap_finish", "ptmcmc_am_ring", "ptmcmc_maintain",
    "ptmcmc_state_bytes", "ptmcmc_save_state", "ptmcmc_load_state", "ptmcmc_replay", "ptmcmc_mh_kernel_name",
    "ptmcmc_test_normals", "ptmcmc_measure_fp64_peak", "ptmcmc_set_sink", "ptmcmc_sink_wait", "ptmcmc_snapshot_bytes",
    "ptmcmc_snapshot", "ptmcmc_snapshot_wait", "ptmcmc_adapt_begin_dev", "ptmcmc_adapt_finish_dev", "ptmcmc_factor_dev",
    "ptmcmc_factor_refresh", "ptmcmc_user_compile_check", "ptmcmc_callback_buffers", "ptmcmc_propose_pinned",
    "ptmcmc_accept_pinned", "ptmcmc_state_seed", "ptmcmc_p2p_open", "ptmcmc_p2p_connect", "ptmcmc_swap_p2p",
    "ptmcmc_p2p_error", "ptmcmc_p2p_seq",
]

_lib = None


def load():
    """Load libptmcmc_b200.so; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build the CUDA engine with `python -m ptmcmcsampler_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    h = C.c_void_p
    L.ptmcmc_abi_version.restype = C.c_int32
    L.ptmcmc_device_count.restype = C.c_int32
    L.ptmcmc_create_error.restype = C.c_char_p
    L.ptmcmc_last_error.restype = C.c_char_p
    L.ptmcmc_last_error.argtypes = [h]
    L.ptmcmc_create.restype = h
    L.ptmcmc_create.argtypes = [C.POINTER(Config)]
    L.ptmcmc_destroy.restype = None
    L.ptmcmc_destroy.argtypes = [h]
    L.ptmcmc_set_state.argtypes = [h, _dp]
    L.ptmcmc_set_state_external.argtypes = [h, _dp, _dp, _dp]
    L.ptmcmc_run.argtypes = [h, C.c_int64]
    L.ptmcmc_propose.argtypes = [h, _dp, _ip]
    L.ptmcmc_accept.argtypes = [h, _dp, _dp, _dp, _dp]
    L.ptmcmc_callback_buffers.argtypes = [h, C.POINTER(_dp), C.POINTER(_ip), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp),
                                          C.POINTER(_dp)]
    L.ptmcmc_propose_pinned.argtypes = [h, C.c_int32]
    L.ptmcmc_accept_pinned.argtypes = [h, C.c_int32]
    L.ptmcmc_iteration.restype = C.c_int64
    L.ptmcmc_iteration.argtypes = [h]
    L.ptmcmc_sync.argtypes = [h]
    L.ptmcmc_get_state.argtypes = [h, _dp, _dp, _dp, _dp]
    L.ptmcmc_rows.restype = C.c_int64
    L.ptmcmc_rows.argtypes = [h]
    L.ptmcmc_row_base.restype = C.c_int64
    L.ptmcmc_row_base.argtypes = [h]
    L.ptmcmc_get_chain.argtypes = [h, C.c_int64, C.c_int64, _dp, _dp, _dp]
    L.ptmcmc_release_rows.argtypes = [h, C.c_int64]
    L.ptmcmc_get_adapt.argtypes = [h, _dp, _dp, _dp, _i64p]
    L.ptmcmc_get_factor.argtypes = [h, _dp, _dp]
    L.ptmcmc_set_factor.argtypes = [h, _dp, _dp]
    L.ptmcmc_get_buffers.argtypes = [h, _dp, _dp]
    L.ptmcmc_adapt_begin.argtypes = [h, _dp]
    L.ptmcmc_adapt_finish.argtypes = [h, _dp]
    L.ptmcmc_njumps.argtypes = [h]
    L.ptmcmc_get_counters.argtypes = [h, _i64p, _i64p, _i64p, _i64p]
    L.ptmcmc_get_trace.argtypes = [h, C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int16), C.c_int64]
    L.ptmcmc_get_timing.argtypes = [h, C.POINTER(Timing)]
    L.ptmcmc_reset_timing.argtypes = [h]
    L.ptmcmc_stream.restype = C.c_void_p
    L.ptmcmc_stream.argtypes = [h]
    L.ptmcmc_set_timing.argtypes = [h, C.c_int32]
    L.ptmcmc_host_alloc.restype = C.c_void_p
    L.ptmcmc_host_alloc.argtypes = [C.c_int64]
    L.ptmcmc_host_free.restype = None
    L.ptmcmc_host_free.argtypes = [C.c_void_p]
    L.ptmcmc_swap_msg_doubles.restype = C.c_int64
    L.ptmcmc_swap_msg_doubles.argtypes = [h]
    L.ptmcmc_swap_pending.argtypes = [h]
    L.ptmcmc_swap_pack_top.argtypes = [h, C.c_void_p]
    L.ptmcmc_swap_sweep.argtypes = [h, C.c_void_p, C.c_void_p]
    L.ptmcmc_swap_finish.argtypes = [h, C.c_void_p]
    L.ptmcmc_p2p_open.argtypes = [h, C.c_void_p, C.POINTER(C.c_void_p)]
    L.ptmcmc_p2p_connect.argtypes = [h, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64]
    L.ptmcmc_p2p_seq.restype = C.c_int64
    L.ptmcmc_p2p_seq.argtypes = [h]
    L.ptmcmc_swap_p2p.argtypes = [h, C.c_int32]
    L.ptmcmc_p2p_error.argtypes = [h]
    L.ptmcmc_am_ring.argtypes = [h, C.POINTER(C.c_void_p), _i64p]
    L.ptmcmc_maintain.argtypes = [h]
    L.ptmcmc_state_bytes.restype = C.c_int64
    L.ptmcmc_state_bytes.argtypes = [h]
    L.ptmcmc_save_state.argtypes = [h, C.c_void_p, C.c_int64]
    L.ptmcmc_load_state.argtypes = [h, C.c_void_p, C.c_int64]
    L.ptmcmc_state_seed.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64)]
    L.ptmcmc_replay.argtypes = [h, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp]
    L.ptmcmc_mh_kernel_name.restype = C.c_char_p
    L.ptmcmc_mh_kernel_name.argtypes = [h]
    L.ptmcmc_measure_fp64_peak.argtypes = [C.c_int32, _dp]
    L.ptmcmc_user_compile_check.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int64]
    L.ptmcmc_adapt_begin_dev.argtypes = [h, C.POINTER(C.c_void_p), _i64p]
    L.ptmcmc_adapt_finish_dev.argtypes = [h, C.c_void_p, C.c_int32, C.c_int64]
    L.ptmcmc_factor_dev.argtypes = [h, C.POINTER(C.c_void_p), _i64p, C.POINTER(C.c_void_p), _i64p]
    L.ptmcmc_factor_refresh.argtypes = [h]
    L.ptmcmc_set_sink.argtypes = [h, _dp, _dp, _dp, C.c_int64]
    L.ptmcmc_sink_wait.argtypes = [h]
    L.ptmcmc_snapshot_bytes.restype = C.c_int64
    L.ptmcmc_snapshot_bytes.argtypes = [h]
    L.ptmcmc_snapshot.argtypes = [h, C.c_void_p, C.c_int64, C.c_int32]
    L.ptmcmc_snapshot_wait.argtypes = [h, C.c_int32]
    L.ptmcmc_test_normals.argtypes = [C.c_int32, C.POINTER(C.c_uint64), C.c_int64, _dp, _dp]
    for name in SYMBOLS:
        getattr(L, name)
    if L.ptmcmc_abi_version() != ABI_VERSION:
        raise ImportError("libptmcmc_b200.so ABI %d != binding %d" % (L.ptmcmc_abi_version(), ABI_VERSION))
    _lib = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class _PinnedBlock(object):
    """One page-locked allocation; returned to a small size-keyed pool when its array dies (pinning
    memory costs ~0.4 ms per MB, more than a whole sampling step for large records)."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def release(self):
        if self.ptr is None:
            return
        if _lib is None:
            return
        pool = _pinned_pool.setdefault(self.nbytes, [])
        if sum(len(v) * k for k, v in _pinned_pool.items()) + self.nbytes <= _PINNED_POOL_LIMIT:
            pool.append(self.ptr)
        else:
            _lib.ptmcmc_host_free(self.ptr)
        self.ptr = None


_pinned_pool = {}
_PINNED_POOL_LIMIT = 2 << 30


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked memory from the engine library (falls back to pageable
    memory when no CUDA device is present, e.g. for host-only construction).  Contents are undefined."""
    import weakref

    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = max(count * dtype.itemsize, 1)
    n = (n + 4095) & ~4095
    pool = _pinned_pool.get(n)
    ptr = pool.pop() if pool else load().ptmcmc_host_alloc(n)
    if not ptr:
        return np.empty(shape, dtype=dtype)
    block = _PinnedBlock(ptr, n)
    buf = (C.c_char * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    weakref.finalize(buf, block.release)  # buf lives as long as any view of arr
    return arr


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "ptmcmc engine error %d: %s" % (code, msg))
        self.code = code


class Engine(object):
    """Thin object wrapper over one ptmcmc_engine handle.  Arrays are [T][W][d] numpy float64."""

    def __init__(self, ndim, nwalkers, ntemps, cov, ladder, mh_temp=None, seed=0, groups=None,
                 cycle=((JUMP_SCAM, 20), (JUMP_AM, 20)), de_weight=20, cov_update=1000, burn=10000, tskip=100,
                 thin=10, logl_kind=LOGL_GAUSSIAN, logl_params=None, logp_kind=LOGP_UNIFORM, logp_params=None,
                 record_hot=False, record_rows=1024, trace_iters=0, timing=False, device=0, walker_offset=0,
                 temp_offset=0, ntemps_global=0, ladder_above=0.0, ladder_below=0.0, logl_source=None, logp_source=None,
                 logl_user_params=None, logp_user_params=None):
        L = load()
        self._L = L
        self.d, self.W, self.T = int(ndim), int(nwalkers), int(ntemps)
        self.cov_update, self.burn, self.thin = int(cov_update), int(burn), int(thin)
        cfg = Config()
        cfg.abi_version, cfg.device = ABI_VERSION, int(device)
        cfg.ndim, cfg.nwalkers, cfg.ntemps = self.d, self.W, self.T
        cfg.walker_offset, cfg.temp_offset, cfg.seed = int(walker_offset), int(temp_offset), int(seed) & (2**64 - 1)
        keep = []
        ladder = np.ascontiguousarray(ladder, dtype=np.float64)
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        assert ladder.shape == (self.T,) and cov.shape == (self.d, self.d)
        keep += [ladder, cov]
        cfg.ladder, cfg.cov = _d(ladder), _d(cov)
        if mh_temp is not None:
            mh_temp = np.ascontiguousarray(mh_temp, dtype=np.float64)
            keep.append(mh_temp)
            cfg.mh_temp = _d(mh_temp)
        if groups is None:
            self.groups = [np.arange(self.d, dtype=np.int32)]
            cfg.ngroups = 0
        else:
            self.groups = [np.asarray(g, dtype=np.int32) for g in groups]
            offs = np.zeros(len(self.groups) + 1, dtype=np.int32)
            offs[1:] = np.cumsum([len(g) for g in self.groups])
            idx = np.ascontiguousarray(np.concatenate(self.groups), dtype=np.int32)
            keep += [offs, idx]
            cfg.ngroups, cfg.group_offsets, cfg.group_indices = len(self.groups), _i(offs), _i(idx)
        cycle = [(int(j), int(w)) for j, w in cycle]
        if len(cycle) >= MAX_CYCLE:
            raise ValueError("at most %d proposal-cycle segments" % (MAX_CYCLE - 1))
        cfg.ncycle, cfg.de_weight = len(cycle), int(de_weight)
        for i, (j, w) in enumerate(cycle):
            cfg.cycle_jump[i], cfg.cycle_weight[i] = j, w
        cfg.cov_update, cfg.burn, cfg.tskip, cfg.thin = self.cov_update, self.burn, int(tskip), self.thin
        cfg.logl_kind, cfg.logp_kind = int(logl_kind), int(logp_kind)
        if logl_params is not None:
            lpar = np.ascontiguousarray(logl_params, dtype=np.float64)
            keep.append(lpar)
            cfg.logl_params = _d(lpar)
        if logp_params is not None:
            ppar = np.ascontiguousarray(logp_params, dtype=np.float64)
            keep.append(ppar)
            cfg.logp_params = _d(ppar)
        if logl_source is not None:
            cfg.logl_source = logl_source.encode() if isinstance(logl_source, str) else logl_source
        if logp_source is not None:
            cfg.logp_source = logp_source.encode() if isinstance(logp_source, str) else logp_source
        upar = np.concatenate([np.asarray(a, dtype=np.float64).ravel() if a is not None else np.zeros(0)
                               for a in (logl_user_params, logp_user_params)])
        cfg.n_logl_user_params = 0 if logl_user_params is None else int(np.size(logl_user_params))
        cfg.n_logp_user_params = 0 if logp_user_params is None else int(np.size(logp_user_params))
        if upar.size:
            upar = np.ascontiguousarray(upar)
            keep.append(upar)
            cfg.user_params = _d(upar)
        cfg.record_hot, cfg.record_rows = int(bool(record_hot)), int(record_rows)
        cfg.trace, cfg.trace_iters = int(trace_iters > 0), int(trace_iters)
        cfg.timing = int(bool(timing))
        cfg.ntemps_global, cfg.ladder_above, cfg.ladder_below = int(ntemps_global), float(ladder_above), float(ladder_below)
        self.temp_offset, self.ntemps_global = int(temp_offset), int(ntemps_global) or self.T
        self.device = int(device)
        self.ntr = self.T if record_hot else 1
        self.usize = sum(len(g) ** 2 for g in self.groups)
        self.ssize = sum(len(g) for g in self.groups)
        self._h = L.ptmcmc_create(C.byref(cfg))
        if not self._h:
            msg = L.ptmcmc_create_error().decode()
            if "No jump proposals" in msg or "NVRTC" in msg:
                raise ValueError(msg)
            raise EngineError(ERR_CUDA, msg)
        self.njumps = L.ptmcmc_njumps(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.ptmcmc_destroy(self._h)
            self._h = None
        self._sink = None
        self._snap = None
        self._cb = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            msg = self._L.ptmcmc_last_error(self._h).decode()
            if rc == ERR_DE_SHAPE:
                raise ValueError(msg)
            raise EngineError(rc, msg)
        return rc

    def _full(self, x0):
        return np.ascontiguousarray(np.broadcast_to(np.asarray(x0, dtype=np.float64), (self.T, self.W, self.d)))

    def set_state(self, x0):
        self._check(self._L.ptmcmc_set_state(self._h, _d(self._full(x0))))

    def set_state_external(self, x0, lnl, lnprior):
        lnl = np.ascontiguousarray(np.broadcast_to(lnl, (self.T, self.W)), dtype=np.float64)
        lnprior = np.ascontiguousarray(np.broadcast_to(lnprior, (self.T, self.W)), dtype=np.float64)
        self._check(self._L.ptmcmc_set_state_external(self._h, _d(self._full(x0)), _d(lnl), _d(lnprior)))

    def run(self, niter):
        self._check(self._L.ptmcmc_run(self._h, int(niter)))

    def propose(self):
        q = np.empty((self.T, self.W, self.d))
        jump = np.empty((self.T, self.W), dtype=np.int32)
        self._check(self._L.ptmcmc_propose(self._h, _d(q), _i(jump)))
        return q, jump

    def accept(self, q, qxy, lnl, lnprior):
        q = np.ascontiguousarray(q, dtype=np.float64)
        qxy = np.ascontiguousarray(qxy, dtype=np.float64)
        lnl = np.ascontiguousarray(lnl, dtype=np.float64)
        lnprior = np.ascontiguousarray(lnprior, dtype=np.float64)
        self._check(self._L.ptmcmc_accept(self._h, _d(q), _d(qxy), _d(lnl), _d(lnprior)))

    def callback_buffers(self):
        """numpy views of the engine's page-locked round-trip buffers: dict(q, jump, qxy, lnl, lp, x)."""
        if getattr(self, "_cb", None) is None:
            q, qxy, lnl, lp, x = _dp(), _dp(), _dp(), _dp(), _dp()
            jump = _ip()
            self._check(self._L.ptmcmc_callback_buffers(self._h, C.byref(q), C.byref(jump), C.byref(qxy), C.byref(lnl),
                                                        C.byref(lp), C.byref(x)))
            Cn = self.T * self.W
            as_arr = lambda ptr, n, shp: np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shp)  # noqa: E731
            self._cb = dict(q=as_arr(q, Cn * self.d, (self.T, self.W, self.d)), x=as_arr(x, Cn * self.d, (self.T, self.W, self.d)),
                            jump=as_arr(jump, Cn, (self.T, self.W)), qxy=as_arr(qxy, Cn, (self.T, self.W)),
                            lnl=as_arr(lnl, Cn, (self.T, self.W)), lp=as_arr(lp, Cn, (self.T, self.W)))
        return self._cb

    def propose_pinned(self, want_x=False):
        self._check(self._L.ptmcmc_propose_pinned(self._h, int(bool(want_x))))

    def accept_pinned(self, q_modified=True):
        self._check(self._L.ptmcmc_accept_pinned(self._h, int(bool(q_modified))))

    @property
    def iteration(self):
        return int(self._L.ptmcmc_iteration(self._h))

    def sync(self):
        self._check(self._L.ptmcmc_sync(self._h))

    def state(self):
        x = np.empty((self.T, self.W, self.d))
        lnl, lp, lnp = (np.empty((self.T, self.W)) for _ in range(3))
        self._check(self._L.ptmcmc_get_state(self._h, _d(x), _d(lnl), _d(lp), _d(lnp)))
        return x, lnl, lp, lnp

    @property
    def rows(self):
        return int(self._L.ptmcmc_rows(self._h))

    @property
    def row_base(self):
        return int(self._L.ptmcmc_row_base(self._h))

    def chain(self, row0=None, nrows=None, out=None):
        row0 = self.row_base if row0 is None else row0
        nrows = self.rows - row0 if nrows is None else nrows
        if out is None:
            ch = np.empty((nrows, self.ntr, self.W, self.d))
            lnl, lnp = np.empty((nrows, self.ntr, self.W)), np.empty((nrows, self.ntr, self.W))
        else:
            ch, lnl, lnp = out
            for a in out:
                assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
        self._check(self._L.ptmcmc_get_chain(self._h, row0, nrows, _d(ch), _d(lnl), _d(lnp)))
        return ch, lnl, lnp

    def release_rows(self, upto_row):
        self._check(self._L.ptmcmc_release_rows(self._h, int(upto_row)))

    def adapt(self):
        cov, mu, m2 = np.empty((self.d, self.d)), np.empty(self.d), np.empty((self.d, self.d))
        n = C.c_int64()
        self._check(self._L.ptmcmc_get_adapt(self._h, _d(cov), _d(mu), _d(m2), C.byref(n)))
        return cov, mu, m2, n.value

    def factor(self):
        U, S = np.empty(self.usize), np.empty(self.ssize)
        self._check(self._L.ptmcmc_get_factor(self._h, _d(U), _d(S)))
        return U, S

    def set_factor(self, U, S):
        U = np.ascontiguousarray(U, dtype=np.float64).ravel()
        S = np.ascontiguousarray(S, dtype=np.float64).ravel()
        assert U.size == self.usize and S.size == self.ssize
        self._check(self._L.ptmcmc_set_factor(self._h, _d(U), _d(S)))

    def buffers(self):
        am = np.empty((self.cov_update, self.W, self.d))
        de = np.empty((self.burn, self.W, self.d))
        self._check(self._L.ptmcmc_get_buffers(self._h, _d(am), _d(de)))
        return am, de

    def adapt_begin(self):
        batch = np.empty(1 + self.d + self.d * self.d)
        rc = self._check(self._L.ptmcmc_adapt_begin(self._h, _d(batch)))
        return batch if rc == 1 else None

    def adapt_finish(self, batch):
        batch = np.ascontiguousarray(batch, dtype=np.float64)
        self._check(self._L.ptmcmc_adapt_finish(self._h, _d(batch)))

    # ---- device-resident forms (collectives on device memory, no host synchronisation) ------------------
    def adapt_begin_dev(self):
        """(due, device address, doubles) of this engine's batch moments; filled in stream order when due."""
        ptr, n = C.c_void_p(), C.c_int64()
        rc = self._check(self._L.ptmcmc_adapt_begin_dev(self._h, C.byref(ptr), C.byref(n)))
        return rc == 1, int(ptr.value), int(n.value)

    def adapt_finish_dev(self, parts_ptr, nparts, nsamples):
        self._check(self._L.ptmcmc_adapt_finish_dev(self._h, parts_ptr, int(nparts), int(nsamples)))

    def factor_dev(self):
        """((U device address, doubles), (S device address, doubles))."""
        pu, ps, nu, ns = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        self._check(self._L.ptmcmc_factor_dev(self._h, C.byref(pu), C.byref(nu), C.byref(ps), C.byref(ns)))
        return (int(pu.value), int(nu.value)), (int(ps.value), int(ns.value))

    def factor_refresh(self):
        self._check(self._L.ptmcmc_factor_refresh(self._h))

    def counters(self):
        """(proposed[T][W][njumps], accepted[T][W][njumps], swap_accepted[T][W], swapProposed); the first
        two are views of the engine's [njumps][T][W] layout."""
        shp = (self.njumps, self.T, self.W)
        prop, acc = np.empty(shp, dtype=np.int64), np.empty(shp, dtype=np.int64)
        sw = np.empty((self.T, self.W), dtype=np.int64)
        n = C.c_int64()
        p64 = lambda a: a.ctypes.data_as(_i64p)  # noqa: E731
        self._check(self._L.ptmcmc_get_counters(self._h, p64(prop), p64(acc), p64(sw), C.byref(n)))
        return prop.transpose(1, 2, 0), acc.transpose(1, 2, 0), sw, n.value

    def trace(self, iters, events=0):
        tr = np.empty((iters, self.T, self.W), dtype=np.uint8)
        sm = np.empty((events, self.W, self.T), dtype=np.int16)
        self._check(self._L.ptmcmc_get_trace(self._h, tr.ctypes.data_as(C.POINTER(C.c_uint8)), iters,
                                             sm.ctypes.data_as(C.POINTER(C.c_int16)), events))
        return tr, sm

    # ---- ladder sharding: pointers are DEVICE addresses (e.g. torch tensor .data_ptr()) -------------
    @property
    def swap_msg_doubles(self):
        return int(self._L.ptmcmc_swap_msg_doubles(self._h))

    @property
    def swap_pending(self):
        return bool(self._L.ptmcmc_swap_pending(self._h))

    def swap_pack_top(self, msg_ptr):
        self._check(self._L.ptmcmc_swap_pack_top(self._h, msg_ptr))

    def swap_sweep(self, carry_in_ptr, carry_out_ptr):
        self._check(self._L.ptmcmc_swap_sweep(self._h, carry_in_ptr or None, carry_out_ptr or None))

    def swap_finish(self, below_ptr):
        self._check(self._L.ptmcmc_swap_finish(self._h, below_ptr or None))

    # the same exchange through peer memory: the kernels write into the neighbour's mailbox themselves
    def p2p_open(self, want_handle=True):
        """Allocate this shard's mailbox: (64-byte CUDA IPC handle or None, device address)."""
        hd = (C.c_ubyte * 64)() if want_handle else None
        ptr = C.c_void_p()
        self._check(self._L.ptmcmc_p2p_open(self._h, hd, C.byref(ptr)))
        return (bytes(hd) if want_handle else None), ptr.value

    @property
    def p2p_seq(self):
        return int(self._L.ptmcmc_p2p_seq(self._h))

    def p2p_connect(self, above, below, seq0):
        """``above`` / ``below``: the hotter / colder neighbour's mailbox, as 64-byte IPC handles (other processes) or as
        device addresses (same process); None where there is no neighbour.  ``seq0``: the shards' common sequence number
        (at least the largest ``p2p_seq`` among them)."""
        ipc = isinstance(above if above is not None else below, (bytes, bytearray))
        if ipc:
            keep = [C.create_string_buffer(bytes(v), 64) if v is not None else None for v in (above, below)]
            args = [C.cast(k, C.c_void_p) if k is not None else None for k in keep]
        else:
            args = [C.c_void_p(v) if v is not None else None for v in (above, below)]
        self._check(self._L.ptmcmc_p2p_connect(self._h, args[0], args[1], 1 if ipc else 0, int(seq0)))

    def swap_p2p(self, phase):
        self._check(self._L.ptmcmc_swap_p2p(self._h, phase))

    def p2p_error(self):
        self._check(self._L.ptmcmc_p2p_error(self._h))

    def am_ring(self):
        """(device address, number of doubles) of the AM ring."""
        ptr, n = C.c_void_p(), C.c_int64()
        self._check(self._L.ptmcmc_am_ring(self._h, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)

    def maintain(self):
        """Run the covariance / DE maintenance due at the start of the next iteration now."""
        self._check(self._L.ptmcmc_maintain(self._h))

    # ---- checkpoint / resume ---------------------------------------------------------------------
    def save_state(self):
        """Complete sampling state as a uint8 array (see ptmcmc_save_state)."""
        n = int(self._L.ptmcmc_state_bytes(self._h))
        buf = np.empty(n, dtype=np.uint8)
        self._check(self._L.ptmcmc_save_state(self._h, buf.ctypes.data, n))
        return buf

    def load_state(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._check(self._L.ptmcmc_load_state(self._h, buf.ctypes.data, buf.size))

    def replay(self, niter, repeat, x, lnl, lnprior):
        """Advance ``niter`` iterations taking the states from stored rows (reference-style resume)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        nrows = x.shape[0]
        assert x.shape == (nrows, self.T, self.W, self.d)
        lnl = np.ascontiguousarray(np.broadcast_to(lnl, (nrows, self.T, self.W)), dtype=np.float64)
        lnprior = np.ascontiguousarray(np.broadcast_to(lnprior, (nrows, self.T, self.W)), dtype=np.float64)
        self._check(self._L.ptmcmc_replay(self._h, int(niter), int(repeat), nrows, _d(x), _d(lnl), _d(lnprior)))

    # ---- asynchronous record sink and write-time snapshots --------------------------------------------
    def set_sink(self, chain, lnl, lnprob):
        """Stream recorded rows into the page-locked arrays chain [rows][ntr][W][d], lnl / lnprob [rows][ntr][W]
        (kept alive by this object until ``clear_sink`` / ``close``)."""
        for a in (chain, lnl, lnprob):
            assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
        rows = chain.shape[0]
        assert chain.size == rows * self.ntr * self.W * self.d and lnl.size == rows * self.ntr * self.W == lnprob.size
        self._sink = (chain, lnl, lnprob)
        self._check(self._L.ptmcmc_set_sink(self._h, _d(chain), _d(lnl), _d(lnprob), rows))

    def clear_sink(self):
        if getattr(self, "_h", None):
            self._L.ptmcmc_set_sink(self._h, None, None, None, 0)
        self._sink = None

    def sink_wait(self):
        self._check(self._L.ptmcmc_sink_wait(self._h))

    def snapshot(self, slot):
        """Enqueue a snapshot into slot 0 / 1; ``snapshot_result(slot)`` waits for it and decodes it."""
        if getattr(self, "_snap", None) is None:
            n = int(self._L.ptmcmc_snapshot_bytes(self._h))
            self._snap = [pinned_empty((n // 8,), dtype=np.int64) for _ in range(2)]
        buf = self._snap[slot]
        self._check(self._L.ptmcmc_snapshot(self._h, buf.ctypes.data, buf.nbytes, slot))

    def snapshot_result(self, slot):
        self._check(self._L.ptmcmc_snapshot_wait(self._h, slot))
        buf = self._snap[slot]
        nj, T, d = self.njumps, self.T, self.d
        it, nsw, nsamp = int(buf[0]), int(buf[1]), int(buf[2])
        n = nj * T
        summ = buf[4:4 + 4 * n + 2 * T]
        out = dict(iteration=it, swap_proposed=nsw, nsamp=nsamp,
                   prop_sum=summ[0:n].reshape(nj, T).copy(), acc_sum=summ[n:2 * n].reshape(nj, T).copy(),
                   prop_w0=summ[2 * n:3 * n].reshape(nj, T).copy(), acc_w0=summ[3 * n:4 * n].reshape(nj, T).copy(),
                   swap_sum=summ[4 * n:4 * n + T].copy(), swap_w0=summ[4 * n + T:4 * n + 2 * T].copy())
        dbl = buf[4 + 4 * n + 2 * T:].view(np.float64)
        o = 0
        for name, cnt, shp in (("cov", d * d, (d, d)), ("mu", d, (d,)), ("m2", d * d, (d, d)), ("U", self.usize, (-1,)),
                               ("S", self.ssize, (-1,))):
            out[name] = dbl[o:o + cnt].reshape(shp).copy()
            o += cnt
        return out

    def timing(self):
        t = Timing()
        self._check(self._L.ptmcmc_get_timing(self._h, C.byref(t)))
        return dict(launches={K_NAMES[i]: int(t.launches[i]) for i in range(8)},
                    ms={K_NAMES[i]: float(t.ms[i]) for i in range(8)}, chain_steps=int(t.chain_steps))

    def reset_timing(self):
        self._check(self._L.ptmcmc_reset_timing(self._h))

    def set_timing(self, on):
        self._check(self._L.ptmcmc_set_timing(self._h, int(bool(on))))

    @property
    def stream(self):
        return self._L.ptmcmc_stream(self._h)

    @property
    def mh_kernel_name(self):
        return self._L.ptmcmc_mh_kernel_name(self._h).decode()


def checkpoint_seed(buf):
    """Seed recorded in an engine checkpoint (``Engine.save_state`` bytes); None if ``buf`` is not one."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    seed = C.c_uint64()
    rc = load().ptmcmc_state_seed(buf.ctypes.data, buf.size, C.byref(seed))
    return int(seed.value) if rc == 0 else None


def user_compile_check(logl_source=None, logp_source=None, cc=(10, 0)):
    """Compile user target sources with NVRTC for compute capability ``cc`` (no device needed).  Returns the cubin size;
    raises ``ValueError`` with the compiler log on failure."""
    log = C.create_string_buffer(1 << 16)
    enc = lambda s: None if s is None else (s.encode() if isinstance(s, str) else s)  # noqa: E731
    rc = load().ptmcmc_user_compile_check(enc(logl_source), enc(logp_source), int(cc[0]), int(cc[1]), log, len(log))
    if rc < 0:
        raise ValueError(log.value.decode(errors="replace"))
    return rc


def measure_fp64_peak(device=0):
    """Measured fp64 FMA throughput of the device in TFLOP/s."""
    v = C.c_double()
    rc = load().ptmcmc_measure_fp64_peak(int(device), C.byref(v))
    if rc < 0:
        raise EngineError(rc, "ptmcmc_measure_fp64_peak failed")
    return v.value


def device_normals(words, device=0):
    """Box-Muller pairs of 64-bit words as the device computes them (test hook)."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    z0, z1 = np.empty(words.size), np.empty(words.size)
    rc = load().ptmcmc_test_normals(int(device), words.ctypes.data_as(C.POINTER(C.c_uint64)), words.size, _d(z0), _d(z1))
    if rc < 0:
        raise EngineError(rc, "ptmcmc_test_normals failed")
    return z0, z1
