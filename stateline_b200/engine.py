"""ctypes binding of include/slgpu.h -- the Python face of the engine.

Mirrors the reference's server-side surface for the hot path: a stateline config (the keys of
StatelineSettings::fromJSON, src/app/serverwrapper.hpp:64-91) goes in, cold-chain rows in the
CSVChainArrayWriter layout (src/db/db.cpp:18-25) come out.  All compute happens in libslgpu.so and the
plugin kernels; there is no Python or CPU fallback -- a missing library or GPU raises.
"""
import ctypes as C
import json
import os

import numpy as np

from . import build as _build

OK, E_CONFIG, E_PLUGIN, E_CUDA, E_ARG, E_STATE, E_IO, E_FULL = range(8)
_STATUS_NAMES = ["OK", "E_CONFIG", "E_PLUGIN", "E_CUDA", "E_ARG", "E_STATE", "E_IO", "E_FULL"]

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)

# every symbol include/slgpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "slgpu_last_error", "slgpu_version", "slgpu_create", "slgpu_destroy", "slgpu_set_replay",
    "slgpu_init_chains", "slgpu_step_rounds", "slgpu_sync", "slgpu_run", "slgpu_rounds_done",
    "slgpu_drain", "slgpu_peek", "slgpu_advance", "slgpu_rows_emitted", "slgpu_timer_start",
    "slgpu_timer_stop", "slgpu_timer_step_kernel", "slgpu_launch_count", "slgpu_n_dims", "slgpu_n_chains", "slgpu_get_last_rows",
    "slgpu_get_sigma_beta", "slgpu_get_shaper", "slgpu_get_models", "slgpu_get_ladder",
    "slgpu_get_counters", "slgpu_rhat", "slgpu_rhat_stage1", "slgpu_rhat_stage2", "slgpu_rhat_finish",
    "slgpu_summary", "slgpu_eval_likelihood", "slgpu_format_row", "slgpu_measure_fp64_peak",
    "slgpu_save_state", "slgpu_load_state", "slgpu_get_logger", "slgpu_format_table",
    "slgpu_next_segment", "slgpu_release_segment", "slgpu_d2h_bytes", "slgpu_timer_start_coarse", "slgpu_discard_segment", "slgpu_request_key_segment", "slgpu_wait", "slgpu_group_wait",
    "slgpu_group_create", "slgpu_group_destroy", "slgpu_group_size", "slgpu_group_engine", "slgpu_group_init_chains",
    "slgpu_group_step_rounds", "slgpu_group_sync", "slgpu_group_run", "slgpu_group_rounds_done", "slgpu_group_rhat",
    "slgpu_group_summary", "slgpu_group_save_state", "slgpu_group_load_state",
]


class _Segment(C.Structure):
    _fields_ = [("n_rounds", C.c_uint32), ("n_stacks", C.c_uint32), ("n_chunks", C.c_uint32), ("payload_doubles", C.c_uint32),
                ("count", C.c_uint32), ("key", C.c_uint32), ("row_begin", C.c_uint64), ("base", C.POINTER(C.c_uint32)),
                ("flags", C.POINTER(C.c_ubyte)), ("payload", _dp)]

BUILTIN_ENTRIES = {
    "demo_gauss": "slgpu_plugin_demo_gauss",
    "double_well": "slgpu_plugin_double_well",
    "gauss_fact": "slgpu_plugin_gauss_fact",
    "gauss_corr": "slgpu_plugin_gauss_corr",
    "rosenbrock": "slgpu_plugin_rosenbrock",
}


class SlgpuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_STATUS_NAMES[status] if 0 <= status < len(_STATUS_NAMES) else status, message))
        self.status = status


_lib = None


def lib():
    """Load libslgpu.so (building it in-tree if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SLGPU_ENGINE_SO") or _build.build_engine()  # env: development A/B builds
    L = C.CDLL(path)
    L.slgpu_last_error.restype = C.c_char_p
    L.slgpu_version.restype = C.c_char_p
    L.slgpu_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.slgpu_destroy.argtypes = [C.c_void_p]
    L.slgpu_destroy.restype = None
    L.slgpu_set_replay.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_uint64]
    L.slgpu_init_chains.argtypes = [C.c_void_p]
    L.slgpu_step_rounds.argtypes = [C.c_void_p, C.c_uint64]
    L.slgpu_sync.argtypes = [C.c_void_p]
    L.slgpu_wait.argtypes = [C.c_void_p]
    L.slgpu_group_wait.argtypes = [C.c_void_p]
    L.slgpu_run.argtypes = [C.c_void_p]
    L.slgpu_rounds_done.argtypes = [C.c_void_p]
    L.slgpu_rounds_done.restype = C.c_uint64
    L.slgpu_drain.argtypes = [C.c_void_p, _dp, C.c_size_t, _szp]
    L.slgpu_peek.argtypes = [C.c_void_p, C.POINTER(_dp), _szp, C.POINTER(_dp), _szp]
    L.slgpu_advance.argtypes = [C.c_void_p, C.c_size_t]
    L.slgpu_rows_emitted.argtypes = [C.c_void_p]
    L.slgpu_rows_emitted.restype = C.c_uint64
    L.slgpu_timer_start.argtypes = [C.c_void_p]
    L.slgpu_timer_start_coarse.argtypes = [C.c_void_p]
    L.slgpu_timer_stop.argtypes = [C.c_void_p, _dp]
    L.slgpu_timer_step_kernel.argtypes = [C.c_void_p, _dp, _u64p]
    L.slgpu_launch_count.argtypes = [C.c_void_p]
    L.slgpu_launch_count.restype = C.c_uint64
    L.slgpu_n_dims.argtypes = [C.c_void_p]
    L.slgpu_n_dims.restype = C.c_uint32
    L.slgpu_n_chains.argtypes = [C.c_void_p]
    L.slgpu_n_chains.restype = C.c_uint32
    L.slgpu_get_last_rows.argtypes = [C.c_void_p, _dp]
    L.slgpu_get_sigma_beta.argtypes = [C.c_void_p, _dp, _dp]
    L.slgpu_get_shaper.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp, _dp]
    L.slgpu_get_models.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    L.slgpu_get_ladder.argtypes = [C.c_void_p, _dp]
    L.slgpu_get_counters.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, _dp]
    L.slgpu_rhat.argtypes = [C.c_void_p, _dp]
    L.slgpu_rhat_stage1.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.slgpu_rhat_stage2.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.slgpu_rhat_finish.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_uint32, _dp]
    L.slgpu_rhat_finish.restype = None
    L.slgpu_summary.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, _szp]
    L.slgpu_eval_likelihood.argtypes = [C.c_void_p, _dp, C.c_uint32, _dp]
    L.slgpu_format_row.argtypes = [_dp, C.c_uint32, C.c_char_p, C.c_size_t]
    L.slgpu_format_row.restype = C.c_size_t
    L.slgpu_measure_fp64_peak.argtypes = [C.c_int, _dp]
    L.slgpu_get_logger.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.slgpu_format_table.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, _szp]
    L.slgpu_save_state.argtypes = [C.c_void_p, C.c_char_p]
    L.slgpu_load_state.argtypes = [C.c_void_p, C.c_char_p]
    L.slgpu_next_segment.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Segment)]
    L.slgpu_release_segment.argtypes = [C.c_void_p]
    L.slgpu_discard_segment.argtypes = [C.c_void_p]
    L.slgpu_request_key_segment.argtypes = [C.c_void_p]
    L.slgpu_d2h_bytes.argtypes = [C.c_void_p]
    L.slgpu_d2h_bytes.restype = C.c_uint64
    L.slgpu_group_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_int,
                                     C.POINTER(C.c_void_p)]
    L.slgpu_group_destroy.argtypes = [C.c_void_p]
    L.slgpu_group_destroy.restype = None
    L.slgpu_group_size.argtypes = [C.c_void_p]
    L.slgpu_group_engine.argtypes = [C.c_void_p, C.c_int]
    L.slgpu_group_engine.restype = C.c_void_p
    for name in ("slgpu_group_init_chains", "slgpu_group_sync", "slgpu_group_run"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.slgpu_group_step_rounds.argtypes = [C.c_void_p, C.c_uint64]
    L.slgpu_group_rounds_done.argtypes = [C.c_void_p]
    L.slgpu_group_rounds_done.restype = C.c_uint64
    L.slgpu_group_rhat.argtypes = [C.c_void_p, _dp]
    L.slgpu_group_summary.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, _szp]
    L.slgpu_group_save_state.argtypes = [C.c_void_p, C.c_char_p]
    L.slgpu_group_load_state.argtypes = [C.c_void_p, C.c_char_p]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(status):
    if status != OK:
        raise SlgpuError(status, lib().slgpu_last_error().decode())


def make_config(nStacks, nTemperatures, mn, mx, nJobTypes, swapInterval=10, nSamplesTotal=0,
                optimalAcceptRate=0.234, optimalSwapRate=0.3874, outputPath="slgpu-output",
                loggingRateSec=1.0, initial=None, **gpu):
    """A stateline server config (src/bin/demo-config.json keys) plus the optional "gpu" block."""
    cfg = {
        "nStacks": int(nStacks), "nTemperatures": int(nTemperatures), "nSamplesTotal": int(nSamplesTotal),
        "swapInterval": int(swapInterval), "loggingRateSec": float(loggingRateSec), "nJobTypes": int(nJobTypes),
        "outputPath": outputPath, "optimalAcceptRate": float(optimalAcceptRate),
        "optimalSwapRate": float(optimalSwapRate), "useInitial": initial is not None,
        "min": [float(v) for v in mn], "max": [float(v) for v in mx],
    }
    if initial is not None:
        cfg["initial"] = [float(v) for v in initial]
    if gpu:
        cfg["gpu"] = gpu
    return cfg


def format_row(row, d):
    """One cold-chain row as CSVChainArrayWriter prints it (src/db/db.cpp:18-25)."""
    row = _f64(row)
    buf = C.create_string_buffer(64 * (d + 8))
    lib().slgpu_format_row(_p(row), d, buf, len(buf))
    return buf.value.decode()


class Engine:
    """One engine = one GPU's share of the stacks (the role ServerWrapper + Sampler play on the host)."""

    def __init__(self, config, plugin=None, entry=None, payload=None):
        """config: dict or JSON text.  plugin: path of a plugin .so, or the name of a built-in target
        (demo_gauss, double_well, gauss_fact, gauss_corr, rosenbrock) from libslgpu_targets.so."""
        L = lib()
        text = config if isinstance(config, str) else json.dumps(config)
        if plugin is None or plugin in BUILTIN_ENTRIES:
            entry = BUILTIN_ENTRIES[plugin or "demo_gauss"]
            plugin = os.environ.get("SLGPU_TARGETS_SO") or _build.build_targets()  # env: development A/B builds
        pl = None
        n = 0
        if payload is not None:
            self._payload = np.ascontiguousarray(payload)
            pl = self._payload.ctypes.data_as(C.c_void_p)
            n = self._payload.nbytes
        h = C.c_void_p()
        _check(L.slgpu_create(text.encode(), os.fspath(plugin).encode(), entry.encode() if entry else None, pl, n, C.byref(h)))
        self.h = h
        self.cfg = json.loads(text)
        self.S = self.cfg["nStacks"]
        self.T = self.cfg["nTemperatures"]
        self.d = L.slgpu_n_dims(h)
        self.C = L.slgpu_n_chains(h)
        self.J = self.cfg["nJobTypes"]

    def close(self):
        if getattr(self, "h", None):
            lib().slgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- run control
    def set_replay(self, z, u_mh, u_swap, x_init):
        z, u_mh, u_swap = _f64(z), _f64(u_mh), _f64(u_swap)
        xi = _f64(x_init) if x_init is not None else None
        _check(lib().slgpu_set_replay(self.h, _p(z), _p(u_mh), _p(u_swap), _p(xi) if xi is not None else None, u_mh.shape[0]))

    def init_chains(self):
        _check(lib().slgpu_init_chains(self.h))

    def step_rounds(self, n):
        _check(lib().slgpu_step_rounds(self.h, int(n)))

    run_rounds = step_rounds  # the oracle wrapper's name for the same thing

    def sync(self):
        _check(lib().slgpu_sync(self.h))

    def run(self):
        _check(lib().slgpu_run(self.h))

    def logger(self, rates=True):
        """TableLogger columns per chain: (cur_energy, accept_rate, swap_rate); the windowed rates need
        "gpu": {"windowRates": true} (pass rates=False to get the energies alone)."""
        cur = np.empty(self.C)
        ar = np.empty(self.C) if rates else None
        sr = np.empty(self.C) if rates else None
        _check(lib().slgpu_get_logger(self.h, _p(cur), _p(ar) if rates else None, _p(sr) if rates else None))
        return cur, ar, sr

    def format_table(self):
        """The table TableLogger::update prints (src/infer/logging.cpp:52-87) plus its convergence line."""
        need = C.c_size_t()
        _check(lib().slgpu_format_table(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().slgpu_format_table(self.h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def save_state(self, path):
        """Checkpoint every chain, the tier models and the round counter (slgpu_save_state)."""
        _check(lib().slgpu_save_state(self.h, os.fspath(path).encode()))

    def load_state(self, path):
        """Resume from a checkpoint written by an engine with the same config and plugin; replaces init_chains."""
        _check(lib().slgpu_load_state(self.h, os.fspath(path).encode()))

    @property
    def rounds_done(self):
        return lib().slgpu_rounds_done(self.h)

    @property
    def rows_emitted(self):
        return lib().slgpu_rows_emitted(self.h)

    @property
    def launch_count(self):
        return lib().slgpu_launch_count(self.h)

    def timer_start(self, per_launch=True):
        """per_launch=False: no event pair around every step-kernel launch (they keep a kernel launched with programmatic
        dependent launch from starting early): the step loop as it runs outside a measurement."""
        _check((lib().slgpu_timer_start if per_launch else lib().slgpu_timer_start_coarse)(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _check(lib().slgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def timer_step_kernel(self):
        """(summed ms, launches) of the step-kernel launches inside the last timed region."""
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().slgpu_timer_step_kernel(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- cold rows
    def drain(self, max_rows=None):
        """Landed cold-chain rows, [n][d+5], in emission order (round-major, stack-minor)."""
        out = []
        w = self.d + 5
        while True:
            cap = max_rows if max_rows is not None else max(self.S * 64, 4096)
            buf = np.empty((cap, w))
            n = C.c_size_t()
            _check(lib().slgpu_drain(self.h, _p(buf), cap, C.byref(n)))
            out.append(buf[: n.value])
            if max_rows is not None or n.value < cap:
                break
        return np.concatenate(out) if len(out) > 1 else out[0]

    def peek(self):
        """Zero-copy numpy views of the landed rows in the pinned host ring (up to two spans)."""
        p0, p1, n0, n1 = _dp(), _dp(), C.c_size_t(), C.c_size_t()
        _check(lib().slgpu_peek(self.h, C.byref(p0), C.byref(n0), C.byref(p1), C.byref(n1)))
        w = self.d + 5
        spans = []
        for p, n in ((p0, n0.value), (p1, n1.value)):
            if n:
                spans.append(np.ctypeslib.as_array(p, shape=(n, w)))
        return spans

    def advance(self, n):
        _check(lib().slgpu_advance(self.h, int(n)))

    def next_segment(self, block=False):
        """The oldest landed segment of the delta stream (sink "host", wire "delta") as zero-copy numpy views into the
        pinned ring: dict(n_rounds, row_begin, count, key, base[n_rounds, n_chunks], flags[n_rounds, S],
        payload[count, d + 3]), or None when nothing has landed.  Release it with release_segment()."""
        sg = _Segment()
        _check(lib().slgpu_next_segment(self.h, 1 if block else 0, C.byref(sg)))
        if sg.n_rounds == 0:
            return None
        as_array = np.ctypeslib.as_array
        return dict(n_rounds=sg.n_rounds, row_begin=sg.row_begin, count=sg.count, key=bool(sg.key),
                    base=as_array(sg.base, shape=(sg.n_rounds, sg.n_chunks)),
                    flags=as_array(sg.flags, shape=(sg.n_rounds, sg.n_stacks)),
                    payload=as_array(sg.payload, shape=(max(sg.count, 1), sg.payload_doubles))[:sg.count])

    def release_segment(self, keep_rows_state=True):
        """keep_rows_state=False (slgpu_discard_segment): the host stops tracking the rows behind the stream (cheaper; a later
        drain() needs a key segment first: request_key_segment())."""
        _check((lib().slgpu_release_segment if keep_rows_state else lib().slgpu_discard_segment)(self.h))

    def request_key_segment(self):
        """The next launch starts a key segment (every row of its first round in the payload): after discarded segments,
        drain() / peek() can rebuild rows again from there."""
        _check(lib().slgpu_request_key_segment(self.h))

    @property
    def d2h_bytes(self):
        return int(lib().slgpu_d2h_bytes(self.h))

    # ---- inspection (same shapes as oracle.Sampler)
    def last_rows(self):
        out = np.empty((self.C, self.d + 5))
        _check(lib().slgpu_get_last_rows(self.h, _p(out)))
        return out

    def sigma_beta(self):
        s, b = np.empty(self.C), np.empty(self.C)
        _check(lib().slgpu_get_sigma_beta(self.h, _p(s), _p(b)))
        return s, b

    def shaper(self, chain):
        L, N, cnt = np.empty((self.d, self.d)), np.empty((self.d, self.d)), C.c_double()
        _check(lib().slgpu_get_shaper(self.h, chain, _p(L), _p(N), C.byref(cnt)))
        return L, N, cnt.value

    def models(self, which):
        T = self.T
        xx, xy, w, cnt = np.empty((T, 9)), np.empty((T, 3)), np.empty((T, 3)), np.empty(T)
        _check(lib().slgpu_get_models(self.h, which, _p(xx), _p(xy), _p(w), _p(cnt)))
        return xx, xy, w, cnt

    def ladder(self):
        out = np.empty(self.T)
        _check(lib().slgpu_get_ladder(self.h, _p(out)))
        return out

    def counters(self):
        arrs = [np.empty(self.C, dtype=np.uint64) for _ in range(4)]
        mine = np.empty(self.C)
        _check(lib().slgpu_get_counters(self.h, *[a.ctypes.data_as(_u64p) for a in arrs], _p(mine)))
        return dict(length=arrs[0], accepts=arrs[1], swaps=arrs[2], swap_attempts=arrs[3], min_energy=mine)

    def rhat(self):
        out = np.empty(self.d)
        _check(lib().slgpu_rhat(self.h, _p(out)))
        return out

    def rhat_stage1(self):
        sm, ns, nn = np.empty(self.d), C.c_double(), C.c_double()
        _check(lib().slgpu_rhat_stage1(self.h, _p(sm), C.byref(ns), C.byref(nn)))
        return sm, ns.value, nn.value

    def rhat_stage2(self, mean):
        mean = _f64(mean)
        b, w = np.empty(self.d), np.empty(self.d)
        _check(lib().slgpu_rhat_stage2(self.h, _p(mean), _p(b), _p(w)))
        return b, w

    def summary(self):
        need = C.c_size_t()
        _check(lib().slgpu_summary(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().slgpu_summary(self.h, buf, need.value, None))
        return json.loads(buf.value.decode())

    def eval_likelihood(self, x):
        x = _f64(x).reshape(-1, self.d)
        out = np.empty((x.shape[0], self.J))
        _check(lib().slgpu_eval_likelihood(self.h, _p(x), x.shape[0], _p(out)))
        return out


def measure_fp64_peak(device=0):
    """Sustained FP64 FMA throughput of the device in TFLOP/s (register-resident DFMA loop)."""
    v = C.c_double()
    _check(lib().slgpu_measure_fp64_peak(int(device), C.byref(v)))
    return v.value


def rhat_finish(sum_dm2, sum_s, n_stacks_total, n_samples):
    sum_dm2, sum_s = _f64(sum_dm2), _f64(sum_s)
    out = np.empty(sum_dm2.size)
    lib().slgpu_rhat_finish(_p(sum_dm2), _p(sum_s), float(n_stacks_total), float(n_samples), sum_dm2.size, _p(out))
    return out


class _BorrowedEngine(Engine):
    """An engine owned by an EngineGroup: same methods, the group destroys it."""

    def __init__(self, handle, cfg, n_stacks):
        L = lib()
        self.h = C.c_void_p(handle)
        self.cfg = cfg
        self.S = n_stacks
        self.T = cfg["nTemperatures"]
        self.d = L.slgpu_n_dims(self.h)
        self.C = L.slgpu_n_chains(self.h)
        self.J = cfg["nJobTypes"]

    def close(self):
        self.h = None


class EngineGroup:
    """Several GPUs driven from one process through slgpu_group_* (include/slgpu.h): what one `stateline -c cfg` server
    process is on the reference side (src/bin/stateline.cpp:52-90).  devices: None = every visible device from
    gpu.device up, an int = that many, or an explicit list (a device may appear more than once)."""

    def __init__(self, config, plugin=None, entry=None, payload=None, devices=None):
        L = lib()
        text = config if isinstance(config, str) else json.dumps(config)
        if plugin is None or plugin in BUILTIN_ENTRIES:
            entry = BUILTIN_ENTRIES[plugin or "demo_gauss"]
            plugin = os.environ.get("SLGPU_TARGETS_SO") or _build.build_targets()
        pl, n = None, 0
        if payload is not None:
            self._payload = np.ascontiguousarray(payload)
            pl, n = self._payload.ctypes.data_as(C.c_void_p), self._payload.nbytes
        if devices is None or isinstance(devices, int):
            dev_ptr, n_dev = None, int(devices or 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            dev_ptr, n_dev = arr, len(devices)
        h = C.c_void_p()
        _check(L.slgpu_group_create(text.encode(), os.fspath(plugin).encode(), entry.encode() if entry else None, pl, n,
                                    dev_ptr, n_dev, C.byref(h)))
        self.h = h
        self.cfg = json.loads(text)
        self.d = len(self.cfg["min"])
        size = L.slgpu_group_size(h)
        S = self.cfg["nStacks"]
        base, rem = divmod(S, size)
        self.engines = [_BorrowedEngine(L.slgpu_group_engine(h, i), self.cfg, base + (1 if i < rem else 0)) for i in range(size)]

    def close(self):
        if getattr(self, "h", None):
            for e in self.engines:
                e.close()
            lib().slgpu_group_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def init_chains(self):
        _check(lib().slgpu_group_init_chains(self.h))

    def step_rounds(self, n):
        _check(lib().slgpu_group_step_rounds(self.h, int(n)))

    def sync(self):
        _check(lib().slgpu_group_sync(self.h))

    def run(self):
        _check(lib().slgpu_group_run(self.h))

    def save_state(self, path):
        _check(lib().slgpu_group_save_state(self.h, os.fspath(path).encode()))

    def load_state(self, path):
        _check(lib().slgpu_group_load_state(self.h, os.fspath(path).encode()))

    @property
    def rounds_done(self):
        return int(lib().slgpu_group_rounds_done(self.h))

    def rhat(self):
        out = np.empty(self.d)
        _check(lib().slgpu_group_rhat(self.h, _p(out)))
        return out

    def summary(self):
        need = C.c_size_t()
        _check(lib().slgpu_group_summary(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().slgpu_group_summary(self.h, buf, need.value, None))
        return json.loads(buf.value.decode())
