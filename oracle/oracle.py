"""ctypes loader for the CPU oracle (oracle/sl_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by stateline_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsl_oracle.so")

TARGET_DEMO_GAUSS, TARGET_DOUBLE_WELL, TARGET_GAUSS_FACT, TARGET_GAUSS_CORR, TARGET_ROSENBROCK = range(5)
SCHED_SEQUENTIAL, SCHED_BATCHED, SCHED_PER_STACK = range(3)
RNG_REPLAY, RNG_PHILOX = 0, 1

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


class SloConfig(C.Structure):
    _fields_ = [
        ("n_stacks", C.c_uint32), ("n_temps", C.c_uint32), ("n_dims", C.c_uint32),
        ("n_job_types", C.c_uint32), ("swap_interval", C.c_uint32), ("adapt_interval", C.c_uint32),
        ("opt_accept", C.c_double), ("opt_swap", C.c_double),
        ("min", _dp), ("max", _dp),
        ("use_initial", C.c_int32), ("initial", _dp),
        ("target", C.c_int32), ("target_params", _dp),
        ("schedule", C.c_int32), ("rng_mode", C.c_int32),
        ("seed", C.c_uint64), ("stack_offset", C.c_uint32), ("wire_quantize", C.c_int32),
    ]


def build(force=False):
    src = os.path.join(_HERE, "sl_oracle.cpp")
    hdr = os.path.join(_HERE, "sl_oracle.h")
    mth = os.path.join(os.path.dirname(_HERE), "include", "slgpu_math.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(mth))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libsl_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_LIBM_PATH = os.path.join(_HERE, "libsl_oracle_libm.so")
_lib = None
_lib_libm = None


def build_libm(force=False):
    """The oracle with glibc exp/log (-DSLO_LIBM): compared bit for bit with oracle/_ref (oracle/ref.py)."""
    src = os.path.join(_HERE, "sl_oracle.cpp")
    hdr = os.path.join(_HERE, "sl_oracle.h")
    if (not force and os.path.exists(_LIBM_PATH)
            and os.path.getmtime(_LIBM_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIBM_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libsl_oracle_libm.so"], stdout=subprocess.DEVNULL)
    return _LIBM_PATH


def lib(libm=False):
    global _lib, _lib_libm
    if libm:
        if _lib_libm is None:
            build_libm()
            _lib_libm = _declare(C.CDLL(_LIBM_PATH))
        return _lib_libm
    if _lib is not None:
        return _lib
    build()
    _lib = _declare(C.CDLL(_LIB_PATH))
    return _lib


def _declare(L):
    L.slo_create.restype = C.c_void_p
    L.slo_create.argtypes = [C.POINTER(SloConfig)]
    L.slo_destroy.argtypes = [C.c_void_p]
    L.slo_set_replay.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_uint64]
    L.slo_init_chains.argtypes = [C.c_void_p]
    L.slo_run_rounds.argtypes = [C.c_void_p, C.c_uint64]
    L.slo_flush.argtypes = [C.c_void_p]
    L.slo_rounds_done.restype = C.c_uint64
    L.slo_rounds_done.argtypes = [C.c_void_p]
    L.slo_get_last_rows.argtypes = [C.c_void_p, _dp]
    L.slo_get_sigma_beta.argtypes = [C.c_void_p, _dp, _dp]
    L.slo_get_shaper.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp, _dp]
    L.slo_n_models.restype = C.c_uint32
    L.slo_n_models.argtypes = [C.c_void_p]
    L.slo_get_models.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    L.slo_get_counters.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, _dp]
    L.slo_get_logger.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.slo_window_rates.argtypes = [C.POINTER(C.c_int), C.c_uint32, _dp]
    L.slo_format_table.restype = C.c_size_t
    L.slo_format_table.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.slo_cold_rows.restype = C.c_uint64
    L.slo_cold_rows.argtypes = [C.c_void_p, C.c_uint32, _dp, C.c_uint64]
    L.slo_rhat.argtypes = [C.c_void_p, _dp]
    L.slo_bouncy_bounds.argtypes = [_dp, _dp, _dp, C.c_uint32, _dp]
    L.slo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.slo_draw_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, _dp]
    L.slo_draw_uniform.restype = C.c_double
    L.slo_draw_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
    L.slo_draw_init.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, _dp]
    L.slo_qr_solve3.argtypes = [_dp, _dp, _dp]
    L.slo_shaper_update.argtypes = [_dp, _dp, _dp, _dp, C.c_uint32, C.c_double]
    L.slo_epsr.argtypes = [_dp, C.c_uint32, C.c_uint32, C.c_uint32, _dp]
    L.slo_regression_init.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp, _dp]
    L.slo_regression_predict.restype = C.c_double
    L.slo_regression_predict.argtypes = [_dp, C.c_double, C.c_double, C.c_double, C.c_double]
    L.slo_regression_update.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.slo_reduce_tree256.restype = C.c_double
    L.slo_reduce_tree256.argtypes = [_dp, C.c_uint32, C.c_uint32]
    L.slo_csv_row.restype = C.c_size_t
    L.slo_csv_row.argtypes = [_dp, C.c_uint32, C.c_char_p, C.c_size_t]
    L.slo_target_eval.restype = C.c_double
    L.slo_target_eval.argtypes = [C.c_int32, _dp, C.c_uint32, C.c_uint32, C.c_uint32, _dp]
    L.slo_time_server_workers.restype = C.c_double
    L.slo_time_server_workers.argtypes = [C.POINTER(SloConfig), C.c_uint32, C.c_uint64, _u64p]
    return L


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_config(S, T, J, swap_interval, opt_accept, opt_swap, mn, mx, target, target_params=None,
                schedule=SCHED_BATCHED, rng_mode=RNG_PHILOX, seed=1234, adapt_interval=0,
                initial=None, stack_offset=0, wire_quantize=0):
    """Returns (SloConfig, keepalive list)."""
    mn = _f64(mn)
    mx = _f64(mx)
    keep = [mn, mx]
    cfg = SloConfig()
    cfg.n_stacks, cfg.n_temps, cfg.n_dims, cfg.n_job_types = S, T, mn.size, J
    cfg.swap_interval, cfg.adapt_interval = swap_interval, adapt_interval
    cfg.opt_accept, cfg.opt_swap = opt_accept, opt_swap
    cfg.min, cfg.max = _p(mn), _p(mx)
    if initial is not None:
        ini = _f64(initial)
        keep.append(ini)
        cfg.use_initial, cfg.initial = 1, _p(ini)
    cfg.target = target
    if target_params is not None:
        tp = _f64(target_params)
        keep.append(tp)
        cfg.target_params = _p(tp)
    cfg.schedule, cfg.rng_mode, cfg.seed = schedule, rng_mode, seed
    cfg.stack_offset, cfg.wire_quantize = stack_offset, wire_quantize
    return cfg, keep


class Sampler:
    """Thin object wrapper over slo_sampler."""

    def __init__(self, libm=False, **kw):
        self._L = lib(libm)
        self.cfg, self._keep = make_config(**kw)
        self.S, self.T, self.d = self.cfg.n_stacks, self.cfg.n_temps, self.cfg.n_dims
        self.C = self.S * self.T
        self.h = self._L.slo_create(C.byref(self.cfg))

    def close(self):
        if self.h:
            self._L.slo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_replay(self, z, u_mh, u_swap, x_init):
        self._replay = [_f64(z), _f64(u_mh), _f64(u_swap), _f64(x_init)]
        n_rounds = self._replay[1].shape[0]
        self._L.slo_set_replay(self.h, *[_p(a) for a in self._replay], n_rounds)

    def init_chains(self):
        self._L.slo_init_chains(self.h)

    def run_rounds(self, n):
        self._L.slo_run_rounds(self.h, n)

    def flush(self):
        self._L.slo_flush(self.h)

    def last_rows(self):
        out = np.empty((self.C, self.d + 5))
        self._L.slo_get_last_rows(self.h, _p(out))
        return out

    def sigma_beta(self):
        s = np.empty(self.C)
        b = np.empty(self.C)
        self._L.slo_get_sigma_beta(self.h, _p(s), _p(b))
        return s, b

    def shaper(self, chain):
        L = np.empty((self.d, self.d))
        N = np.empty((self.d, self.d))
        cnt = np.empty(1)
        self._L.slo_get_shaper(self.h, chain, _p(L), _p(N), _p(cnt))
        return L, N, float(cnt[0])

    def models(self, which):
        n = self._L.slo_n_models(self.h)
        xx = np.empty((n, 9))
        xy = np.empty((n, 3))
        w = np.empty((n, 3))
        cnt = np.empty(n)
        self._L.slo_get_models(self.h, which, _p(xx), _p(xy), _p(w), _p(cnt))
        return xx, xy, w, cnt

    def counters(self):
        arrs = [np.empty(self.C, dtype=np.uint64) for _ in range(4)]
        mine = np.empty(self.C)
        self._L.slo_get_counters(self.h, *[a.ctypes.data_as(_u64p) for a in arrs], _p(mine))
        return dict(length=arrs[0], accepts=arrs[1], swaps=arrs[2], swap_attempts=arrs[3], min_energy=mine)

    def logger(self):
        cur, ar, sr = np.empty(self.C), np.empty(self.C), np.empty(self.C)
        self._L.slo_get_logger(self.h, _p(cur), _p(ar), _p(sr))
        return cur, ar, sr

    def format_table(self):
        n = self._L.slo_format_table(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._L.slo_format_table(self.h, buf, n + 1)
        return buf.value.decode()

    def cold_rows(self, stack):
        n = self._L.slo_cold_rows(self.h, stack, None, 0)
        out = np.empty((n, self.d + 5))
        self._L.slo_cold_rows(self.h, stack, _p(out), n)
        return out

    def rhat(self):
        out = np.empty(self.d)
        self._L.slo_rhat(self.h, _p(out))
        return out


# ---- stand-alone helpers -------------------------------------------------------------------------
def bouncy_bounds(val, mn, mx):
    val, mn, mx = _f64(val), _f64(mn), _f64(mx)
    out = np.empty_like(val)
    lib().slo_bouncy_bounds(_p(val), _p(mn), _p(mx), val.size, _p(out))
    return out


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().slo_philox4x32_10(c, k, o)
    return list(o)


def draw_normals(seed, chain, rnd, d):
    z = np.empty(d)
    lib().slo_draw_normals(seed, chain, rnd, d, _p(z))
    return z


def draw_uniform(seed, chain, rnd, purpose):
    return lib().slo_draw_uniform(seed, chain, rnd, purpose)


def draw_init(seed, chain, d):
    x = np.empty(d)
    lib().slo_draw_init(seed, chain, d, _p(x))
    return x


def qr_solve3(A, b):
    A, b = _f64(A).reshape(9), _f64(b)
    w = np.empty(3)
    lib().slo_qr_solve3(_p(A), _p(b), _p(w))
    return w


def shaper_update(L, count, step, prop_norm):
    L = _f64(L).copy()
    d = L.shape[0]
    N = np.zeros_like(L)
    cnt = np.array([float(count)])
    step = _f64(step)
    lib().slo_shaper_update(_p(L), _p(N), _p(cnt), _p(step), d, prop_norm)
    return L, N, float(cnt[0])


def epsr(samples):
    s = _f64(samples)
    n, m, d = s.shape
    out = np.empty(d)
    lib().slo_epsr(_p(s), n, m, d, _p(out))
    return out


def regression_init(min_cap, max_cap):
    xx, xy, w, cnt = np.empty(9), np.empty(3), np.empty(3), np.empty(1)
    lib().slo_regression_init(min_cap, max_cap, _p(xx), _p(xy), _p(w), _p(cnt))
    return xx, xy, w, float(cnt[0])


def regression_predict(w, opt, min_cap, max_cap, side):
    w = _f64(w)
    return lib().slo_regression_predict(_p(w), opt, min_cap, max_cap, side)


def regression_update(xx, xy, w, count, min_cap, max_cap, logval, side, acc):
    xx, xy, w = _f64(xx).copy(), _f64(xy).copy(), _f64(w).copy()
    cnt = np.array([float(count)])
    lib().slo_regression_update(_p(xx), _p(xy), _p(w), _p(cnt), min_cap, max_cap, logval, side, int(acc))
    return xx, xy, w, float(cnt[0])


def reduce_tree256(v):
    v = _f64(v)
    return lib().slo_reduce_tree256(_p(v), v.size, 1)


def window_rates(flags):
    """RegressionAdapter's logging window (adaptive.cpp:80-90): the rate after each flag."""
    f = np.ascontiguousarray(flags, dtype=np.int32)
    out = np.empty(f.size)
    lib().slo_window_rates(f.ctypes.data_as(C.POINTER(C.c_int)), f.size, _p(out))
    return out


def csv_row(row, d):
    row = _f64(row)
    buf = C.create_string_buffer(64 * (d + 6))
    lib().slo_csv_row(_p(row), d, buf, len(buf))
    return buf.value.decode()


def target_eval(target, params, d, J, job, x):
    x = _f64(x)
    prm = _f64(params) if params is not None else np.zeros(1)
    return lib().slo_target_eval(target, _p(prm), d, J, job, _p(x))


def time_server_workers(n_workers, n_rounds, **kw):
    cfg, keep = make_config(**kw)
    steps = C.c_uint64(0)
    secs = lib().slo_time_server_workers(C.byref(cfg), n_workers, n_rounds, C.byref(steps))
    return secs, steps.value
