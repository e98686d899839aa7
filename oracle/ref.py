"""ctypes loader for oracle/_ref/libsl_ref.so: the reference's own src/infer + src/db sources, compiled
unmodified from /root/reference against the stand-in headers of oracle/ref_shim/ (oracle/Makefile.ref,
oracle/ref_harness.cpp say exactly what is reference code and what is a stand-in).

TEST INFRASTRUCTURE ONLY.  Used by tests/test_ref_pin.py and tests/golden/make_ref_golden.py to pin the
oracle (oracle/sl_oracle.cpp) against the reference's code.  /root/reference does not exist on the GPU box:
the library is built here and travels as a built file; nothing builds it at test time unless the sources are
present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as slo

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsl_ref.so")
REFERENCE_ROOT = os.environ.get("STATELINE_REFERENCE", "/root/reference")

_dp = C.POINTER(C.c_double)


class SlrefConfig(C.Structure):
    _fields_ = [
        ("n_stacks", C.c_uint32), ("n_temps", C.c_uint32), ("n_dims", C.c_uint32),
        ("n_job_types", C.c_uint32), ("swap_interval", C.c_uint32),
        ("opt_accept", C.c_double), ("opt_swap", C.c_double),
        ("min", _dp), ("max", _dp),
        ("seed", C.c_uint32), ("wire", C.c_int32),
        ("output_path", C.c_char_p),
        ("lik", C.c_void_p), ("lik_ctx", C.c_void_p),
        ("target_fn", C.c_void_p), ("target", C.c_int32), ("target_params", _dp),
        ("max_rounds", C.c_uint64),
    ]


def reference_present():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "src", "infer", "sampler.cpp"))


def build(force=False):
    """Builds oracle/_ref when the reference sources are present (this container); otherwise the prebuilt
    library is used as it is.  Returns the library path or None when neither exists."""
    if reference_present():
        args = ["make", "-C", _HERE, "-f", "Makefile.ref", "REF=" + REFERENCE_ROOT]
        if force:
            args.append("-B")
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    if path is None:
        raise RuntimeError("oracle/_ref/libsl_ref.so is not built and the reference sources are absent")
    L = C.CDLL(path)
    L.slref_create.restype = C.c_void_p
    L.slref_create.argtypes = [C.POINTER(SlrefConfig), _dp]
    L.slref_step.argtypes = [C.c_void_p, C.c_uint64]
    L.slref_flush.argtypes = [C.c_void_p]
    L.slref_destroy.argtypes = [C.c_void_p]
    L.slref_unattributed.restype = C.c_uint64
    L.slref_unattributed.argtypes = [C.c_void_p]
    L.slref_get_draws.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, _dp]
    L.slref_get_last_rows.argtypes = [C.c_void_p, _dp]
    L.slref_get_sigma_beta.argtypes = [C.c_void_p, _dp, _dp]
    L.slref_get_shaper.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp, _dp]
    L.slref_get_models.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    L.slref_cold_rows.restype = C.c_uint64
    L.slref_cold_rows.argtypes = [C.c_void_p, C.c_uint32, _dp, C.c_uint64]
    L.slref_length.restype = C.c_uint32
    L.slref_length.argtypes = [C.c_void_p, C.c_uint32]
    L.slref_step_print.restype = C.c_size_t
    L.slref_step_print.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.slref_get_logger.argtypes = [C.c_void_p] + [_dp] * 7
    L.slref_bouncy_bounds.argtypes = [_dp, _dp, _dp, C.c_uint32, _dp]
    L.slref_epsr.argtypes = [_dp, C.c_uint32, C.c_uint32, C.c_uint32, _dp, C.POINTER(C.c_int32)]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefRun:
    """One run of the reference's Sampler over the in-process requester.

    The likelihood is the oracle's slo_target_eval (a C function pointer: no Python in the loop)."""

    def __init__(self, S, T, J, swap_interval, opt_accept, opt_swap, mn, mx, target, x_init, output_path,
                 target_params=None, seed=1, wire=0, max_rounds=0):
        self._L = lib()
        mn, mx = _f64(mn), _f64(mx)
        self.S, self.T, self.d, self.J = S, T, mn.size, J
        self.C = S * T
        x_init = _f64(x_init).reshape(self.C, self.d)
        self._keep = [mn, mx, x_init]
        cfg = SlrefConfig()
        cfg.n_stacks, cfg.n_temps, cfg.n_dims, cfg.n_job_types, cfg.swap_interval = S, T, mn.size, J, swap_interval
        cfg.opt_accept, cfg.opt_swap = opt_accept, opt_swap
        cfg.min, cfg.max = _p(mn), _p(mx)
        cfg.seed, cfg.wire = seed, wire
        self._path = os.fsencode(output_path)
        cfg.output_path = self._path
        cfg.lik, cfg.lik_ctx = None, None
        cfg.target_fn = C.cast(slo.lib(libm=True).slo_target_eval, C.c_void_p)
        cfg.target = target
        if target_params is not None:
            tp = _f64(target_params)
            self._keep.append(tp)
            cfg.target_params = _p(tp)
        cfg.max_rounds = max_rounds
        self.max_rounds = max_rounds
        self.cfg = cfg
        self.h = self._L.slref_create(C.byref(cfg), _p(x_init))
        if not self.h:
            raise RuntimeError("slref_create failed (one reference run at a time)")

    def close(self):
        """Destroys the sampler and the chain array: the ChainArray destructor writes the CSV files."""
        if self.h:
            self._L.slref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def run_rounds(self, n):
        """n canonical rounds = n * S*T calls of Sampler::step()."""
        self._L.slref_step(self.h, n * self.C)

    def step_print(self):
        """One more Sampler::step() on which TableLogger prints (logging.cpp:50-87); returns the captured text."""
        buf = C.create_string_buffer(1 << 20)
        self._L.slref_step_print(self.h, buf, len(buf))
        return buf.value.decode()

    def logger(self):
        """TableLogger's counters and EPSRDiagnostic::rHat()."""
        a = [np.empty(self.C) for _ in range(6)]
        rhat = np.empty(self.d)
        self._L.slref_get_logger(self.h, *[_p(x) for x in a], _p(rhat))
        return dict(length=a[0], min_energy=a[1], cur_energy=a[2], accepts=a[3], swaps=a[4], swap_attempts=a[5],
                    rhat=rhat)

    def flush(self):
        self._L.slref_flush(self.h)

    def unattributed(self):
        return self._L.slref_unattributed(self.h)

    def draws(self, n_rounds):
        z = np.empty((n_rounds, self.C, self.d))
        u_mh = np.empty((n_rounds, self.C))
        u_swap = np.empty((n_rounds, self.C))
        self._L.slref_get_draws(self.h, n_rounds, _p(z), _p(u_mh), _p(u_swap))
        return z, u_mh, u_swap

    def last_rows(self):
        out = np.empty((self.C, self.d + 5))
        self._L.slref_get_last_rows(self.h, _p(out))
        return out

    def sigma_beta(self):
        s, b = np.empty(self.C), np.empty(self.C)
        self._L.slref_get_sigma_beta(self.h, _p(s), _p(b))
        return s, b

    def shaper(self, chain):
        L, N, cnt = np.empty((self.d, self.d)), np.empty((self.d, self.d)), np.empty(1)
        self._L.slref_get_shaper(self.h, chain, _p(L), _p(N), _p(cnt))
        return L, N, float(cnt[0])

    def models(self, which):
        T = self.T
        xx, xy, w, cnt = np.empty((T, 9)), np.empty((T, 3)), np.empty((T, 3)), np.empty(T)
        values, rates = np.empty(self.C), np.empty(self.C)
        self._L.slref_get_models(self.h, which, _p(xx), _p(xy), _p(w), _p(cnt), _p(values), _p(rates))
        return xx, xy, w, cnt, values, rates

    def cold_rows(self, stack):
        n = self._L.slref_cold_rows(self.h, stack, None, 0)
        out = np.empty((n, self.d + 5))
        self._L.slref_cold_rows(self.h, stack, _p(out), n)
        return out

    def length(self, chain):
        return self._L.slref_length(self.h, chain)


def bouncy_bounds(val, mn, mx):
    """bouncyBounds of src/infer/sampler.cpp:23-56, the reference's own object code."""
    val, mn, mx = _f64(val), _f64(mn), _f64(mx)
    out = np.empty_like(val)
    lib().slref_bouncy_bounds(_p(val), _p(mn), _p(mx), val.size, _p(out))
    return out


def epsr(samples):
    """EPSRDiagnostic of src/infer/diagnostics.cpp (object code): samples[n][m][d] -> (rhat[d], hasConverged)."""
    x = _f64(samples)
    n, m, d = x.shape
    out = np.empty(d)
    conv = C.c_int32(0)
    lib().slref_epsr(_p(x), n, m, d, _p(out), C.byref(conv))
    return out, bool(conv.value)
