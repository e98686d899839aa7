"""CPU tests of the drop-in boundary: libslgpu.so loads, exports every symbol include/slgpu.h declares,
and the host logic in front of the device (config reader, plugin loading, error convention) behaves.
No compute call is made: there is no GPU here and the engine has no CPU fallback."""
import numpy as np
import ctypes as C
import json
import os
import re

import pytest

import stateline_b200 as sb
from stateline_b200 import configs, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sb.build.build_engine())
    names = _declared("slgpu.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(engine.SYMBOLS) == names  # the Python binding covers the whole header


def test_builtin_plugin_exports_tables():
    lib = C.CDLL(sb.build.build_targets())
    for sym in ["slgpu_plugin_entry"] + list(engine.BUILTIN_ENTRIES.values()):
        assert hasattr(lib, sym), sym


def test_version_string():
    assert b"sm_100a" in engine.lib().slgpu_version()


def _create(cfg, plugin=None, entry=None):
    h = C.c_void_p()
    L = engine.lib()
    st = L.slgpu_create(json.dumps(cfg).encode() if not isinstance(cfg, (bytes, str)) else (cfg if isinstance(cfg, bytes) else cfg.encode()),
                        (plugin or sb.build.build_targets()).encode(), entry.encode() if entry else None, None, 0, C.byref(h))
    return st, L.slgpu_last_error().decode(), h


@pytest.mark.parametrize("key", ["nStacks", "nTemperatures", "nSamplesTotal", "swapInterval", "loggingRateSec", "nJobTypes",
                                 "outputPath", "optimalAcceptRate", "optimalSwapRate", "useInitial", "min", "max"])
def test_missing_key_is_config_error(key):
    """readSettings (src/app/jsonsettings.hpp:22-35): every key of serverwrapper.hpp:64-91 is mandatory.
    The reference exits the process; the C-ABI returns SLGPU_E_CONFIG and names the key."""
    cfg = configs.engine_config(configs.workload("config1"))
    del cfg[key]
    st, msg, h = _create(cfg)
    assert st == engine.E_CONFIG and key in msg and not h


def test_initial_needed_only_with_use_initial():
    cfg = configs.engine_config(configs.workload("config1"))
    cfg["useInitial"] = True
    st, msg, _ = _create(cfg)
    assert st == engine.E_CONFIG and "initial" in msg


def test_bounds_mismatch_is_config_error():
    cfg = configs.engine_config(configs.workload("config1"))
    cfg["max"] = cfg["max"][:-1]
    st, msg, _ = _create(cfg)  # ProposalBoundsFromJSON, src/infer/sampler.hpp:42-46
    assert st == engine.E_CONFIG and "mismatch" in msg


def test_malformed_json_and_bad_values():
    st, msg, _ = _create(b"{ not json")
    assert st == engine.E_CONFIG
    cfg = configs.engine_config(configs.workload("config1"))
    cfg["nTemperatures"] = 33  # a stack's ladder must fit one warp
    assert _create(cfg)[0] == engine.E_CONFIG
    cfg = configs.engine_config(configs.workload("config1"), sink="tape")
    assert _create(cfg)[0] == engine.E_CONFIG
    st, msg, _ = _create(configs.engine_config(configs.workload("config1"), schedule="eventually"))
    assert st == engine.E_CONFIG and "schedule" in msg
    # a valid schedule gets past the parser (and stops at the missing device on a CPU-only host)
    st, msg, _ = _create(configs.engine_config(configs.workload("config1"), schedule="sequential"))
    assert st != engine.E_CONFIG


def test_plugin_errors():
    cfg = configs.engine_config(configs.workload("config1"))
    st, msg, _ = _create(cfg, plugin="/nonexistent/libnope.so")
    assert st == engine.E_PLUGIN and "dlopen" in msg
    st, msg, _ = _create(cfg, entry="no_such_symbol")
    assert st == engine.E_PLUGIN
    st, msg, _ = _create(cfg, entry="slgpu_plugin_double_well")  # fixed J=2, d=2 does not match J=3, d=4
    assert st == engine.E_PLUGIN and "does not match" in msg


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    st, msg, h = _create(configs.engine_config(configs.workload("config1")))
    assert st == engine.E_CUDA and "no CPU fallback" in msg and not h
    with pytest.raises(sb.SlgpuError):
        sb.Engine(configs.engine_config(configs.workload("config1")), plugin="demo_gauss")


def test_null_arguments():
    L = engine.lib()
    assert L.slgpu_create(None, None, None, None, 0, None) == engine.E_ARG
    assert L.slgpu_init_chains(None) == engine.E_ARG
    assert L.slgpu_rounds_done(None) == 0
    L.slgpu_destroy(None)


def test_format_row_matches_reference_stream_format():
    # operator<<(ostream&, State), src/db/db.cpp:18-25: default precision 6, accepted as bool, swapType as int
    row = [0.25, -1234567.0, 1e-7, 2.5, 0.123456789, 1.0, 1.0, 2.0]
    assert sb.format_row(row, 3) == "0.25,-1.23457e+06,1e-07,2.5,0.123457,1,1,2"
    from oracle import oracle as O
    assert sb.format_row(row, 3) == O.csv_row(row, 3)


def test_format_row_is_printf_g_for_every_kind_of_value():
    """The csv sink formats with std::to_chars(general, 6); the reference streams doubles with default flags,
    i.e. printf's %g.  Same text on magnitudes across the fixed / scientific switch, rounding ties at the sixth
    digit, subnormals, zeros and non-finite values."""
    rng = np.random.default_rng(11)
    vals = np.concatenate([
        rng.standard_normal(4000) * 10.0 ** rng.integers(-12, 13, 4000),
        rng.integers(-2_000_000, 2_000_000, 2000).astype(np.float64) / 1000.0,
        np.array([0.0, -0.0, 1e-5, 9.9999949e-5, 9.9999951e-5, 1e-4, 99999.95, 999999.5, 999999.4, 1e6, 123456.5, 0.1234565,
                  2.5e-310, 1.7976931348623157e308, 5e-324, np.inf, -np.inf, 100000.0, 1234565.0, 0.5, 1.0, -1.0, 1e21, 1e-21]),
    ])
    d = 20
    vals = vals[: (vals.size // (d + 3)) * (d + 3)].reshape(-1, d + 3)
    for i, v in enumerate(vals):
        row = np.concatenate([v, [i % 2, i % 3]])
        want = ",".join("%g" % x for x in v) + ",%d,%d" % (i % 2, i % 3)
        assert sb.format_row(row, d) == want


def test_build_digest_does_not_depend_on_the_checkout_path(tmp_path, monkeypatch):
    """The source hash that decides whether an in-tree .so is stale (and that ties profiles/pt_step_traffic.json to the
    kernels it was measured on) must be the same wherever the tree is copied: the GPU boxes run from a scratch path."""
    from stateline_b200 import build as B
    src = os.path.join(B.PKG, "plugins", "builtin_targets.cu")
    flags = B.ARCH + B.COMMON + ["-O3"]
    here = B._digest([src] + B._plugin_headers(), flags)
    monkeypatch.setattr(B, "ROOT", "/somewhere/else")
    moved_flags = [f.replace(ROOT, "/somewhere/else") for f in flags]
    assert B._digest([src] + B._plugin_headers(), moved_flags) == here
    assert not any(h.endswith("slgpu.h") for h in B._plugin_headers())  # the host API is no part of a plugin build
    if os.path.exists(B.TARGETS_SO + ".srchash"):
        traffic = json.load(open(os.path.join(ROOT, "profiles", "pt_step_traffic.json")))
        assert traffic["plugin_digest"] == open(B.TARGETS_SO + ".srchash").read().strip()[:16], \
            "profiles/pt_step_traffic.json was measured on other kernel sources: re-capture it (tools/make_profiles_r2.py)"
