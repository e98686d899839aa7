"""The reference's own schedule on the GPU ("gpu": {"schedule": "sequential"}) -- north-star check 1.

In this validation mode every launch advances the chains by one round and k_tier_scan_sequential then replays
Sampler::step()'s updates of the shared tier models chain by chain in the reference's FIFO order (descending chain id:
sampler.cpp:128-132,237-257), including the cascade's quirk that only beta[1] takes the new ladder at once.

Two checks:
  * GPU == oracle (SLO_SCHED_SEQUENTIAL) bit for bit, native Philox draws and recorded draws;
  * the committed REFERENCE fixtures (tests/golden/ref_*.json: runs of the reference's own src/infer object code with
    the variates it drew, tests/golden/make_ref_golden.py) replayed on the GPU: identical accept / swap decisions and
    logger counters; sigma / beta within 1e-12 relative; samples and energies of the cold chains within 1e-12 of
    max(|v|, 1) (1e-10 on the double well), of all chains within 1e-9 (the device uses include/slgpu_math.h for exp / log,
    the reference glibc: at most 1 ulp apart per call; tests/test_ref_pin.py states the per-tier measurements).
"""
import json
import os

import numpy as np
import pytest

from helpers import engine_for, oracle_for
from oracle import oracle as O
from stateline_b200 import configs
from test_parity_gpu import _compare
from test_ref_pin import FIXTURES, REPLAY_HOT_BOUND, load_fixture, replay_cold_bound

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw,rounds", [("config1", dict(), 61), ("config2", dict(S=5), 45), ("config3", dict(S=3), 45),
                                            ("config5", dict(S=2), 23)])
def test_sequential_schedule_bit_exact(name, kw, rounds):
    w = configs.workload(name, **kw)
    orc = oracle_for(w, schedule=O.SCHED_SEQUENTIAL, seed=31)
    orc.init_chains()
    with engine_for(w, seed=31, schedule="sequential") as eng:
        eng.init_chains()
        _compare(eng, orc, w, "init")
        done = 0
        for chunk in (1, 8, 1, 11, rounds - 21):  # across swap rounds (every 10th) one round at a time and in bulk
            orc.run_rounds(chunk)
            eng.step_rounds(chunk)
            done += chunk
            _compare(eng, orc, w, "sequential, after %d rounds" % done)
        rows = eng.drain().reshape(done + 1, w["S"], w["d"] + 5)
        for s in range(w["S"]):
            assert np.array_equal(rows[:, s], orc.cold_rows(s))


def _fixture_workload(cfg):
    d = len(cfg["mn"])
    plugin = {O.TARGET_DEMO_GAUSS: "demo_gauss", O.TARGET_DOUBLE_WELL: "double_well", O.TARGET_GAUSS_FACT: "gauss_fact"}[cfg["target"]]
    return dict(S=cfg["S"], T=cfg["T"], J=cfg["J"], d=d, mn=cfg["mn"], mx=cfg["mx"], plugin=plugin,
                payload=cfg["target_params"], oracle_target=cfg["target"], oracle_params=cfg["target_params"],
                swapInterval=cfg["swap_interval"], optAccept=cfg["opt_accept"], optSwap=cfg["opt_swap"], name="fixture")


GPU_FIXTURES = [p for p in FIXTURES if json.load(open(p))["config"]["wire"] == 0]  # the device path has no "%f" wire text


@pytest.mark.parametrize("path", GPU_FIXTURES, ids=[os.path.basename(p)[:-5] for p in GPU_FIXTURES])
def test_reference_fixture_replayed_on_gpu(path):
    fx, cfg, C, d, n, x_init, z, u_mh, u_swap = load_fixture(path)
    w = _fixture_workload(cfg)
    orc = oracle_for(w, schedule=O.SCHED_SEQUENTIAL, rng_mode=O.RNG_REPLAY)
    orc.set_replay(z, u_mh, u_swap, x_init)
    orc.init_chains()
    orc.run_rounds(n)
    with engine_for(w, rng="replay", schedule="sequential") as eng:
        eng.set_replay(z, u_mh, u_swap, x_init)
        eng.init_chains()
        eng.step_rounds(n)
        _compare(eng, orc, w, "reference draws")  # GPU == oracle (same exp/log), bit for bit
        g = eng.last_rows()
        gs, gb = eng.sigma_beta()
        gc = eng.counters()
    ref = np.array([float.fromhex(h) for h in fx["last_rows"]]).reshape(C, d + 5)
    # against what the reference's own code computed from the same draws
    assert np.array_equal(g[:, d + 3:], ref[:, d + 3:])  # accepted, swapType of every chain's last row
    lg = {k: np.array([float.fromhex(h) for h in v]) for k, v in fx["logger"].items()}
    for k in ("length", "accepts", "swaps", "swap_attempts"):  # every accept / swap decision of the whole run
        assert np.array_equal(gc[k].astype(np.float64), lg[k]), k
    scale = np.maximum(np.abs(ref[:, :d + 1]), 1.0)
    err = np.max(np.abs(g[:, :d + 1] - ref[:, :d + 1]) / scale, axis=1).reshape(cfg["S"], cfg["T"])
    # per tier (tests/test_ref_pin.py: replay_cold_bound): the cold chains meet the north star's 1e-12 (1e-10 where the hot
    # states reach the cold chain within the run), the chains at the sigma cap stay below 1e-9
    assert err.max() < REPLAY_HOT_BOUND and err[:, 0].max() < replay_cold_bound(path)
    rs = np.array([float.fromhex(h) for h in fx["sigma"]])
    rb = np.array([float.fromhex(h) for h in fx["beta"]])
    assert np.max(np.abs(gs - rs) / rs) < 1e-12 and np.max(np.abs(gb - rb) / rb) < 1e-12
