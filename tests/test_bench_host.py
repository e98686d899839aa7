"""Host-side pieces of bench.py that need no GPU: the reference arm's JSON contract, the ESS estimator, the
algorithmic byte count of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                                   "--steps", "2", "--warmup", "1", "--ref-seconds", "3"], cwd=ROOT, timeout=300)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "chain_steps_per_sec" and line["unit"] == "chain-steps/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    # oracle/_ref (the reference's own object code) when it is built or buildable here, else the oracle port
    assert line["cpu_baseline"]["kind"] == ("reference" if bench.reference_available() else "port")
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] >= 1 and "config #3" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("config3") and line["dtype"] == "f64"


def test_reference_arm_port_fallback_keeps_the_contract():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-port",
                                   "--steps", "1", "--warmup", "1", "--ref-seconds", "2"], cwd=ROOT, timeout=300)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert "server +" in line["cpu_baseline"]["sample"]


def test_reference_pool_counts_chain_steps():
    if not bench.reference_available():
        import pytest
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    pool = bench.ReferencePool(2, 3)
    try:
        secs, steps = pool.step(5)
    finally:
        pool.close()
    assert steps == 2 * 3 * 8 * 5 and secs > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == b""


def test_ess_of_an_ar1_series():
    """AR(1) with coefficient rho has ESS/n = (1 - rho) / (1 + rho)."""
    rng = np.random.default_rng(3)
    n, rho = 200_000, 0.8
    e = rng.standard_normal(n)
    x = np.empty(n)
    x[0] = e[0]
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    assert abs(bench.ess_geyer(x) / n - (1 - rho) / (1 + rho)) < 0.015
    assert abs(bench.ess_geyer(rng.standard_normal(50_000)) / 50_000 - 1.0) < 0.05


def test_algorithmic_bytes_closed_form():
    # (d + d(d+1)/2 + 6) doubles in and out once per launch of K rounds + (d+5) doubles per cold row
    assert bench.algorithmic_bytes_per_chain_step(20, 8, 10) == (20 + 210 + 6) * 16 / 10 + 25 * 8 / 8
    assert round(bench.algorithmic_bytes_per_chain_step(20, 8, 10)) == 403


def test_stdout_carries_only_the_json_line():
    """Native libraries print on fd 1 (NCCL's version line under NCCL_DEBUG): after bench._claim_stdout() such
    writes go to stderr and only bench._emit() reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._claim_stdout(); "
            "os.write(1, b'NCCL version 0.0\\n'); print('python-level noise'); bench._emit('{\"ok\": 1}')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=120)
    assert r.returncode == 0
    assert r.stdout.decode().splitlines() == ['{"ok": 1}']
    assert b"NCCL version 0.0" in r.stderr and b"python-level noise" in r.stderr
