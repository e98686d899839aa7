"""slgpu_group_*: several devices driven from one process (what one `stateline -c cfg` server is on the reference side,
src/bin/stateline.cpp:52-90).  A group must equal, bit for bit, the engines one would create by hand for its shards; its
R-hat is the reference's EPSRDiagnostic over all stacks.  On a box with one GPU the two shards share the device (the
sharding, the global stack offsets and the in-process reduction are exercised all the same); with two or more GPUs the
shards sit on devices 0 and 1, which also covers two engines on different devices in one process."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle as O
from stateline_b200 import EngineGroup, build, configs
from helpers import engine_for

pytestmark = pytest.mark.gpu


def _devices():
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


@pytest.mark.parametrize("name,S,rounds", [("config3", 37, 31), ("config4", 5, 12)])
def test_group_equals_the_union_of_its_shards(name, S, rounds):
    w = configs.workload(name, S=S)
    devs = _devices()
    cfg = configs.engine_config(w, sink="host", seed=11, stackOffset=5, hostRingRounds=64)
    with EngineGroup(cfg, plugin=w["plugin"], payload=w["payload"], devices=devs) as g:
        assert [e.S for e in g.engines] == [S - S // 2, S // 2]
        g.init_chains()
        g.step_rounds(rounds)
        g.sync()
        assert g.rounds_done == rounds
        got = [(e.drain(), e.last_rows(), e.sigma_beta(), e.models(0), e.models(1), e.rhat_stage1()) for e in g.engines]
        rhat_group = g.rhat()
        summ = g.summary()
    assert summ["devices"] == 2 and summ["engines"][1]["stackOffset"] == 5 + S - S // 2
    off = 5
    for i, e_got in enumerate(got):
        n = S - S // 2 if i == 0 else S // 2
        wi = configs.workload(name, S=n)
        with engine_for(wi, seed=11, stack_offset=off, device=devs[i], hostRingRounds=64) as eng:
            eng.init_chains()
            eng.step_rounds(rounds)
            want = (eng.drain(), eng.last_rows(), eng.sigma_beta(), eng.models(0), eng.models(1), eng.rhat_stage1())
        for a, b in zip(e_got, want):
            for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
                assert np.array_equal(np.asarray(x), np.asarray(y))
        off += n
    assert np.all(np.isfinite(rhat_group)) and rhat_group.shape == (w["d"],)


def test_group_rhat_is_the_reference_epsr_over_all_stacks():
    """One shard per device, R-hat over the union: equals EPSRDiagnostic::rHat (diagnostics.cpp:49-71) evaluated by the
    oracle's restatement on the cold-chain rows of all stacks."""
    S, rounds = 12, 80
    w = configs.workload("config2", S=S)
    cfg = configs.engine_config(w, sink="host", seed=3, hostRingRounds=rounds + 4)
    with EngineGroup(cfg, plugin=w["plugin"], devices=_devices()) as g:
        g.init_chains()
        g.step_rounds(rounds)
        rhat = g.rhat()
        rows = [e.drain().reshape(rounds + 1, e.S, w["d"] + 5) for e in g.engines]
    x = np.concatenate(rows, axis=1)[1:, :, :w["d"]]  # EPSR sees the rows step() returned: not the initial state
    n, m = x.shape[0], x.shape[1]
    M, V = x.mean(0), x.var(0, ddof=0) * n  # Welford M and S per stack and dim
    B = n / (m - 1.0) * ((M - M.mean(0)) ** 2).sum(0)
    Wv = (V / (n - 1.0)).mean(0)
    want = np.sqrt(((n - 1.0) / n * Wv + B / n) / (Wv + 1e-30))
    assert np.allclose(rhat, want, rtol=1e-9, atol=0)


def test_group_run_applies_the_stop_rule_to_the_total(tmp_path):
    w = configs.workload("config1", S=5)
    cfg = configs.engine_config(w, sink="csv", seed=4)
    cfg["outputPath"] = str(tmp_path / "out")
    cfg["nSamplesTotal"] = 123
    with EngineGroup(cfg, plugin=w["plugin"], devices=_devices()) as g:
        g.run()
        assert g.rounds_done == 25  # ceil(123 / 5 stacks), on every shard
    names = sorted(os.listdir(cfg["outputPath"]), key=lambda f: int(f.split(".")[0]))
    assert names == ["%d.csv" % s for s in range(5)]  # global stack indices
    assert all(len(open(os.path.join(cfg["outputPath"], f)).read().splitlines()) == 26 for f in names)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="stateline-gpu --gpus 2 needs two devices")
def test_cli_on_two_gpus(tmp_path):
    exe = build.build_cli()
    w = configs.workload("config1", S=4)
    c = configs.engine_config(w, seed=4)
    c["gpu"].update(plugin=build.TARGETS_SO, entry="slgpu_plugin_demo_gauss")
    c["outputPath"] = str(tmp_path / "out")
    c["nSamplesTotal"] = 400
    (tmp_path / "cfg.json").write_text(json.dumps(c))
    out = subprocess.check_output([exe, "-c", str(tmp_path / "cfg.json"), "-l", "-1", "--gpus", "2"]).decode()
    summ = json.loads(out.strip().splitlines()[-1])
    assert summ["devices"] == 2 and len(summ["rhat"]) == w["d"]
    for i, off in enumerate((0, 2)):
        wi = configs.workload("config1", S=2)
        orc = O.Sampler(S=2, T=wi["T"], J=wi["J"], swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=wi["mn"], mx=wi["mx"],
                        target=wi["oracle_target"], target_params=None, schedule=O.SCHED_BATCHED, rng_mode=O.RNG_PHILOX, seed=4,
                        stack_offset=off)
        orc.init_chains()
        orc.run_rounds(100)
        for s in range(2):
            lines = open(os.path.join(c["outputPath"], "%d.csv" % (off + s))).read().splitlines()
            assert lines == [O.csv_row(r, wi["d"]) for r in orc.cold_rows(s)]
