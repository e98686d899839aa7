"""GPU tests of the engine around the kernels: sinks, stop rule, error states, statistical equivalence
with the reference's own (sequential) schedule, and size-independent properties at the bench size."""
import os

import numpy as np
import pytest

import stateline_b200 as sb
from oracle import oracle as O
from stateline_b200 import Engine, configs, engine
from helpers import engine_for, oracle_for

pytestmark = pytest.mark.gpu


def test_csv_sink_matches_reference_layout(tmp_path):
    """sink "csv": <outputPath>/<stack>.csv, one row per cold-chain state, initial state first, formatted
    as CSVChainArrayWriter does (src/db/db.cpp:18-47); slgpu_run applies the stop rule of
    serverwrapper.cpp:100-107 (cold samples over all stacks >= nSamplesTotal)."""
    w = configs.workload("config1")
    cfg = configs.engine_config(w, sink="csv", seed=4)
    cfg["outputPath"] = str(tmp_path / "out")
    cfg["nSamplesTotal"] = 123
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.run()
        rounds = eng.rounds_done
    assert rounds == 62  # ceil(123 / 2 stacks)
    orc = oracle_for(w, seed=4)
    orc.init_chains()
    orc.run_rounds(rounds)
    for s in range(w["S"]):
        lines = open(os.path.join(cfg["outputPath"], "%d.csv" % s)).read().splitlines()
        want = [O.csv_row(r, w["d"]) for r in orc.cold_rows(s)]
        assert lines == want


def test_host_ring_backpressure_and_peek():
    w = configs.workload("config2", S=8)
    with engine_for(w, sink="host", hostRingRounds=25) as eng:
        eng.init_chains()
        eng.step_rounds(20)
        with pytest.raises(sb.SlgpuError) as ei:
            eng.step_rounds(20)  # 21 + 20 rounds of rows do not fit a 25-round ring
        assert ei.value.status == engine.E_FULL
        eng.sync()
        spans = eng.peek()
        assert sum(s.shape[0] for s in spans) == 21 * w["S"]
        first = np.concatenate(spans)
        eng.advance(first.shape[0])
        eng.step_rounds(20)
        rest = eng.drain()
        assert rest.shape[0] == 20 * w["S"]
    with engine_for(w, sink="host", hostRingRounds=25) as ref:
        ref.init_chains()
        ref.step_rounds(20)
        a = ref.drain()
        ref.step_rounds(20)
        b = ref.drain()
    assert np.array_equal(first, a) and np.array_equal(rest, b)


def test_state_errors():
    w = configs.workload("config1")
    with engine_for(w, sink="device") as eng:
        with pytest.raises(sb.SlgpuError) as ei:
            eng.step_rounds(1)  # not initialised
        assert ei.value.status == engine.E_STATE
        eng.init_chains()
        with pytest.raises(sb.SlgpuError):
            eng.init_chains()
        with pytest.raises(sb.SlgpuError):
            eng.drain()  # sink is not "host"
    with engine_for(w, rng="replay") as eng:
        with pytest.raises(sb.SlgpuError):
            eng.init_chains()  # replay needs recorded draws first


def test_deterministic_and_launch_split_invariant():
    """Same seed => same bits, however the rounds are cut into launches."""
    w = configs.workload("config3", S=16)
    outs = []
    for chunks, mrl in (((60,), 64), ((1, 2, 3, 54), 64), ((60,), 3)):
        with engine_for(w, seed=11, maxRoundsPerLaunch=mrl, hostRingRounds=80) as eng:
            eng.init_chains()
            for c in chunks:
                eng.step_rounds(c)
            outs.append((eng.last_rows(), eng.drain()))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


def test_statistical_equivalence_with_reference_schedule():
    """North-star check 2: with native Philox the posterior moments and the per-temperature accept / swap
    rates match the reference's own schedule (oracle, sequential: one shared tier model updated chain by
    chain in result-arrival order) within Monte-Carlo error.  z-bound 5 on means (standard errors from the
    between-stack spread), 10% relative on rates and standard deviations."""
    w = configs.workload("config1", S=64)
    rounds, burn = 6000, 1000
    orc = oracle_for(w, schedule=O.SCHED_SEQUENTIAL, seed=1)
    orc.init_chains()
    orc.run_rounds(rounds)
    o_cold = np.stack([orc.cold_rows(s)[burn:, :w["d"]] for s in range(w["S"])])
    with engine_for(w, seed=2, hostRingRounds=rounds + 8) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        g_cold = eng.drain().reshape(rounds + 1, w["S"], w["d"] + 5)[burn:, :, :w["d"]].transpose(1, 0, 2)
        gc, gsb = eng.counters(), eng.sigma_beta()
    oc, osb = orc.counters(), orc.sigma_beta()
    g_m, o_m = g_cold.mean(axis=1), o_cold.mean(axis=1)  # per-stack means
    se = np.sqrt(g_m.var(axis=0, ddof=1) / w["S"] + o_m.var(axis=0, ddof=1) / w["S"])
    z = np.abs(g_m.mean(axis=0) - o_m.mean(axis=0)) / se
    assert z.max() < 5.0, z
    assert np.allclose(g_cold.std(axis=(0, 1)), o_cold.std(axis=(0, 1)), rtol=0.1)
    T = w["T"]
    g_acc = (gc["accepts"] / gc["length"]).reshape(-1, T).mean(0)
    o_acc = (oc["accepts"] / oc["length"]).reshape(-1, T).mean(0)
    assert np.allclose(g_acc, o_acc, rtol=0.1), (g_acc, o_acc)
    g_sw = (gc["swaps"] / gc["swap_attempts"]).reshape(-1, T).mean(0)[:T - 1]
    o_sw = (oc["swaps"] / oc["swap_attempts"]).reshape(-1, T).mean(0)[:T - 1]
    assert np.allclose(g_sw, o_sw, rtol=0.1), (g_sw, o_sw)
    assert np.allclose(gsb[1].reshape(-1, T).mean(0), osb[1].reshape(-1, T).mean(0), rtol=0.25)  # beta ladder


def test_double_well_mode_hopping():
    """Config #2: swaps let the cold chain visit both wells; P(x0 > 0) = 0.5 within 5 standard errors."""
    w = configs.workload("config2")
    rounds, burn = 4000, 500
    with engine_for(w, seed=6, hostRingRounds=rounds + 8) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        rows = eng.drain().reshape(rounds + 1, w["S"], w["d"] + 5)[burn:]
    x0 = rows[:, :, 0]
    frac = (x0 > 0).mean(axis=0)  # per stack
    hops = (np.diff(np.sign(x0), axis=0) != 0).sum(axis=0)
    assert hops.min() > 0
    se = frac.std(ddof=1) / np.sqrt(w["S"])
    assert abs(frac.mean() - 0.5) < 5.0 * se + 0.01


def test_bench_size_properties():
    """BASELINE config #3 at full size (8192 stacks x 8 temperatures): properties that do not need the oracle."""
    w = configs.workload("config3")
    S, T, d = w["S"], w["T"], w["d"]
    with engine_for(w, seed=1234, sink="host", hostRingRounds=64) as eng:
        eng.init_chains()
        eng.step_rounds(39)
        rows = eng.drain()
        assert rows.shape == (40 * S, d + 5)
        rows = rows.reshape(40, S, d + 5)
        x, e = rows[:, :, :d], rows[:, :, d]
        assert np.all(x >= -10.0) and np.all(x <= 10.0)                       # bouncyBounds keeps samples in the box
        assert np.all(rows[:, :, d + 2] == 1.0)                               # the cold chain's beta is never rewritten
        assert set(np.unique(rows[:, :, d + 3])) <= {0.0, 1.0}
        st = rows[:, :, d + 4]
        assert set(np.unique(st)) <= {0.0, 1.0, 2.0}
        swap_rows = [r for r in range(1, 40) if (r + 1) % 10 == 0]           # row r = round r-1; chain length r+1
        assert np.all(st[swap_rows] != 0.0) and np.all(np.delete(st, swap_rows, axis=0) == 0.0)
        # energy column == sum over job types of the functor at the stored sample (checked on a slice)
        pts = x[-1, :256]
        assert np.array_equal(eng.eval_likelihood(pts).cumsum(axis=1)[:, -1], e[-1, :256]) or \
            np.allclose(eng.eval_likelihood(pts).sum(axis=1), e[-1, :256], rtol=1e-14)
        sig, beta = eng.sigma_beta()
        beta = beta.reshape(S, T)
        assert np.all(beta[:, 0] == 1.0) and np.all(np.diff(beta, axis=1) <= 0.0)  # non-increasing ladder (adaptive.cpp:118-132)
        assert np.all(sig > 0) and np.all(sig <= np.exp(4.0) * (1 + 1e-15))
        c = eng.counters()
        assert np.all(c["length"] == 40) and np.all(c["accepts"] <= 40)
        assert np.all(c["swap_attempts"].reshape(S, T)[:, :T - 1] == 1 + 4)  # swap rounds 8, 18, 28, 38
        assert np.all(c["swap_attempts"].reshape(S, T)[:, T - 1] == 1)
        r = eng.rhat()
        assert np.all(np.isfinite(r))
        summ = eng.summary()
        assert summ["rounds"] == 39 and len(summ["accept_rate"]) == T


def test_user_plugin_and_cpp_host(tmp_path):
    """examples/: a user plugin .so with a POD payload and a +inf region (chainarray.cpp:36-37: certain reject),
    driven once through the Python binding and once through the C++ host example linked against libslgpu.so."""
    import subprocess
    from stateline_b200 import build
    so = build.build_plugin(os.path.join(build.ROOT, "examples", "ring_plugin.cu"), str(tmp_path / "libring_plugin.so"))
    payload = np.array([1.5, 0.2, 2.0])  # radius, width, cutoff
    cfg = sb.make_config(8, 4, [-3.0, -3.0, -3.0], [3.0, 3.0, 3.0], 2, initial=[1.5, 0.0, 0.0], sink="host", hostRingRounds=600, seed=3)
    with Engine(cfg, plugin=so, payload=payload) as eng:
        eng.init_chains()
        eng.step_rounds(500)
        rows = eng.drain()
    r = np.linalg.norm(rows[:, :3], axis=1)
    assert r.max() <= 2.0 and np.all(np.isfinite(rows[:, 3]))           # never left the ball: +inf proposals always rejected
    assert abs(np.median(r[8 * 200:]) - 1.5) < 0.2                        # and sits on the ring
    # stateline-gpu, the C++ host over the C-ABI: the reference's command line, csv files, table logger, checkpoints
    exe = build.build_cli()
    w = configs.workload("config1")
    c2 = configs.engine_config(w, seed=4)
    assert "sink" not in c2["gpu"]  # the CLI defaults to csv like the reference
    c2["gpu"].update(plugin=build.TARGETS_SO, entry="slgpu_plugin_demo_gauss")
    c2["outputPath"] = str(tmp_path / "demo-output")
    c2["nSamplesTotal"] = 200
    c2["loggingRateSec"] = 0.0
    import json
    (tmp_path / "cfg.json").write_text(json.dumps(c2, indent=2))
    ck = str(tmp_path / "ck.slgpu")
    out = subprocess.check_output([exe, "-c", str(tmp_path / "cfg.json"), "-p", "5555", "--checkpoint", ck]).decode()
    summ = json.loads(out.strip().splitlines()[-1])
    assert summ["rounds"] == 100
    assert len(open(os.path.join(c2["outputPath"], "1.csv")).read().splitlines()) == 101
    orc = oracle_for(w, seed=4)
    orc.init_chains()
    orc.run_rounds(100)
    assert out.count("  ID    Length    MinEngy") >= 1 and orc.format_table() in out  # the final table is the oracle's
    # resume for another 60 rounds, quietly: the files continue, no table is printed
    c2["nSamplesTotal"] = 320
    (tmp_path / "cfg2.json").write_text(json.dumps(c2))
    out = subprocess.check_output([exe, "-c", str(tmp_path / "cfg2.json"), "-l", "-1", "--resume", ck]).decode()
    assert json.loads(out.strip().splitlines()[-1])["rounds"] == 160 and "MinEngy" not in out
    orc.run_rounds(60)
    for s in range(w["S"]):
        lines = open(os.path.join(c2["outputPath"], "%d.csv" % s)).read().splitlines()
        assert lines == [O.csv_row(r, w["d"]) for r in orc.cold_rows(s)]
    # errors: the reference's message for a missing config; a config without a plugin
    r = subprocess.run([exe, "-c", str(tmp_path / "nope.json")], capture_output=True)
    assert r.returncode == 1 and b"Could not find config file" in r.stderr
    del c2["gpu"]["plugin"]
    (tmp_path / "cfg3.json").write_text(json.dumps(c2))
    assert subprocess.run([exe, "-c", str(tmp_path / "cfg3.json")], capture_output=True).returncode == 2


def _full_state(eng, chains):
    sig, beta = eng.sigma_beta()
    out = [eng.last_rows(), sig, beta, eng.ladder(), eng.rhat()]
    for which in (0, 1):
        out += list(eng.models(which))
    c = eng.counters()
    out += [c[k] for k in sorted(c)]
    for ch in chains:
        out += list(eng.shaper(ch))
    return [np.asarray(v) for v in out]


@pytest.mark.parametrize("name,S,cut", [("config3", 16, 17), ("config3", 16, 19), ("config4", 4, 8), ("config2", 8, 1)])
def test_checkpoint_resume_is_bit_exact(tmp_path, name, S, cut):
    """slgpu_save_state / slgpu_load_state (the recoverFromDisk the reference declares at chainarray.hpp:158-162
    and never defines): a run cut at any round -- inside a swap interval, or right before the adapt round (19,
    where the tier statistics of 9 rounds are pending) -- and resumed in a fresh engine gives the same bits as
    the uninterrupted run: rows, chain state, shaper factors, tier models, ladder, counters and R-hat."""
    w = configs.workload(name, S=S)
    total = 41
    chains = [0, 1, w["T"], S * w["T"] - 1]
    with engine_for(w, seed=21) as eng:
        eng.init_chains()
        eng.step_rounds(total)
        want_rows, want_state = eng.drain(), _full_state(eng, chains)
    ck = tmp_path / "state.slgpu"
    with engine_for(w, seed=21) as eng:
        eng.init_chains()
        eng.step_rounds(cut)
        eng.save_state(ck)
        head = eng.drain()
        eng.step_rounds(3)  # the checkpoint is a snapshot: the donor carries on unharmed
    assert not os.path.exists(str(ck) + ".tmp")
    with engine_for(w, seed=21) as eng:
        eng.load_state(ck)
        assert eng.rounds_done == cut
        eng.step_rounds(total - cut)
        tail = eng.drain()
        got_state = _full_state(eng, chains)
    assert np.array_equal(np.concatenate([head, tail]), want_rows)
    for a, b in zip(got_state, want_state):
        assert np.array_equal(a, b)


def test_checkpoint_rejects_foreign_and_damaged_files(tmp_path):
    w = configs.workload("config1")
    ck = tmp_path / "a.slgpu"
    with engine_for(w, seed=3) as eng:
        with pytest.raises(sb.SlgpuError) as ei:
            eng.save_state(ck)  # nothing to save before the chains exist
        assert ei.value.status == engine.E_STATE
        eng.init_chains()
        eng.step_rounds(12)
        eng.save_state(ck)
        before = eng.last_rows()
        with pytest.raises(sb.SlgpuError) as ei:
            eng.load_state(tmp_path / "missing.slgpu")
        assert ei.value.status == engine.E_IO
        blob = open(ck, "rb").read()
        open(tmp_path / "cut.slgpu", "wb").write(blob[: len(blob) // 2])
        with pytest.raises(sb.SlgpuError) as ei:
            eng.load_state(tmp_path / "cut.slgpu")
        assert ei.value.status == engine.E_IO
        open(tmp_path / "junk.slgpu", "wb").write(b"x" * 4096)
        with pytest.raises(sb.SlgpuError) as ei:
            eng.load_state(tmp_path / "junk.slgpu")
        assert ei.value.status == engine.E_IO
        assert np.array_equal(eng.last_rows(), before) and eng.rounds_done == 12  # failed loads change nothing
    for other in (dict(seed=4), dict(seed=3, stack_offset=2)):
        with engine_for(w, **other) as eng:
            with pytest.raises(sb.SlgpuError) as ei:
                eng.load_state(ck)
            assert ei.value.status == engine.E_CONFIG
    w2 = configs.workload("config2", S=2)
    with engine_for(w2, seed=3) as eng:
        with pytest.raises(sb.SlgpuError) as ei:
            eng.load_state(ck)
        assert ei.value.status == engine.E_CONFIG


def test_csv_resume_appends(tmp_path):
    """A run checkpointed and resumed with "gpu": {"append": true} leaves the same <stack>.csv files as one run."""
    w = configs.workload("config1")

    def cfg_for(out, **gpu):
        cfg = configs.engine_config(w, sink="csv", seed=4, **gpu)
        cfg["outputPath"] = str(tmp_path / out)
        return cfg

    cfg = cfg_for("whole")
    cfg["nSamplesTotal"] = 2 * 50
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.run()
    ck = tmp_path / "mid.slgpu"
    cfg = cfg_for("parts")
    cfg["nSamplesTotal"] = 2 * 23
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.run()
        eng.save_state(ck)
    cfg = cfg_for("parts", append=True)
    cfg["nSamplesTotal"] = 2 * 50
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.load_state(ck)
        eng.run()
    for s in range(w["S"]):
        a = open(tmp_path / "whole" / ("%d.csv" % s)).read()
        b = open(tmp_path / "parts" / ("%d.csv" % s)).read()
        assert a == b and a.count("\n") == 51


@pytest.mark.parametrize("name,S,rounds,swap_interval", [("config1", 2, 30, None), ("config2", 8, 2300, 2), ("config4", 2, 45, 5)])
def test_table_logger_matches_oracle(name, S, rounds, swap_interval):
    """TableLogger parity (src/infer/logging.cpp:35-87): energies_, the adapters' windowed rates (adaptive.cpp:80-90;
    2300 rounds with swapInterval 2 wrap both 1000-entry windows) and the printed table, character for character."""
    w = configs.workload(name, S=S)
    if swap_interval:
        w["swapInterval"] = swap_interval
    with engine_for(w, seed=9, sink="none", windowRates=True) as eng:
        eng.init_chains()
        first = eng.format_table()
        eng.step_rounds(rounds)
        got_log, got_table = eng.logger(), eng.format_table()
    orc = oracle_for(w, seed=9)
    orc.init_chains()
    assert first == orc.format_table()
    orc.run_rounds(rounds)
    for a, b in zip(got_log, orc.logger()):
        assert np.array_equal(a, b)
    assert got_table == orc.format_table()
    with engine_for(w, seed=9, sink="none") as eng:
        eng.init_chains()
        eng.step_rounds(3)
        assert np.array_equal(eng.logger(rates=False)[0], oracle_cur3(w))
        with pytest.raises(sb.SlgpuError) as ei:
            eng.format_table()  # the windowed-rate columns were not asked for
        assert ei.value.status == engine.E_STATE


def oracle_cur3(w):
    orc = oracle_for(w, seed=9)
    orc.init_chains()
    orc.run_rounds(3)
    return orc.logger()[0]


def test_checkpoint_keeps_the_rate_windows(tmp_path):
    w = configs.workload("config2", S=4)
    w["swapInterval"] = 3
    with engine_for(w, seed=2, sink="none", windowRates=True) as eng:
        eng.init_chains()
        eng.step_rounds(1300)
        want = eng.format_table()
    ck = tmp_path / "w.slgpu"
    with engine_for(w, seed=2, sink="none", windowRates=True) as eng:
        eng.init_chains()
        eng.step_rounds(1100)
        eng.save_state(ck)
    with engine_for(w, seed=2, sink="none", windowRates=True) as eng:
        eng.load_state(ck)
        eng.step_rounds(200)
        assert eng.format_table() == want
    # the windows are logging only: a checkpoint with them loads into an engine that keeps none, and the other way
    # round the windows start empty; the chains continue identically either way
    with engine_for(w, seed=2, sink="none") as eng:
        eng.load_state(ck)
        eng.step_rounds(200)
        plain = eng.last_rows()
        ck2 = tmp_path / "plain.slgpu"
        eng.save_state(ck2)
    with engine_for(w, seed=2, sink="none", windowRates=True) as eng:
        eng.load_state(ck2)
        assert np.array_equal(eng.last_rows(), plain)
        cur, ar, sr = eng.logger()
        assert not ar.any() and not sr.any()
        eng.step_rounds(10)
        assert eng.logger()[1].max() <= 1.0 and eng.rounds_done == 1310


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_csv_formatter_pool(tmp_path, threads):
    """The csv sink's formatter threads own disjoint stacks; whatever their number, every <stack>.csv holds that
    stack's rows in order, as the oracle formats them (a 7-round ring forces many pump / back-pressure cycles)."""
    w = configs.workload("config2", S=37)
    cfg = configs.engine_config(w, sink="csv", seed=6, csvThreads=threads, hostRingRounds=7)
    cfg["outputPath"] = str(tmp_path / "out")
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.init_chains()
        for n in (1, 5, 33, 20):
            eng.step_rounds(n)
        eng.sync()  # flushes
        orc = oracle_for(w, seed=6)
        orc.init_chains()
        orc.run_rounds(59)
        for s in (0, 1, 17, 36):
            lines = open(os.path.join(cfg["outputPath"], "%d.csv" % s)).read().splitlines()
            assert lines == [O.csv_row(r, w["d"]) for r in orc.cold_rows(s)]
        eng.step_rounds(4)
    # destroy flushed the last four rounds
    assert open(os.path.join(cfg["outputPath"], "36.csv")).read().count("\n") == 64
