"""The delta stream (gpu.wire "delta"): what crosses PCIe is one flag byte per cold row plus the d + 3 doubles of the rows
whose cold state changed; the host rebuilds the exact rows.  Everything a consumer sees -- slgpu_drain, slgpu_peek, the csv
files, the binary chain file -- must equal the full-row wire and the oracle, bit for bit."""
import os

import numpy as np
import pytest

from stateline_b200 import Engine, binlog, configs
from helpers import engine_for, oracle_for

pytestmark = pytest.mark.gpu


def _oracle_rows(w, seed, rounds):
    orc = oracle_for(w, seed=seed)
    orc.init_chains()
    orc.run_rounds(rounds)
    return np.stack([orc.cold_rows(s) for s in range(w["S"])], axis=1)  # [rounds + 1, S, d + 5]


@pytest.mark.parametrize("name,S,rounds", [("config1", 2, 57), ("config2", 300, 43), ("config3", 520, 33), ("config4", 5, 13)])
def test_drain_is_identical_on_both_wires(name, S, rounds):
    w = configs.workload(name, S=S)
    want = _oracle_rows(w, 21, rounds)
    got = {}
    for wire in ("full", "delta"):
        with engine_for(w, seed=21, wire=wire) as eng:
            eng.init_chains()
            for chunk in (1, 2, rounds - 3):  # launches of 1, 2 and many rounds: segments of every size
                eng.step_rounds(chunk)
            got[wire] = (eng.drain().reshape(rounds + 1, S, w["d"] + 5), eng.d2h_bytes)
    assert np.array_equal(got["full"][0], want)
    assert np.array_equal(got["delta"][0], want)
    assert got["delta"][1] < got["full"][1]


def test_delta_stream_bytes_and_segments():
    """Config #3 slice: the segment view rebuilds the rows, and once the chains have adapted the copy engine moves at
    least 2.5x fewer bytes than full rows would (the first segments are copied whole: nothing has landed yet to size them)."""
    S, warm, rounds = 1024, 60, 300
    w = configs.workload("config3", S=S)
    want = _oracle_rows(w, 5, warm + rounds + 7)
    with engine_for(w, seed=5, wire="delta") as eng:
        eng.init_chains()
        eng.step_rounds(warm)
        eng.sync()
        prev = np.zeros((S, w["d"] + 3))
        rows = []
        first = True
        while True:
            sg = eng.next_segment(block=True)
            if sg is None:
                break
            assert sg["key"] == first and sg["row_begin"] == sum(r.shape[0] for r in rows) * S
            first = False
            rows.append(binlog.expand_segment(prev, sg["n_rounds"], sg["base"], sg["flags"], sg["payload"]))
            eng.release_segment()
        assert np.array_equal(np.concatenate(rows), want[:warm + 1])
        b0 = eng.d2h_bytes
        rows = []
        for _ in range(rounds // 10):  # the way a consumer runs: step, take what has landed
            eng.step_rounds(10)
            while True:
                sg = eng.next_segment(block=False)
                if sg is None:
                    break
                rows.append(binlog.expand_segment(prev, sg["n_rounds"], sg["base"], sg["flags"], sg["payload"]))
                eng.release_segment()
        eng.sync()
        delta_bytes = eng.d2h_bytes - b0
        while True:
            sg = eng.next_segment(block=True)
            if sg is None:
                break
            rows.append(binlog.expand_segment(prev, sg["n_rounds"], sg["base"], sg["flags"], sg["payload"]))
            eng.release_segment()
        assert np.array_equal(np.concatenate(rows), want[warm + 1:warm + 1 + rounds])
        eng.step_rounds(7)  # segments and drain mix at segment boundaries
        tail = eng.drain().reshape(7, S, w["d"] + 5)
    assert np.array_equal(tail, want[-7:])
    full_bytes = rounds * S * (w["d"] + 5) * 8
    # 2.5x, not the bench's 3.2x: these are rounds 60..360 (the accept rates are still settling), and a copy is sized from the
    # last segment that had LANDED when it was enqueued (+25 %), which depends on timing
    assert delta_bytes * 2.5 <= full_bytes, (delta_bytes, full_bytes)


def test_csv_and_bin_sinks_on_the_delta_wire(tmp_path):
    S, rounds = 37, 41
    w = configs.workload("config2", S=S)
    want = _oracle_rows(w, 8, rounds)
    texts = {}
    for wire in ("full", "delta"):
        cfg = configs.engine_config(w, sink="csv", seed=8, wire=wire, csvThreads=3)
        cfg["outputPath"] = str(tmp_path / wire)
        with Engine(cfg, plugin=w["plugin"]) as eng:
            eng.init_chains()
            eng.step_rounds(rounds)
            eng.sync()
        texts[wire] = [open(os.path.join(cfg["outputPath"], "%d.csv" % s)).read() for s in range(S)]
    assert texts["full"] == texts["delta"]
    cfg = configs.engine_config(w, sink="bin", seed=8, stackOffset=3)
    cfg["outputPath"] = str(tmp_path / "bin")
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        eng.sync()
    hdr, rows = binlog.read(os.path.join(cfg["outputPath"], "chains-3.slgpu"))
    assert (hdr["n_stacks"], hdr["n_dims"], hdr["stack_offset"]) == (S, w["d"], 3)
    orc = oracle_for(w, seed=8, stack_offset=3)
    orc.init_chains()
    orc.run_rounds(rounds)
    assert np.array_equal(rows, np.stack([orc.cold_rows(s) for s in range(S)], axis=1))
    assert want.shape == rows.shape


def test_overwrite_and_resume_start_with_key_segments(tmp_path):
    """gpu.overwrite: a lagging consumer loses the oldest segments, and what it reads afterwards is still exact (every
    segment is a key segment); a resumed engine's first segment is a key segment as well."""
    S, d = 64, 20
    w = configs.workload("config3", S=S)
    want = _oracle_rows(w, 2, 120)
    with engine_for(w, seed=2, overwrite=True, hostRingRounds=30) as eng:
        eng.init_chains()
        eng.step_rounds(120)
        eng.sync()
        rows = eng.drain()
        n = rows.shape[0] // S
        assert 10 <= n <= 31
        assert np.array_equal(rows.reshape(n, S, d + 5), want[-n:])
    ck = str(tmp_path / "ck.bin")
    with engine_for(w, seed=2) as eng:
        eng.init_chains()
        eng.step_rounds(55)
        eng.save_state(ck)
    with engine_for(w, seed=2) as eng:
        eng.load_state(ck)
        eng.step_rounds(65)
        sg = eng.next_segment(block=True)
        assert sg["key"]
        rows = eng.drain().reshape(65, S, d + 5)
    assert np.array_equal(rows, want[-65:])


def test_discarded_segments_and_the_next_key_segment():
    """slgpu_discard_segment: the cheap release for consumers that only take segments.  The host's row state is then
    stale: drain() works again from a key segment (slgpu_request_key_segment), a non-key segment before that is refused."""
    from stateline_b200 import SlgpuError
    S, d = 48, 20
    w = configs.workload("config3", S=S)
    want = _oracle_rows(w, 13, 60)
    with engine_for(w, seed=13) as eng:
        eng.init_chains()
        eng.step_rounds(30)
        eng.sync()
        n = 0
        while True:
            sg = eng.next_segment(block=True)
            if sg is None:
                break
            n += sg["n_rounds"]
            eng.release_segment(keep_rows_state=False)
        assert n == 31
        eng.request_key_segment()
        eng.step_rounds(30)
        sg = eng.next_segment(block=True)
        assert sg["key"]
        rows = eng.drain().reshape(30, S, d + 5)
        assert np.array_equal(rows, want[31:])
    with engine_for(w, seed=13) as eng:
        eng.init_chains()
        eng.step_rounds(30)  # several non-key segments are queued behind the initial one
        eng.sync()
        eng.next_segment(block=True)
        eng.release_segment(keep_rows_state=False)
        with pytest.raises(SlgpuError):
            eng.drain()
