"""GPU parity: the CUDA path through the C-ABI against the CPU oracle on the same draws.

Bar (north_star): fed the same recorded proposal / uniform draws, accept and swap decisions match
exactly and fp64 chain states agree within 1e-12 relative.  Because both sides use the same
correctly-rounded operation sequence (include/slgpu_math.h for exp / log / sincos, -fmad=false /
-ffp-contract=off elsewhere), the tests below demand more: BIT-IDENTICAL states, adaptation state,
counters and cold-chain rows, in replay mode and with native Philox draws."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from stateline_b200 import configs, dist as sd
from helpers import engine_for, make_replay, oracle_for

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

SMALL = {
    "config1": dict(),                 # README demo, S=2 T=5 d=4 J=3 (fast kernel, idle lanes: T does not divide 32)
    "config2": dict(S=16),             # double well, T=8 d=2
    "config3": dict(S=12),             # 20-D factorised Gaussian, T=8 (fast kernel, the bench workload)
    "config4": dict(S=3),              # 100-D correlated Gaussian, T=8 (runtime-d kernel)
    "config5": dict(S=3),              # 50-D Rosenbrock, T=16 (runtime-d kernel)
}


def _compare(eng, orc, w, tag):
    d = w["d"]
    g_rows, o_rows = eng.last_rows(), orc.last_rows()
    assert np.array_equal(g_rows[:, d + 3:], o_rows[:, d + 3:]), tag      # decisions: accepted, swapType
    assert np.array_equal(g_rows, o_rows), tag                            # x, energy, sigma, beta of the last row
    gs, gb = eng.sigma_beta()
    os_, ob = orc.sigma_beta()
    assert np.array_equal(gs, os_) and np.array_equal(gb, ob), tag
    gc, oc = eng.counters(), orc.counters()
    for k in ("length", "accepts", "swaps", "swap_attempts", "min_energy"):
        assert np.array_equal(gc[k], oc[k]), (tag, k)
    for which in (0, 1):
        gm, om = eng.models(which), orc.models(which)
        for a, b in zip(gm, om):
            assert np.array_equal(a, b), (tag, which)
    for chain in (0, w["T"] - 1, w["S"] * w["T"] - 1):
        gL, gN, gcnt = eng.shaper(chain)
        oL, oN, ocnt = orc.shaper(chain)
        assert gcnt == ocnt
        assert np.array_equal(gL, oL) and np.array_equal(gN, oN), (tag, chain)


@pytest.mark.parametrize("name", list(SMALL))
def test_replay_exact(name):
    w = configs.workload(name, **SMALL[name])
    n_rounds = 45 if w["d"] <= 20 else 25
    z, u_mh, u_swap, x_init = make_replay(w["S"], w["T"], w["d"], n_rounds, seed=sum(map(ord, name)))
    orc = oracle_for(w, rng_mode=O.RNG_REPLAY)
    orc.set_replay(z, u_mh, u_swap, x_init)
    orc.init_chains()
    with engine_for(w, rng="replay") as eng:
        eng.set_replay(z, u_mh, u_swap, x_init)
        eng.init_chains()
        _compare(eng, orc, w, "init")
        done = 0
        for chunk in (1, 7, 2, n_rounds - 10):  # uneven chunks: launches split at adapt points
            orc.run_rounds(chunk)
            eng.step_rounds(chunk)
            done += chunk
            _compare(eng, orc, w, "after %d rounds" % done)
        # cold-chain rows, emission order: initial row then one per round, per stack
        rows = eng.drain()
        assert rows.shape == ((done + 1) * w["S"], w["d"] + 5)
        rows = rows.reshape(done + 1, w["S"], w["d"] + 5)
        for s in range(w["S"]):
            assert np.array_equal(rows[:, s], orc.cold_rows(s))
        if w["S"] >= 2:
            assert np.allclose(eng.rhat(), orc.rhat(), rtol=1e-12, atol=0)
            assert np.allclose(sd.allreduce_rhat(eng), orc.rhat(), rtol=1e-10, atol=0)  # two-stage form, 1 rank


@pytest.mark.parametrize("name,rounds", [("config1", 399), ("config2", 199), ("config3", 159), ("config5", 21)])
def test_philox_bit_exact(name, rounds):
    """Native Philox streams: same counter scheme, same Box-Muller arithmetic => bit-identical runs."""
    w = configs.workload(name, **SMALL[name])
    orc = oracle_for(w, seed=77, stack_offset=5)
    orc.init_chains()
    orc.run_rounds(rounds)
    with engine_for(w, seed=77, stack_offset=5) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        _compare(eng, orc, w, "philox %d rounds" % rounds)


GOLDEN = sorted(f[:-5] for f in os.listdir(os.path.join(HERE, "golden")) if f.startswith("oracle_") and f.endswith(".json"))


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_fixture(name):
    """Committed oracle output on the five workloads (tests/golden/make_golden.py): the GPU reproduces it bit for bit,
    the printed logger table included."""
    g = json.load(open(os.path.join(HERE, "golden", name + ".json")))
    w = configs.workload(g["workload"], **g["workload_args"])
    with engine_for(w, seed=g["seed"], windowRates=True) as eng:
        eng.init_chains()
        eng.step_rounds(g["rounds"])
        assert np.array_equal(eng.last_rows(), np.array(g["last_rows"]))
        sig, beta = eng.sigma_beta()
        assert np.array_equal(sig, np.array(g["sigma"])) and np.array_equal(beta, np.array(g["beta"]))
        assert np.array_equal(eng.rhat(), np.array(g["rhat"])) and eng.format_table() == g["table"]
        rows = eng.drain().reshape(g["rounds"] + 1, w["S"], w["d"] + 5)
        assert np.array_equal(rows[-3:, 0], np.array(g["cold_rows_stack0_tail"]))


@pytest.mark.parametrize("name", list(SMALL))
def test_likelihood_contract(name):
    """f(job, x) of the device functor == the oracle's target on random points."""
    w = configs.workload(name, **SMALL[name])
    rng = np.random.default_rng(3)
    x = rng.uniform(np.array(w["mn"]), np.array(w["mx"]), (64, w["d"]))
    with engine_for(w, sink="none") as eng:
        got = eng.eval_likelihood(x)
    want = np.array([[O.target_eval(w["oracle_target"], w["oracle_params"], w["d"], w["J"], j, xi) for j in range(w["J"])] for xi in x])
    assert np.array_equal(got, want)


def test_use_initial():
    """useInitial / initial (serverwrapper.hpp:76-86): every chain starts from the given point."""
    w = configs.workload("config1")
    init = [0.5, 0.25, -0.5, 1.0]
    orc = oracle_for(w, seed=9, initial=init)
    orc.init_chains()
    orc.run_rounds(29)
    from stateline_b200 import Engine
    cfg = configs.engine_config(w, seed=9, sink="host")
    cfg["useInitial"] = True
    cfg["initial"] = init
    with Engine(cfg, plugin=w["plugin"]) as eng:
        eng.init_chains()
        rows0 = eng.drain()
        assert np.array_equal(rows0[:, :4], np.tile(init, (w["S"], 1)))
        eng.step_rounds(29)
        assert np.array_equal(eng.last_rows(), orc.last_rows())


EDGE_SHAPES = [
    # S, T, d, J, rounds: smallest sizes, single temperature, ladders that fill a warp, ragged lane maps,
    # dimensions around the kernel boundaries (32 | 33: thread-per-chain vs warp-per-chain; 64 | 65 | 96 | 97: row groups)
    (1, 1, 1, 1, 25), (3, 1, 2, 2, 25), (2, 32, 3, 1, 31), (5, 3, 5, 4, 31), (3, 7, 32, 3, 21), (3, 7, 33, 3, 21),
    (2, 5, 64, 40, 15), (2, 9, 65, 2, 15), (1, 16, 97, 5, 12), (1, 32, 128, 1, 11),
]


@pytest.mark.parametrize("S,T,d,J,rounds", EDGE_SHAPES)
def test_edge_shapes_bit_exact(S, T, d, J, rounds):
    """Edge shapes through the generic likelihood (demo Gaussian accepts any d and J): every kernel family
    (thread-per-chain runtime-d, warp-per-chain with 2 / 3 / 4 row groups) against the oracle, bit for bit."""
    mn = [-3.0 - 0.1 * i for i in range(d)]
    mx = [2.0 + 0.2 * i for i in range(d)]
    w = dict(S=S, T=T, d=d, J=J, mn=mn, mx=mx, swapInterval=4, optAccept=0.3, optSwap=0.35, plugin="demo_gauss", payload=None,
             oracle_target=configs.T_DEMO_GAUSS, oracle_params=None)
    orc = oracle_for(w, seed=5, stack_offset=3)
    orc.init_chains()
    orc.run_rounds(rounds)
    with engine_for(w, seed=5, stack_offset=3) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        _compare(eng, orc, w, "edge shape S=%d T=%d d=%d J=%d" % (S, T, d, J))
        rows = eng.drain().reshape(rounds + 1, S, d + 5)
        for s in range(S):
            assert np.array_equal(rows[:, s], orc.cold_rows(s))


def _random_shapes(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        T = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 11, 16, 32]))
        d = int(rng.choice([1, 2, 3, 4, 7, 12, 20, 31, 32, 33, 40, 63, 64, 65, 100, 128]))
        S = int(rng.integers(1, 40 if d <= 32 else 4))
        J = int(rng.integers(1, 6))
        swap = int(rng.integers(1, 9))
        adapt = int(rng.choice([swap, swap, 1, 3, 7]))
        mrl = int(rng.choice([64, 64, 1, 2, 5]))
        rounds = int(rng.integers(8, 40 if d <= 32 else 14))
        out.append((S, T, d, J, swap, adapt, mrl, rounds, bool(rng.integers(0, 2)), int(rng.integers(0, 1000))))
    return out


@pytest.mark.parametrize("S,T,d,J,swap,adapt,mrl,rounds,use_initial,seed", _random_shapes(28, 2026))
def test_random_shapes_bit_exact(S, T, d, J, swap, adapt, mrl, rounds, use_initial, seed):
    """A seeded sweep over shapes and schedules: stack / ladder / dimension / job counts, swap interval, an adapt
    interval different from it, launches cut at 1, 2, 5 rounds, fixed initial state -- always the oracle's bits."""
    mn = [-2.0 - 0.05 * i for i in range(d)]
    mx = [1.5 + 0.1 * i for i in range(d)]
    w = dict(S=S, T=T, d=d, J=J, mn=mn, mx=mx, swapInterval=swap, optAccept=0.25, optSwap=0.4, plugin="demo_gauss", payload=None,
             oracle_target=configs.T_DEMO_GAUSS, oracle_params=None)
    initial = [0.1 * ((i % 5) - 2) for i in range(d)] if use_initial else None
    orc = oracle_for(w, seed=seed, stack_offset=seed % 7, adapt_interval=adapt, initial=initial)
    orc.init_chains()
    orc.run_rounds(rounds)
    cfg = configs.engine_config(w, seed=seed, sink="host", stackOffset=seed % 7, adaptInterval=adapt, maxRoundsPerLaunch=mrl,
                                hostRingRounds=64)
    if use_initial:
        cfg["useInitial"], cfg["initial"] = True, initial
    from stateline_b200 import Engine
    with Engine(cfg, plugin="demo_gauss") as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        _compare(eng, orc, w, "random shape S=%d T=%d d=%d J=%d swap=%d adapt=%d mrl=%d" % (S, T, d, J, swap, adapt, mrl))
        rows = eng.drain().reshape(rounds + 1, S, d + 5)
        for s in range(S):
            assert np.array_equal(rows[:, s], orc.cold_rows(s))


def test_one_round_launches_keep_their_rows():
    """Launches cut to a single round run back to back while the drain stream still copies the previous
    segment: the row segments must be distinct buffers (they once aliased, and a late copy picked up the next
    round's rows about four runs in five).  Five runs of the racy shape, every cold row against the oracle."""
    S, T, d, J, rounds = 37, 32, 4, 5, 27
    mn = [-2.0 - 0.05 * i for i in range(d)]
    mx = [1.5 + 0.1 * i for i in range(d)]
    w = dict(S=S, T=T, d=d, J=J, mn=mn, mx=mx, swapInterval=8, optAccept=0.25, optSwap=0.4, plugin="demo_gauss", payload=None,
             oracle_target=configs.T_DEMO_GAUSS, oracle_params=None)
    orc = oracle_for(w, seed=345, stack_offset=2, adapt_interval=3)
    orc.init_chains()
    orc.run_rounds(rounds)
    want = np.stack([orc.cold_rows(s) for s in range(S)], axis=1)
    from stateline_b200 import Engine
    for _ in range(5):
        cfg = configs.engine_config(w, seed=345, sink="host", stackOffset=2, adaptInterval=3, maxRoundsPerLaunch=1, hostRingRounds=64)
        with Engine(cfg, plugin="demo_gauss") as eng:
            eng.init_chains()
            eng.step_rounds(rounds)
            rows = eng.drain().reshape(rounds + 1, S, d + 5)
        assert np.array_equal(rows, want)


def test_zero_rounds_and_empty_run():
    w = configs.workload("config1")
    with engine_for(w, seed=1) as eng:
        eng.init_chains()
        eng.step_rounds(0)
        assert eng.rounds_done == 0
        assert eng.drain().shape == (w["S"], w["d"] + 5)  # only the initial rows


def _compare_all_rows(eng, orc, w, rounds):
    rows = eng.drain().reshape(rounds + 1, w["S"], w["d"] + 5)
    for s in range(w["S"]):
        assert np.array_equal(rows[:, s], orc.cold_rows(s)), "cold rows of stack %d" % s


# BASELINE.json's sizes, bit for bit against the oracle: the bench configuration itself (#3 at 8192 x 8 crosses two
# adapt points in 21 rounds and three in 31), #4 at 1024 x 8 and #5 at 8192 x 16 (one GPU's share).  At these sizes
# k_tier_adapt's reduce slots hold 32 / 4 / 32 stacks each, the grids are several waves deep and every block stages a
# realistic number of accepted chains -- none of which the small shapes above reach.
@pytest.mark.parametrize("name,S,rounds", [("config3", 8192, 21), ("config3", 8192, 31), ("config4", 1024, 11), ("config5", 8192, 4)])
def test_baseline_sizes_bit_exact(name, S, rounds):
    w = configs.workload(name, S=S)
    orc = oracle_for(w, seed=1234)
    orc.init_chains()
    orc.run_rounds(rounds)
    with engine_for(w, seed=1234, hostRingRounds=rounds + 2) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        _compare(eng, orc, w, "%s at S=%d after %d rounds" % (name, S, rounds))
        _compare_all_rows(eng, orc, w, rounds)
        # every chain's factor, not only three of them: the scale the next proposal uses is a function of all of L
        rng = np.random.default_rng(S + rounds)
        for chain in rng.integers(0, w["S"] * w["T"], 24):
            gL, gN, gcnt = eng.shaper(int(chain))
            oL, oN, ocnt = orc.shaper(int(chain))
            assert gcnt == ocnt and np.array_equal(gL, oL) and np.array_equal(gN, oN), chain


@pytest.mark.parametrize("S", [257, 1000, 4099])
def test_tier_reduce_across_many_stacks(S):
    """k_tier_adapt's cross-stack sum against the oracle's reduce_tree256 when a slot holds several stacks and S is
    not a multiple of 256: the tier models (mu_xx, mu_xy, weights, counts) after three adapt points, bit for bit."""
    d, T, J = 3, 4, 2
    w = dict(S=S, T=T, d=d, J=J, mn=[-2.0, -1.5, -3.0], mx=[2.5, 1.0, 3.0], swapInterval=5, optAccept=0.25, optSwap=0.4,
             plugin="demo_gauss", payload=None, oracle_target=configs.T_DEMO_GAUSS, oracle_params=None)
    orc = oracle_for(w, seed=S)
    orc.init_chains()
    orc.run_rounds(16)
    with engine_for(w, seed=S, hostRingRounds=20) as eng:
        eng.init_chains()
        eng.step_rounds(16)
        _compare(eng, orc, w, "tier reduce S=%d" % S)
        _compare_all_rows(eng, orc, w, 16)


@pytest.fixture(scope="module")
def tile_plugin():
    """The built-in targets compiled with -DSLGPU_TILE_KERNEL: the opt-in step kernel that keeps a stack group's factors in
    shared memory for the launch (include/slgpu_step_tile.cuh); stateline_b200/build.py builds it in-tree."""
    from stateline_b200 import build
    return build.build_targets_tile()


@pytest.mark.parametrize("name,S,rounds,kw", [("config1", 2, 61, {}), ("config2", 37, 45, {}), ("config3", 520, 31, {}),
                                             ("config3", 8192, 21, {}), ("config3", 5, 33, dict(rng="replay"))])
def test_tile_kernel_bit_exact(tile_plugin, name, S, rounds, kw, monkeypatch):
    """Same bits from the opt-in tile kernel: README demo (T = 5: idle lanes), double well (d = 2: two warps per group),
    config #3 at a ragged size and at the bench size, and replay mode."""
    monkeypatch.setenv("SLGPU_TARGETS_SO", tile_plugin)
    w = configs.workload(name, S=S)
    if kw.get("rng") == "replay":
        z, u_mh, u_swap, x_init = make_replay(w["S"], w["T"], w["d"], rounds, seed=4)
        orc = oracle_for(w, rng_mode=O.RNG_REPLAY)
        orc.set_replay(z, u_mh, u_swap, x_init)
    else:
        orc = oracle_for(w, seed=9)
    orc.init_chains()
    orc.run_rounds(rounds)
    with engine_for(w, seed=9, hostRingRounds=rounds + 2, **kw) as eng:
        if kw.get("rng") == "replay":
            eng.set_replay(z, u_mh, u_swap, x_init)
        eng.init_chains()
        for chunk in (1, 9, rounds - 10):
            eng.step_rounds(chunk)
        _compare(eng, orc, w, "tile kernel %s S=%d" % (name, S))
        _compare_all_rows(eng, orc, w, rounds)
