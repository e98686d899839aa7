""" "gpu": {"shaper": "welford"} -- the opt-in variant BASELINE's north star names ("Welford rank-1 covariance update with
periodic re-Cholesky").  It is NOT the reference's arithmetic (ProposalShaper::update, src/infer/adaptive.cpp:173-209, is a
rank-1 Cholesky update on every accept), so there is no bit parity to ask for; what is checked instead:
  * re-factorised after every round it is the same adaptation: factors and chain states agree with the reference mode to
    rounding (the second moment C <- (n-1)/n C + s s^T / n is what L L^T obeys under the reference's update);
  * with the product setting (re-Cholesky every adaptInterval rounds) the sampler is statistically the same: posterior
    moments of the README demo target and the adapted accept / swap rates, with stated bounds;
  * it is deterministic and resumes bit for bit; unsupported shapes are refused."""
import numpy as np
import pytest

from stateline_b200 import Engine, SlgpuError, configs
from helpers import engine_for

pytestmark = pytest.mark.gpu


def test_refactorised_every_round_it_is_the_reference_update():
    w = configs.workload("config3", S=8)
    res = {}
    for shaper in ("reference", "welford"):
        with engine_for(w, seed=6, shaper=shaper, adaptInterval=1) as eng:  # an adapt point, hence a re-Cholesky, after every round
            eng.init_chains()
            eng.step_rounds(40)
            res[shaper] = (eng.last_rows(), [eng.shaper(c) for c in range(0, 64, 5)], eng.sigma_beta())
    (ra, fa, sa), (rb, fb, sb) = res["reference"], res["welford"]
    d = w["d"]
    assert np.array_equal(ra[:, d + 3:], rb[:, d + 3:])  # same decisions
    assert np.allclose(ra[:, :d + 1], rb[:, :d + 1], rtol=1e-9, atol=1e-9)
    assert np.allclose(sa[0], sb[0], rtol=1e-12) and np.allclose(sa[1], sb[1], rtol=1e-12)
    n_updated = 0
    for (La, Na, ca), (Lb, Nb, cb) in zip(fa, fb):
        assert ca == cb
        n_updated += ca > 1000
        assert np.allclose(La, Lb, rtol=1e-9, atol=1e-12) and np.allclose(Na, Nb, rtol=1e-9, atol=1e-12)
    assert n_updated >= 5


def _demo_stats(shaper, rounds=6000, S=48, **kw):
    w = configs.workload("config1", S=S)
    with engine_for(w, seed=17, shaper=shaper, hostRingRounds=rounds + 8, **kw) as eng:
        eng.init_chains()
        eng.step_rounds(rounds)
        rows = eng.drain().reshape(rounds + 1, S, w["d"] + 5)
        summ = eng.summary()
    x = rows[rounds // 3:, :, :w["d"]].reshape(-1, w["d"])
    return x, summ


def test_welford_samples_the_same_posterior():
    """README demo target (E = 3 * 0.5 |x|^2 over three job types => N(0, I/3) truncated to the box; dim 1 is half-normal,
    src/bin/demo-config.json:2-3).  Bounds: means within 0.03, standard deviations within 3 %, adapted accept rates of the
    tiers below the sigma cap within 0.03 of the 0.234 target and of the reference mode, swap rates within 0.05."""
    xw, sw = _demo_stats("welford")
    xr, sr = _demo_stats("reference")
    for x in (xw, xr):
        assert abs(x[:, 0].mean()) < 0.03 and abs(x[:, 2].mean()) < 0.03
        assert abs(x[:, 0].std() / 3 ** -0.5 - 1.0) < 0.03 and abs(x[:, 2].std() / 3 ** -0.5 - 1.0) < 0.03
        assert abs(x[:, 1].mean() - (2.0 / (3.0 * np.pi)) ** 0.5) < 0.03
    aw, ar = np.array(sw["accept_rate"]), np.array(sr["accept_rate"])
    assert np.all(np.abs(aw[:3] - 0.234) < 0.03) and np.all(np.abs(aw[:4] - ar[:4]) < 0.03)
    assert np.all(np.abs(np.array(sw["swap_rate"])[:-1] - np.array(sr["swap_rate"])[:-1]) < 0.05)
    assert max(sw["rhat"]) < 1.05


def test_welford_on_the_double_well_d2():
    """d = 2 (fewer matrix rows than re-Cholesky warps: two of the four warps own no row): the mode runs, adapts to the same
    accept rates as the reference mode and sees the same second moments of the bimodal target (bounds: 0.03 / 5 %)."""
    w = configs.workload("config2", S=64)
    rounds, res = 3000, {}
    for shaper in ("reference", "welford"):
        with engine_for(w, seed=11, shaper=shaper, hostRingRounds=rounds + 8) as eng:
            eng.init_chains()
            eng.step_rounds(rounds)
            rows = eng.drain().reshape(rounds + 1, 64, w["d"] + 5)
            res[shaper] = (rows[rounds // 3:, :, :2].reshape(-1, 2), eng.summary())
    (xr, sr), (xw, sw) = res["reference"], res["welford"]
    assert np.all(np.isfinite(xw))
    m2r, m2w = (xr ** 2).mean(axis=0), (xw ** 2).mean(axis=0)
    assert np.all(np.abs(m2w / m2r - 1.0) < 0.05), (m2r, m2w)
    assert np.all(np.abs(np.array(sw["accept_rate"])[:4] - np.array(sr["accept_rate"])[:4]) < 0.03)


def test_rechol_every_third_adapt_point_same_posterior_and_limits():
    """gpu.recholEvery = 3: the factor is refreshed every 30 rounds instead of every 10 (the jump log holds the accepts in
    between).  Same bounds as above; a product adaptInterval * recholEvery beyond the log's 32 slots is refused."""
    x, sm = _demo_stats("welford", recholEvery=3)
    assert abs(x[:, 0].mean()) < 0.03 and abs(x[:, 2].mean()) < 0.03
    assert abs(x[:, 0].std() / 3 ** -0.5 - 1.0) < 0.03 and abs(x[:, 2].std() / 3 ** -0.5 - 1.0) < 0.03
    assert abs(x[:, 1].mean() - (2.0 / (3.0 * np.pi)) ** 0.5) < 0.03
    assert np.all(np.abs(np.array(sm["accept_rate"])[:3] - 0.234) < 0.03) and max(sm["rhat"]) < 1.05
    w = configs.workload("config3", S=8)
    with pytest.raises(SlgpuError) as ei:
        engine_for(w, shaper="welford", recholEvery=4)  # 10 * 4 > 32
    assert ei.value.status == 1  # E_CONFIG


@pytest.mark.parametrize("every", [1, 3])
def test_welford_is_deterministic_and_resumes(tmp_path, every):
    w = configs.workload("config3", S=16)
    with engine_for(w, seed=3, shaper="welford", recholEvery=every) as eng:
        eng.init_chains()
        eng.step_rounds(47)
        want = (eng.last_rows(), eng.shaper(9)[0], eng.drain())
    ck = str(tmp_path / "w.ck")
    with engine_for(w, seed=3, shaper="welford", recholEvery=every) as eng:
        eng.init_chains()
        eng.step_rounds(23)
        first = eng.drain()
        eng.save_state(ck)
    with engine_for(w, seed=3, shaper="welford", recholEvery=every) as eng:
        eng.load_state(ck)
        eng.step_rounds(24)
        got = (eng.last_rows(), eng.shaper(9)[0], np.concatenate([first, eng.drain()]))
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_welford_refuses_what_it_does_not_support():
    w = configs.workload("config4", S=2)  # d = 100: warp-per-chain kernel
    with pytest.raises(SlgpuError) as ei:
        engine_for(w, shaper="welford")
    assert ei.value.status == 1  # E_CONFIG
    mn, mx = [-2.0] * 7, [2.0] * 7     # d = 7 is not in demo_gauss's list of specialised dimensions
    w7 = dict(S=4, T=4, d=7, J=2, mn=mn, mx=mx, swapInterval=5, optAccept=0.25, optSwap=0.4, plugin="demo_gauss", payload=None)
    with engine_for(w7, shaper="welford") as eng:
        eng.init_chains()
        with pytest.raises(SlgpuError) as ei:
            eng.step_rounds(12)  # the first adapt point asks the plugin for the re-Cholesky
        assert ei.value.status == 2  # E_PLUGIN
