"""CPU tests that pin the oracle (oracle/sl_oracle.cpp) before anything is compared against it.

Pins: the reference's own known answers for this path (src/test/diagnostics.cpp:34,53,78; the swap case of
src/test/chainarray.cpp:104-125), hand-derived known answers for the pieces the reference leaves
untested (bouncyBounds, rank-1 Cholesky update, regression constructor, Philox), the committed golden
fixture, and the README's soft statistics (README.md:421-435)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- EPSR: the reference's three known answers (src/test/diagnostics.cpp) --------------------------
def test_epsr_constant_chains():
    x = np.zeros((10, 5, 1))
    assert O.epsr(x)[0] == 0.0  # diagnostics.cpp:34


def test_epsr_duplicate_linspaced_chains():
    chain = np.linspace(0.0, 1.0, 10)
    x = np.repeat(chain[:, None, None], 5, axis=1)
    assert abs(O.epsr(x)[0] - 0.948683298050513) < 1e-10  # diagnostics.cpp:53


def test_epsr_random_chains():
    x = np.array([[0.1415084, 0.8388452], [0.3565489, 0.8506014], [0.1773983, 0.2258481],
                  [0.6900898, 0.6106332], [0.0096742, 0.6742046]])
    assert abs(O.epsr(x[:, :, None])[0] - 1.340739719234503) < 1e-10  # diagnostics.cpp:78


# ---- swap formula: src/test/chainarray.cpp:104-125 (E=(666,667), beta=(0.1,0.11) => always accepted) ----
def test_swap_formula_case():
    # p = exp((E_l - E_h)(beta_l - beta_h)) = exp((666-667)(0.1-0.11)) = e^0.01 > 1 (chainarray.cpp:55-67)
    p = np.exp((666.0 - 667.0) * (0.1 - 0.11))
    assert p > 1.0
    # and through the sampler: a 1-stack, 2-temperature run whose swap uniform is 0.999999 still swaps
    S, T, d = 1, 2, 1
    n = 9
    z = np.zeros((n, S * T, d))
    u_mh = np.ones((n, S * T))        # u=1 never accepts a proposal: states stay at x_init
    u_swap = np.full((n, S * T), 0.999999)
    x_init = np.array([[-0.75], [0.5]])  # the colder chain starts with the higher energy => p > 1
    s = O.Sampler(S=S, T=T, J=1, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-1.0], mx=[1.0],
                  target=O.TARGET_DEMO_GAUSS, schedule=O.SCHED_BATCHED, rng_mode=O.RNG_REPLAY)
    s.set_replay(z, u_mh, u_swap, x_init)
    s.init_chains()
    rows0 = s.last_rows()
    s.run_rounds(n)  # round index 8 makes the chain length 10 => swap sweep
    rows1 = s.last_rows()
    e_l, e_h = rows0[0, d], rows0[1, d]
    assert 0.999999 < np.exp((e_l - e_h) * (rows0[0, d + 2] - rows0[1, d + 2]))
    assert rows1[0, 0] == rows0[1, 0] and rows1[1, 0] == rows0[0, 0]      # samples exchanged
    assert rows1[0, d] == rows0[1, d] and rows1[1, d] == rows0[0, d]      # energies exchanged
    assert rows1[0, d + 2] == rows0[0, d + 2]                               # beta stays with the rung (chainarray.cpp:176-178)
    assert rows1[0, d + 4] == 1 and rows1[1, d + 4] == 0                    # swapType on the lower chain only


# ---- bouncyBounds: src/infer/sampler.cpp:23-56 ---------------------------------------------------
@pytest.mark.parametrize("v,want", [
    (0.5, 0.5),          # inside
    (1.25, 0.75),        # one bounce off max
    (-1.5, -0.5),        # one bounce off min
    (3.5, -0.5),         # overstep 2.5 = 1 full width + 0.5: odd => min + 0.5
    (5.25, 0.75),        # overstep 4.25 = 2 widths + 0.25: even => max - 0.25
    (-3.5, 0.5),         # understep 2.5: odd => max - 0.5
    (1.0, 1.0), (-1.0, -1.0),
])
def test_bouncy_bounds(v, want):
    got = O.bouncy_bounds([v], [-1.0], [1.0])[0]
    assert got == pytest.approx(want, abs=1e-15)


# ---- ProposalShaper::update == chol(((n-1)/n) L L^T + s s^T / n): src/infer/adaptive.cpp:173-209 -----
def test_shaper_update_is_rank1_cholesky():
    rng = np.random.default_rng(5)
    d = 6
    A = rng.standard_normal((d, d))
    L0 = np.linalg.cholesky(A @ A.T + d * np.eye(d))
    step = rng.standard_normal(d)
    count = 1000.0
    prop_norm = 2.5
    L1, N1, cnt = O.shaper_update(L0, count, step, prop_norm)
    n = count + 1.0
    want = np.linalg.cholesky(((n - 1.0) / n) * (L0 @ L0.T) + np.outer(step, step) / n)
    assert cnt == n
    assert np.allclose(np.tril(L1), want, rtol=1e-12, atol=1e-14)
    scale = prop_norm / np.linalg.norm(np.tril(L1))
    assert np.allclose(N1, np.tril(L1) * scale + 1e-4 * np.eye(d), rtol=1e-13, atol=0)  # adaptive.cpp:200-204


# ---- RegressionAdapter ctor and the initial ladder: adaptive.cpp:31-61,118-132 -----------------------
def test_regression_init_and_initial_ladder():
    xx, xy, w, cnt = O.regression_init(0.0, 4.0)
    assert cnt == 10.0
    assert np.array_equal(w, xy)  # weight_ = mu_xy, not solved (adaptive.cpp:57)
    assert xx[4] == 10.0          # temp_variance_ on the side feature
    # default weights => log-ratio prediction (0.5 - opt)/2 per rung: beta_t = exp(-t (0.5 - optSwap)/2)
    opt = 0.3874
    s = O.Sampler(S=1, T=6, J=1, swap_interval=10, opt_accept=0.234, opt_swap=opt, mn=[-1.0], mx=[1.0],
                  target=O.TARGET_DEMO_GAUSS)
    s.init_chains()
    beta = s.sigma_beta()[1]
    assert np.allclose(beta, np.exp(-np.arange(6) * (0.5 - opt) / 2.0), rtol=1e-14)


def test_regression_update_matches_normal_equations():
    xx, xy, w, cnt = O.regression_init(-8.0, 4.0)
    rng = np.random.default_rng(2)
    X, Y = [], []
    for _ in range(200):
        lv, side, acc = rng.uniform(-3, 2), rng.uniform(0, 3), bool(rng.random() < 0.3)
        xx, xy, w, cnt = O.regression_update(xx, xy, w, cnt, -8.0, 4.0, lv, side, acc)
        X.append([-lv, side, 1.0])
        Y.append(float(acc))
    assert cnt == 210.0
    sol = np.linalg.solve(xx.reshape(3, 3), xy)
    assert np.allclose(w, sol, rtol=1e-8, atol=1e-10)  # colPivHouseholderQr().solve, adaptive.cpp:78


def test_qr_solve3_rank_deficient_is_finite():
    A = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 1.0, 1.0]])
    w = O.qr_solve3(A, np.array([1.0, 2.0, 0.5]))
    assert np.all(np.isfinite(w))
    assert np.allclose(A @ w, [1.0, 2.0, 0.5], atol=1e-12)


# ---- Philox4x32-10: published known-answer vectors (Random123 kat_vectors) ----------------------------
def test_philox_known_answers():
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_native_draws_are_standard():
    z = np.concatenate([O.draw_normals(11, c, r, 20) for c in range(40) for r in range(50)])
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 4.0 * np.sqrt(2.0 / z.size)
    u = np.array([O.draw_uniform(11, c, r, 1) for c in range(100) for r in range(50)])
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 4.0 / np.sqrt(12 * u.size)


# ---- CSV row: operator<<(ostream&, State), src/db/db.cpp:18-25 ---------------------------------------
def test_csv_row_format():
    row = np.array([0.25, -1234567.0, 1e-7, 2.5, 0.123456789, 1.0, 1.0, 2.0])
    assert O.csv_row(row, 3) == "0.25,-1.23457e+06,1e-07,2.5,0.123457,1,1,2"


# ---- golden fixture: oracle output recorded by tests/golden/make_golden.py -----------------------------
GOLDEN = sorted(f[:-5] for f in os.listdir(os.path.join(HERE, "golden")) if f.startswith("oracle_") and f.endswith(".json"))


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_fixture(name):
    from helpers import oracle_for
    from stateline_b200 import configs
    g = json.load(open(os.path.join(HERE, "golden", name + ".json")))
    w = configs.workload(g["workload"], **g["workload_args"])
    s = oracle_for(w, seed=g["seed"])
    s.init_chains()
    s.run_rounds(g["rounds"])
    assert np.array_equal(s.last_rows(), np.array(g["last_rows"]))
    assert np.array_equal(s.cold_rows(0)[-3:], np.array(g["cold_rows_stack0_tail"]))
    sig, beta = s.sigma_beta()
    assert np.array_equal(sig, np.array(g["sigma"])) and np.array_equal(beta, np.array(g["beta"]))
    assert np.array_equal(s.rhat(), np.array(g["rhat"])) and s.format_table() == g["table"]


def test_golden_covers_the_five_workloads():
    assert len(GOLDEN) == 5 and {json.load(open(os.path.join(HERE, "golden", n + ".json")))["workload"] for n in GOLDEN} == {
        "config1", "config2", "config3", "config4", "config5"}


# ---- schedules: the reference order (sequential) and the launch-granular one (batched) agree statistically ----
def test_readme_soft_statistics_both_schedules():
    """README.md:421-435 after ~16k samples per chain on the demo: accept ~0.23 on tiers 0-3, swap ~0.36-0.41,
    beta ~1/0.40/0.15/0.056/0.020, sigma ~0.21/0.34/0.56/0.93/e^4, R-hat ~1.0005."""
    for sched in (O.SCHED_SEQUENTIAL, O.SCHED_BATCHED):
        s = O.Sampler(S=2, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-10, 0, -10, -10],
                      mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS, schedule=sched, seed=3)
        s.init_chains()
        s.run_rounds(16000)
        sig, beta = s.sigma_beta()
        c = s.counters()
        acc = c["accepts"] / c["length"]
        swp = c["swaps"][:4] / c["swap_attempts"][:4]
        assert np.all(np.abs(acc[:4] - 0.234) < 0.02), acc
        assert np.all(np.abs(swp - 0.3874) < 0.05), swp
        assert np.allclose(beta[:5], [1.0, 0.40256, 0.15077, 0.05591, 0.01990], rtol=0.2)
        assert np.allclose(sig[:4], [0.20957, 0.33714, 0.56486, 0.93183], rtol=0.2)
        assert sig[4] == pytest.approx(np.exp(4.0), rel=1e-12)  # hottest tier pinned at the cap (serverwrapper.cpp:48)
        assert np.all(np.abs(s.rhat() - 1.0) < 0.01)


def test_posterior_moments_config1():
    """Truncated-Gaussian moments of the demo target E = 1.5 |x|^2 (J=3 jobs of 0.5|x|^2) on the demo box:
    dims 0 and 2 are N(0,1/3); dim 1 is half-normal on [0,10]; dim 3 is cut at +2 (3.46 sigma)."""
    s = O.Sampler(S=8, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-10, 0, -10, -10],
                  mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS, schedule=O.SCHED_BATCHED, seed=21)
    s.init_chains()
    s.run_rounds(30000)
    x = np.concatenate([s.cold_rows(k)[5000:, :4] for k in range(8)])
    sd = np.sqrt(1.0 / 3.0)
    assert abs(x[:, 0].mean()) < 0.03 and abs(x[:, 2].mean()) < 0.03
    assert abs(x[:, 0].std() - sd) < 0.03 and abs(x[:, 2].std() - sd) < 0.03
    assert abs(x[:, 1].mean() - sd * np.sqrt(2.0 / np.pi)) < 0.03  # half-normal mean
    assert x[:, 1].min() >= 0.0 and x[:, 3].max() <= 2.0


def test_logging_window_quirk():
    """RegressionAdapter's logging window (adaptive.cpp:80-90) pushes, then pops once the deque holds n_window_ = 1000
    flags, and divides by the size before the pop: the rate is sum(all)/k for k < 1000 and sum(last 999)/1000 after."""
    rng = np.random.default_rng(5)
    f = (rng.random(2600) < 0.3).astype(np.int32)
    got = O.window_rates(f)
    want = np.empty(f.size)
    for i in range(f.size):
        k = i + 1
        want[i] = f[:k].sum() / k if k < 1000 else f[k - 999:k].sum() / 1000.0
    assert np.array_equal(got, want)


def test_table_logger_format():
    """The text of TableLogger::update (logging.cpp:52-87): header literal, the dashed rule, one line per chain
    ("%4u %9lu" then eight "%10.5f " cells), a blank line between stacks, the convergence line."""
    s = O.Sampler(S=2, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-10, 0, -10, -10],
                  mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS, schedule=O.SCHED_BATCHED, seed=4)
    s.init_chains()
    before = s.format_table().split("\n")
    assert before[4].split() == ["0", "1", "inf", "0.00000", "1.00000", "0.00000", "1.00000", "1.00000", "0.00000", "0.00000"]
    s.run_rounds(30)
    lines = s.format_table().split("\n")
    assert lines[0] == "" and lines[1] == ""
    assert lines[2] == "  ID    Length    MinEngy   CurrEngy      Sigma     AcptRt  GlbAcptRt       Beta     SwapRt  GlbSwapRt"
    assert lines[3] == "-" * 102
    body = lines[4:4 + 11]
    assert body[5] == "" and all(len(b) == 4 + 1 + 9 + 1 + 8 * 11 + 0 for i, b in enumerate(body) if i != 5)
    c, (cur, ar, sr), (sig, beta) = s.counters(), s.logger(), s.sigma_beta()
    want = "%4d %9d %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f %10.5f " % (
        7, c["length"][7], c["min_energy"][7], cur[7], sig[7], ar[7], c["accepts"][7] / c["length"][7], beta[7], sr[7],
        c["swaps"][7] / c["swap_attempts"][7])
    assert body[8] == want
    assert lines[15] == "" and lines[16] == "Convergence test: %g (not converged)" % s.rhat().mean()
    assert lines[17:] == [""]


def test_row_order_rank1_sweep_is_bit_identical():
    """ProposalShaper::update (adaptive.cpp:184-197) sweeps columns; sweeping ROWS instead (row j: its off-diagonal
    entries k = 0..j-1 left to right, then its own head) applies the same operations to every element in the same order.
    The round-2 kernel plan (profiles/README.md) relies on this to walk the packed factor in memory order."""
    import math
    rng = np.random.default_rng(5)

    def row_order(L, count, step):
        d = L.shape[0]
        L = L.copy()
        count += 1.0
        sq = math.sqrt(count)
        x = [step[i] / sq for i in range(d)]
        scale = math.sqrt((count - 1.0) / count)
        c, s = [0.0] * d, [0.0] * d
        for j in range(d):
            xj = x[j]
            for k in range(j):
                Ljk = (L[j, k] * scale + s[k] * xj) / c[k]
                L[j, k] = Ljk
                xj = c[k] * xj - s[k] * Ljk
            V = max(1e-18, L[j, j] * scale)
            r = math.sqrt(V * V + xj * xj)
            c[j], s[j] = r / V, xj / V
            L[j, j] = r
        return L

    for d in (1, 2, 5, 20, 33, 50):
        for t in range(8):
            A = np.tril(rng.normal(size=(d, d))) * 0.3
            A[np.diag_indices(d)] = np.abs(A.diagonal()) + 0.1
            step = rng.normal(size=d) * rng.uniform(0.01, 3)
            Lo, _, _ = O.shaper_update(A, 1000 + t, step, 2.5)
            assert np.array_equal(np.tril(Lo), np.tril(row_order(A, 1000.0 + t, step)))
