"""Generates tests/golden/ref_*.json from the REFERENCE'S OWN code (oracle/_ref: src/infer + src/db compiled
unmodified from /root/reference, see oracle/Makefile.ref and oracle/ref_harness.cpp).

    python tests/golden/make_ref_golden.py        # needs /root/reference (this container only)

Each fixture records one seeded reference run: the configuration, the start points, every variate the reference
drew (its own std::mt19937 / normal_distribution / uniform_real_distribution, seed pinned), and what the
reference computed from them: the last row of every chain, ChainArray's sigma/beta, the tier models, one proposal
shape, the TableLogger's counters, R-hat and printed table, and the text of the CSV files it wrote.  tests/test_ref_pin.py replays the draws through the oracle
(sequential schedule, glibc exp/log) and requires the same bits; the fixtures travel, the reference does not.
Floats are stored as hex so the JSON round trip is exact.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

CASES = {
    # src/bin/demo-config.json with the demo worker's likelihood, the README example (BASELINE config #1)
    "ref_config1_seed7": dict(S=2, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874,
                              mn=[-10, 0, -10, -10], mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS,
                              target_params=None, seed=7, wire=0, rounds=150),
    # the same through the "%f" wire text of requester.cpp:37 / minion.cpp:53
    "ref_config1_wire_seed9": dict(S=2, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874,
                                   mn=[-10, 0, -10, -10], mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS,
                                   target_params=None, seed=9, wire=1, rounds=100),
    # 2-D double well (BASELINE config #2, fewer stacks): swaps and mode hopping, swap interval 5
    "ref_doublewell_seed5": dict(S=3, T=8, J=2, swap_interval=5, opt_accept=0.234, opt_swap=0.3874,
                                 mn=[-3, -3], mx=[3, 3], target=O.TARGET_DOUBLE_WELL,
                                 target_params=None, seed=5, wire=0, rounds=120),
}


def hexa(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def gauss_fact_case():
    d = 8
    mu = np.linspace(-1.0, 1.0, d)
    inv_s = 1.0 / (2.0 ** ((np.arange(d) - 3.5) / 3.5))
    return dict(S=2, T=4, J=4, swap_interval=10, opt_accept=0.234, opt_swap=0.3874,
                mn=[-10] * d, mx=[10] * d, target=O.TARGET_GAUSS_FACT,
                target_params=np.concatenate([mu, inv_s]), seed=11, wire=0, rounds=100)


CASES["ref_gaussfact8_seed11"] = gauss_fact_case()


def run_case(c):
    S, T = c["S"], c["T"]
    d = len(c["mn"])
    C = S * T
    n = c["rounds"]
    x_init = np.random.default_rng(c["seed"]).uniform(-1.0, 1.0, (C, d))  # what VectorXd::Random spans
    out = tempfile.mkdtemp(prefix="slref_")
    r = R.RefRun(S, T, c["J"], c["swap_interval"], c["opt_accept"], c["opt_swap"], c["mn"], c["mx"], c["target"],
                 x_init, out, target_params=c["target_params"], seed=c["seed"], wire=c["wire"], max_rounds=n + 2)
    r.run_rounds(n - 1)
    r._L.slref_step(r.h, C - 1)
    table = r.step_print()  # the last step of round n-1 is one on which TableLogger prints (logging.cpp:50-87)
    assert r.unattributed() == 0
    lg = r.logger()
    fx = dict(
        generator="tests/golden/make_ref_golden.py (reference sources: src/infer/{sampler,chainarray,adaptive}.cpp, src/db/db.cpp)",
        config=dict(S=S, T=T, J=c["J"], swap_interval=c["swap_interval"], opt_accept=c["opt_accept"],
                    opt_swap=c["opt_swap"], mn=list(map(float, c["mn"])), mx=list(map(float, c["mx"])),
                    target=int(c["target"]),
                    target_params=None if c["target_params"] is None else hexa(c["target_params"]),
                    seed=c["seed"], wire=c["wire"], rounds=n),
        x_init=hexa(x_init), last_rows=hexa(r.last_rows()), table=table, rhat=hexa(lg["rhat"]),
        logger={k: hexa(lg[k]) for k in ("length", "min_energy", "cur_energy", "accepts", "swaps", "swap_attempts")},
    )
    s, b = r.sigma_beta()
    fx["sigma"], fx["beta"] = hexa(s), hexa(b)
    for which, name in ((0, "sigma_model"), (1, "beta_model")):
        xx, xy, w, cnt, values, rates = r.models(which)
        fx[name] = dict(mu_xx=hexa(xx), mu_xy=hexa(xy), weight=hexa(w), count=hexa(cnt), rates=hexa(rates))
    chain = C - 2
    L, N, cnt = r.shaper(chain)
    fx["shaper"] = dict(chain=chain, L=hexa(L), N=hexa(N), count=cnt)
    fx["lengths"] = [r.length(cn) for cn in range(C)]
    r.flush()
    fx["last_rows_after_flush"] = hexa(r.last_rows())
    assert r.unattributed() == 0
    # n + 1 rounds: round n holds the normals of the proposals that were in flight and the uniforms the flush drew
    z, u_mh, u_swap = r.draws(n + 1)
    fx.update(z=hexa(np.nan_to_num(z)), u_mh=hexa(np.nan_to_num(u_mh)), u_swap=hexa(np.nan_to_num(u_swap)),
              n_swap_draws=int((~np.isnan(u_swap[:n])).sum()))
    r.close()  # ~ChainArray writes the CSV files
    fx["csv"] = [open(os.path.join(out, "%d.csv" % s_)).read() for s_ in range(S)]
    return fx


if __name__ == "__main__":
    assert R.reference_present(), "the reference sources are needed to generate these fixtures"
    for name, c in CASES.items():
        fx = run_case(c)
        path = os.path.join(HERE, name + ".json")
        with open(path, "w") as f:
            json.dump(fx, f)
        print(name, os.path.getsize(path) // 1024, "KiB", "swap draws", fx["n_swap_draws"])
