"""Generates the fixtures under tests/golden/ from the oracle (run from the repo root: python tests/golden/make_golden.py).

These fixtures (oracle_*.json) record the ORACLE's output with native Philox draws on small instances of
the five BASELINE workloads (the README demo config at its real size).  They pin the oracle against silent drift
(any later change to oracle/ or to include/slgpu_math.h that alters a single bit fails
tests/test_oracle.py::test_golden_fixture) and give the GPU tests a run-time-independent target.  The fixtures recorded
from the REFERENCE'S OWN code (ref_*.json) are made by make_ref_golden.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_for  # noqa: E402
from stateline_b200 import configs  # noqa: E402

# name -> (workload, kwargs of configs.workload, seed, rounds)
FIXTURES = {
    "oracle_config1_seed7": ("config1", {}, 7, 199),
    "oracle_config2_seed5": ("config2", {"S": 4}, 5, 83),
    "oracle_config3_seed11": ("config3", {"S": 4}, 11, 57),
    "oracle_config4_seed13": ("config4", {"S": 2}, 13, 26),
    "oracle_config5_seed17": ("config5", {"S": 2}, 17, 31),
}


def record(name):
    wl, kw, seed, rounds = FIXTURES[name]
    w = configs.workload(wl, **kw)
    s = oracle_for(w, seed=seed)
    s.init_chains()
    s.run_rounds(rounds)
    sig, beta = s.sigma_beta()
    return {"workload": wl, "workload_args": kw, "seed": seed, "rounds": rounds, "last_rows": s.last_rows().tolist(),
            "cold_rows_stack0_tail": s.cold_rows(0)[-3:].tolist(), "sigma": sig.tolist(), "beta": beta.tolist(),
            "rhat": s.rhat().tolist(), "table": s.format_table()}


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in FIXTURES:
        json.dump(record(name), open(os.path.join(here, name + ".json"), "w"), indent=1)
        print("wrote", name)
