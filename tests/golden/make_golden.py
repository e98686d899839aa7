"""Generates tests/golden/oracle_config1_seed7.json from the oracle (run from the repo root).

The reference itself cannot be built in this image (Eigen and zmq.h are absent) and its own tests hold
no sampler vectors, so this fixture records the ORACLE's output on the README demo config with native
Philox draws.  It pins the oracle against silent drift (any later change to oracle/ or to
include/slgpu_math.h that alters a single bit fails tests/test_oracle.py::test_golden_fixture) and gives
the GPU tests a run-time-independent target."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

ROUNDS = 199
s = O.Sampler(S=2, T=5, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-10, 0, -10, -10], mx=[10, 10, 10, 2],
              target=O.TARGET_DEMO_GAUSS, schedule=O.SCHED_BATCHED, rng_mode=O.RNG_PHILOX, seed=7)
s.init_chains()
s.run_rounds(ROUNDS)
sig, beta = s.sigma_beta()
out = {"rounds": ROUNDS, "last_rows": s.last_rows().tolist(), "cold_rows_stack0_tail": s.cold_rows(0)[-3:].tolist(),
       "sigma": sig.tolist(), "beta": beta.tolist(), "rhat": s.rhat().tolist()}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_config1_seed7.json"), "w"), indent=1)
print("wrote golden fixture, rounds =", ROUNDS)
