"""Quick device-time probe of the step loop on one workload (development aid, not the bench)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stateline_b200 import Engine, configs

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config3")
ap.add_argument("--stacks", type=int, default=None)
ap.add_argument("--warm", type=int, default=200)
ap.add_argument("--rounds", type=int, default=500)
ap.add_argument("--sink", default="device")
ap.add_argument("--swap", type=int, default=0, help="override swapInterval")
ap.add_argument("--mrl", type=int, default=64, help="gpu.maxRoundsPerLaunch")
ap.add_argument("--shaper", default="reference", help="gpu.shaper: reference | welford (opt-in, not the reference's arithmetic)")
ap.add_argument("--rechol-every", type=int, default=1, help="gpu.recholEvery (welford shaper)")
ap.add_argument("--plugin", default=None, help="path of a plugin .so built from the same targets (development: A/B builds)")
ap.add_argument("--coarse", action="store_true", help="no per-launch events (lets programmatic dependent launch overlap)")
a = ap.parse_args()
if a.plugin:
    os.environ["SLGPU_TARGETS_SO"] = os.path.abspath(a.plugin)  # the built-in entry names, looked up in this build
w = configs.workload(a.workload, S=a.stacks)
if a.swap:
    w["swapInterval"] = a.swap
with Engine(configs.engine_config(w, sink=a.sink, overwrite=True, maxRoundsPerLaunch=a.mrl, shaper=a.shaper, **({"recholEvery": a.rechol_every} if a.shaper == "welford" else {})), plugin=w["plugin"], payload=w["payload"]) as e:
    e.init_chains()
    e.step_rounds(a.warm)
    e.sync()
    for rep in range(3):
        e.timer_start(per_launch=not a.coarse)
        e.step_rounds(a.rounds)
        ms = e.timer_stop()
        kms, kn = e.timer_step_kernel()
        e.sync()
        cs = w["S"] * w["T"] * a.rounds
        print(json.dumps({"workload": a.workload, "S": w["S"], "T": w["T"], "d": w["d"], "rounds": a.rounds, "ms": ms,
                          "us_per_round": 1e3 * ms / a.rounds, "kernel_us_per_round": 1e3 * kms / a.rounds, "launches": kn, "chain_steps_per_s": cs / (ms * 1e-3)}))
    s = e.summary()
    print(json.dumps({k: s[k] for k in ("accept_rate", "swap_rate", "sigma", "beta")}))
    print("rhat mean", sum(s["rhat"]) / max(1, len(s["rhat"])))
