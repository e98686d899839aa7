"""Rows per second of the csv sink (format + append to <stack>.csv) at a given formatter thread count."""
import json
import sys
import tempfile
import time

from stateline_b200 import Engine, configs


def main():
    S, rounds = 1024, 400
    out = {}
    for threads in (1, 4, 0):
        w = configs.workload("config3", S=S)
        with tempfile.TemporaryDirectory() as tmp:
            cfg = configs.engine_config(w, sink="csv", seed=1, csvThreads=threads, hostRingRounds=64)
            cfg["outputPath"] = tmp
            with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as eng:
                eng.init_chains()
                eng.step_rounds(40)
                eng.sync()
                t0 = time.perf_counter()
                eng.step_rounds(rounds)
                eng.sync()
                dt = time.perf_counter() - t0
        out["threads_%s" % (threads or "auto")] = {"rows_per_s": S * rounds / dt, "chain_steps_per_s": S * w["T"] * rounds / dt,
                                                  "seconds": dt}
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
