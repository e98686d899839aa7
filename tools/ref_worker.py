"""One reference sampler process for bench.py's CPU legs (cpu_baseline, --impl reference).

Runs the REFERENCE'S OWN src/infer object code (oracle/_ref/libsl_ref.so: Sampler::step, ChainArray, the adapters, the
proposal shaper, the CSV writer -- see oracle/ref_harness.cpp) on `stacks` stacks of the bench workload, with the
likelihood evaluated by an in-process worker (no ZeroMQ).  The reference's sampler is single-threaded by design
(serverwrapper.cpp:97-98), so bench.py starts one of these per host core: N independent stateline servers, each with
its own tier models -- the same partition the GPU arm uses across GPUs.

Protocol (stdin/stdout, one line each way): "<rounds>" -> runs that many rounds (every chain steps once per round) and
answers with the elapsed seconds; "0" ends the process.  TEST / BENCH INFRASTRUCTURE ONLY.
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    idx, stacks, workload = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import numpy as np
    from oracle import ref as R
    from stateline_b200 import configs
    w = configs.workload(workload, S=stacks)
    x_init = np.random.default_rng(100 + idx).uniform(-1.0, 1.0, (w["S"] * w["T"], w["d"]))
    out = tempfile.mkdtemp(prefix="slref_bench_")
    run = R.RefRun(w["S"], w["T"], w["J"], w["swapInterval"], w["optAccept"], w["optSwap"], w["mn"], w["mx"],
                   w["oracle_target"], x_init, out, target_params=w["oracle_params"], seed=1234 + idx, max_rounds=0)
    sys.stdout.write("ready %d\n" % (w["S"] * w["T"]))
    sys.stdout.flush()
    for line in sys.stdin:
        n = int(line)
        if n <= 0:
            break
        t = time.perf_counter()
        run.run_rounds(n)
        sys.stdout.write("%.9f\n" % (time.perf_counter() - t))
        sys.stdout.flush()
    run.close()
    for f in os.listdir(out):
        os.unlink(os.path.join(out, f))
    os.rmdir(out)


if __name__ == "__main__":
    main()
