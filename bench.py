#!/usr/bin/env python
"""bench.py -- chain-steps/s of the parallel-tempering hot path on N B200s (BASELINE.json metric).

Workload: BASELINE config #3, the 20-D Gaussian factorised over 10 job types, 8192 stacks x 8
temperatures PER GPU (weak scaling; stacks are independent, no collective in the step loop), native
Philox draws, swapInterval 10.  One "step" = one swap interval = 10 rounds of every chain
(= one pt_step launch + the tier-model adapt kernels) = 655 360 chain steps per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # N>1: launched under torchrun by the driver
    python bench.py --impl reference ...                         # the reference's CPU path (oracle port) on host cores

One JSON line on stdout (rank 0).  `value` is device-timed (CUDA events on the engine's compute stream,
max over ranks) with the cold-chain rows written to an HBM ring; `e2e` runs the same rounds through the
C-ABI with sink "host": every row is copied to pinned host memory inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "config3"
ROUNDS_PER_STEP = 10  # = swapInterval: launches are cut at adapt points
STEADY_ROUNDS = 5000  # rounds behind the steady-state arm (the accept rates and the ladder have settled by then)
# likelihood flops per chain step of the five workloads
F_LIK = {"config1": 24.0, "config2": 12.0, "config3": 120.0, "config4": 10300.0, "config5": 588.0}


def flops_per_chain_step(d, T, f_lik, a):
    """SURVEY.md 8(d) closed form with the MEASURED mean accept rate a (the shaper update only runs on an accept):
    F = d(d+1) + 6d + F_lik + 226 + a (d(d+1)/2 + 3d(d-1) + d(d+1) + 4d) + 6d/T."""
    return d * (d + 1) + 6 * d + f_lik + 226 + a * (d * (d + 1) / 2 + 3 * d * (d - 1) + d * (d + 1) + 4 * d) + 6 * d / T


def algorithmic_bytes_per_chain_step(d, T, K):
    """SURVEY.md 8(d), factors on chip for a launch (d <= 32): state (d + d(d+1)/2 + 6) doubles loaded+stored once per
    launch of K rounds, plus the cold row (d+5 doubles) written once per cold-chain step."""
    state = (d + d * (d + 1) // 2 + 6) * 8 * 2 / K
    row = (d + 5) * 8 / T
    return state + row


def streamed_bytes_per_chain_step(d, T, a):
    """SURVEY.md 8(d), factors streamed from HBM every round (d > 32): d(d+1)/2 doubles read per step, read and written
    again on an accept, plus the state vector and the cold row."""
    nL = d * (d + 1) // 2
    return nL * 8 * (1 + 2 * a) + 2 * d * 8 + (d + 5) * 8 / T


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.001)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(n_workers, budget_s, stacks=64):
    """The reference's server + N-worker CPU path (oracle port, sequential schedule, FIFO job farm with
    the "%f" text round trip) on a bounded sample of the same workload: `stacks` stacks x 8 temperatures."""
    from oracle import oracle as O
    from stateline_b200 import configs
    w = configs.workload(WORKLOAD, S=stacks)
    kw = dict(S=w["S"], T=w["T"], J=w["J"], swap_interval=w["swapInterval"], opt_accept=w["optAccept"], opt_swap=w["optSwap"],
              mn=w["mn"], mx=w["mx"], target=w["oracle_target"], target_params=w["oracle_params"],
              schedule=O.SCHED_SEQUENTIAL, rng_mode=O.RNG_PHILOX, seed=1234)
    secs, steps = O.time_server_workers(n_workers, 20, **kw)  # calibration
    rounds = max(20, int(20 * budget_s / max(secs, 1e-6)))
    secs, steps = O.time_server_workers(n_workers, rounds, **kw)
    return steps / secs, "%d stacks x %d temps, %d rounds (%d chain steps) of config #3, %s" % (
        w["S"], w["T"], rounds, steps, "server + %d worker threads, FIFO job farm, %%f wire text" % n_workers if n_workers else "inline, 1 thread")


class ReferencePool:
    """N independent reference samplers, one process per host core (tools/ref_worker.py): the reference's own
    src/infer object code (oracle/_ref) with an in-process worker.  step(rounds) runs `rounds` rounds in every process
    at once and returns (wall seconds, chain steps)."""

    def __init__(self, n_procs, stacks_per_proc):
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        self.procs = [subprocess.Popen([sys.executable, "-u", os.path.join(here, "tools", "ref_worker.py"), str(i),
                                        str(stacks_per_proc), WORKLOAD], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                       text=True, bufsize=1) for i in range(n_procs)]
        self.chains = 0
        for p in self.procs:
            line = p.stdout.readline().split()
            if not line or line[0] != "ready":
                self.close()
                raise RuntimeError("reference worker did not start")
            self.chains += int(line[1])

    def step(self, rounds):
        t0 = time.perf_counter()
        for p in self.procs:
            p.stdin.write("%d\n" % rounds)
            p.stdin.flush()
        for p in self.procs:
            float(p.stdout.readline())
        return time.perf_counter() - t0, self.chains * rounds

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write("0\n")
                p.stdin.flush()
                p.stdin.close()
            except Exception:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=30)
            except Exception:
                p.kill()


def reference_available():
    """oracle/_ref (the reference's own sources, built where /root/reference exists; the built file travels)."""
    try:
        from oracle import ref as R
        return R.build() is not None
    except Exception:
        return False


REF_STACKS_PER_PROC = 8
REF_SAMPLE = ("%d independent stateline servers (one per host core), %d stacks x %d temps of config #3 each, %d rounds%s: "
              "the reference's own src/infer + src/db object code (oracle/_ref), in-process worker instead of ZeroMQ")


def cpu_baseline_reference(budget_s):
    """cpu_baseline from oracle/_ref on all host cores, about budget_s seconds."""
    cores = os.cpu_count() or 1
    pool = ReferencePool(cores, REF_STACKS_PER_PROC)
    try:
        pool.step(20)  # warm-up: adaptation settles, caches fill
        secs, _ = pool.step(20)
        rounds = max(20, int(20 * budget_s / max(secs, 1e-6)))
        secs, steps = pool.step(rounds)
    finally:
        pool.close()
    return steps / secs, cores, REF_SAMPLE % (cores, REF_STACKS_PER_PROC, 8, rounds, "")


def cpu_baseline_full(budget_s=8.0, stacks=8, attempts=4):
    """The reference AS SHIPPED: stateline::ServerWrapper + (nproc - 1) stateline::WorkerWrapper over real ZeroMQ sockets
    (oracle/_ref/ref_full: every reference source unmodified, oracle/Makefile.full), on `stacks` stacks x 8 temperatures
    of config #3.  Returns (chain-steps/s, cores, sample) or None when the binary is absent.  A run that exceeds its time
    limit (the reference's own start-up handshake can lose a worker's HELLO) is killed and retried; the leg raises after
    `attempts` tries, so it adds a bounded time to a bench run."""
    import socket
    import subprocess
    import tempfile
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(here, "oracle", "_ref", "ref_full")
    if not os.path.exists(exe):
        try:
            from oracle import ref as R
            if R.reference_present():
                subprocess.check_call(["make", "-C", os.path.join(here, "oracle"), "-f", "Makefile.full", "-j8",
                                       "REF=" + R.REFERENCE_ROOT], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            pass
    if not os.path.exists(exe):
        return None
    import numpy as np
    from stateline_b200 import configs
    w = configs.workload(WORKLOAD, S=stacks)
    cores = os.cpu_count() or 2
    n_workers = max(1, cores - 1)
    tmp = tempfile.mkdtemp(prefix="slfull_")
    np.asarray(w["oracle_params"], dtype=np.float64).tofile(os.path.join(tmp, "params.bin"))
    startup = 0.3 + 0.05 * n_workers  # the harness's own sleeps while it brings the workers up (oracle/ref_full.cpp)

    def run(rounds, limit):
        cfg = {"nStacks": w["S"], "nTemperatures": w["T"], "nSamplesTotal": w["S"] * rounds, "swapInterval": w["swapInterval"],
               "loggingRateSec": 1e6, "nJobTypes": w["J"], "outputPath": tmp, "optimalAcceptRate": w["optAccept"],
               "optimalSwapRate": w["optSwap"], "useInitial": False, "min": list(w["mn"]), "max": list(w["mx"])}
        with open(os.path.join(tmp, "cfg.json"), "w") as f:
            json.dump(cfg, f)
        for _ in range(attempts):
            with socket.socket() as sk:  # a free port for the delegator
                sk.bind(("127.0.0.1", 0))
                port = sk.getsockname()[1]
            p = subprocess.Popen([exe, os.path.join(tmp, "cfg.json"), str(port), str(n_workers), str(int(w["oracle_target"])),
                                  str(w["d"]), os.path.join(tmp, "params.bin")], stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True, cwd=tmp)
            try:
                out, _ = p.communicate(timeout=limit)
                return float(out.split()[0])
            except Exception:
                p.kill()
                p.wait()
        raise RuntimeError("the reference's server + workers did not finish in %d attempts" % attempts)

    try:
        # The server samples from the moment the first worker has said hello, i.e. while the harness is still bringing
        # the others up: a sample has to be several times longer than that start-up for "seconds - start-up" to mean
        # anything.  Grow the run until the work part is at least ~40 % of the budget, then take that run.
        rounds, secs = 200, 0.0
        for _ in range(6):
            secs = run(rounds, startup + 6.0 * budget_s + 10.0) - startup
            if secs >= 0.4 * budget_s or rounds >= 20000:
                break
            rounds = int(rounds * min(10.0, max(1.5, 0.7 * budget_s / max(secs, 0.05))))
    finally:
        for f in os.listdir(tmp):
            os.unlink(os.path.join(tmp, f))
        os.rmdir(tmp)
    steps = w["S"] * w["T"] * rounds
    return steps / secs, n_workers + 1, (
        "%d stacks x %d temps of config #3, %d rounds (%d chain steps): the reference as shipped -- ServerWrapper + %d "
        "WorkerWrapper over ZeroMQ (oracle/_ref/ref_full, all reference sources unmodified), chain initialisation included"
        % (w["S"], w["T"], rounds, steps, n_workers))


def ess_geyer(x):
    """Effective sample size of a 1-D series by Geyer's initial positive sequence."""
    import numpy as np
    n = x.size
    x = x - x.mean()
    f = np.fft.rfft(x, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n]
    if acf[0] <= 0:
        return float(n)
    acf = acf / acf[0]
    s = 0.0
    for k in range(0, n - 1, 2):
        g = acf[k] + acf[k + 1]
        if g <= 0:
            break
        s += g
    tau = max(2.0 * s - 1.0, 1e-12)
    return float(n / tau)


def _ess_of_rows(rows, d):
    """rows [rounds, stacks, d + 5] of cold chains -> ESS per cold sample: min over dims, mean over stacks."""
    import numpy as np
    rounds, stacks = rows.shape[0], rows.shape[1]
    per_stack = [min(ess_geyer(rows[:, s, i]) for i in range(d)) for s in range(stacks)]
    return float(np.mean(per_stack)) / rounds


def ess_per_cold_sample(device, stacks=128, warm=1000, rounds=4000, **gpu):
    """GPU arm: ESS per cold-chain sample from a side run of the same workload with fewer stacks; multiplied by cold
    samples/s it gives ESS/s."""
    from stateline_b200 import Engine, configs
    w = configs.workload(WORKLOAD, S=stacks)
    cfg = configs.engine_config(w, sink="host", device=device, hostRingRounds=rounds + 16, seed=99, **gpu)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(warm)
        e.sync()
        e.drain()
        e.step_rounds(rounds)
        rows = e.drain().reshape(rounds, stacks, w["d"] + 5)
    return _ess_of_rows(rows, w["d"])


def ess_per_cold_sample_reference(stacks=16, warm=1000, rounds=3000):
    """CPU arm: the same estimator on the cold rows the reference's own object code (oracle/_ref) produced."""
    import tempfile
    import numpy as np
    from oracle import ref as R
    from stateline_b200 import configs
    w = configs.workload(WORKLOAD, S=stacks)
    x_init = np.random.default_rng(7).uniform(-1.0, 1.0, (w["S"] * w["T"], w["d"]))
    out = tempfile.mkdtemp(prefix="slref_ess_")
    run = R.RefRun(w["S"], w["T"], w["J"], w["swapInterval"], w["optAccept"], w["optSwap"], w["mn"], w["mx"],
                   w["oracle_target"], x_init, out, target_params=w["oracle_params"], seed=99, max_rounds=0)
    try:
        run.run_rounds(warm + rounds)
        rows = np.stack([run.cold_rows(s)[-rounds:] for s in range(stacks)], axis=1)
    finally:
        run.close()
        for f in os.listdir(out):
            os.unlink(os.path.join(out, f))
        os.rmdir(out)
    return _ess_of_rows(rows, w["d"])


def reference_config(cores, T, d, J):
    """What the reference arm actually runs: a bounded sample of the workload, one server per host core."""
    return {"workload": "config3: 20-D Gaussian factorised over 10 job types -- bounded sample: %d independent stateline "
                        "servers x %d stacks x %d temperatures (the GPU arm runs 8192 stacks x %d per GPU; the cost per "
                        "chain step does not grow with the stack count)" % (cores, REF_STACKS_PER_PROC, T, T),
            "servers": cores, "stacks_per_server": REF_STACKS_PER_PROC, "stacks_total": cores * REF_STACKS_PER_PROC,
            "temperatures": T, "dims": d, "job_types": J, "swap_interval": 10, "rng": "the reference's own mt19937 streams"}


def run_reference_ref(args):
    """--impl reference through oracle/_ref: the reference's own object code on every host core."""
    from stateline_b200 import configs
    w = configs.workload(WORKLOAD, S=REF_STACKS_PER_PROC)
    cores = os.cpu_count() or 1
    pool = ReferencePool(cores, REF_STACKS_PER_PROC)
    try:
        secs, _ = pool.step(10)
        rounds_per_step = max(1, int(10 * args.ref_seconds / max(secs * (args.steps + args.warmup), 1e-9)))
        for _ in range(args.warmup):
            pool.step(rounds_per_step)
        tot_s, tot_steps = 0.0, 0
        for _ in range(args.steps):
            s_, n_ = pool.step(rounds_per_step)
            tot_s += s_
            tot_steps += n_
    finally:
        pool.close()
    value = tot_steps / tot_s
    sample = REF_SAMPLE % (cores, REF_STACKS_PER_PROC, w["T"], rounds_per_step, " per step")
    extra = {}
    try:  # beside it: the reference as shipped (server + workers over ZeroMQ), a few seconds
        full = cpu_baseline_full(budget_s=6.0)
        if full:
            extra["full_server_workers"] = {"value": full[0], "cores": full[1], "kind": "reference", "sample": full[2]}
    except Exception as e:
        sys.stderr.write("reference server + workers leg failed (%r)\n" % (e,))
    out = {
        "impl": "reference", "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": reference_config(cores, w["T"], w["d"], w["J"]),
        "cpu_baseline": dict({"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "reference", "sample": sample}, **extra),
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_ess:
        try:
            eps = ess_per_cold_sample_reference()
            out["ess"] = {"ess_per_cold_sample": eps, "ess_per_sec": eps * value / w["T"],
                          "estimator": "Geyer initial positive sequence, min over dims, mean over 16 stacks of a side run of the reference's own code"}
        except Exception as e:
            sys.stderr.write("reference ESS leg failed (%r)\n" % (e,))
    _emit(json.dumps(out))


def run_reference(args, rank, world):
    if rank != 0:
        return
    if reference_available() and not args.ref_port:
        try:
            return run_reference_ref(args)
        except Exception as e:  # fall back to the port rather than lose the arm
            sys.stderr.write("oracle/_ref arm failed (%r): falling back to the oracle port\n" % (e,))
    cores = os.cpu_count() or 1
    n_workers = max(1, cores - 1)
    from oracle import oracle as O
    from stateline_b200 import configs
    stacks = 64
    w = configs.workload(WORKLOAD, S=stacks)
    kw = dict(S=w["S"], T=w["T"], J=w["J"], swap_interval=w["swapInterval"], opt_accept=w["optAccept"], opt_swap=w["optSwap"],
              mn=w["mn"], mx=w["mx"], target=w["oracle_target"], target_params=w["oracle_params"],
              schedule=O.SCHED_SEQUENTIAL, rng_mode=O.RNG_PHILOX, seed=1234)
    # size one step so that (warmup + steps) steps take about --ref-seconds (60 s)
    secs, steps = O.time_server_workers(n_workers, 10, **kw)
    per_round = secs / 10
    rounds_per_step = max(1, int(args.ref_seconds / max(per_round * (args.steps + args.warmup), 1e-9)))
    for _ in range(args.warmup):
        O.time_server_workers(n_workers, rounds_per_step, **kw)
    tot_s, tot_steps = 0.0, 0
    for _ in range(args.steps):
        s, n = O.time_server_workers(n_workers, rounds_per_step, **kw)
        tot_s += s
        tot_steps += n
    value = tot_steps / tot_s
    sample = "%d stacks x %d temps of config #3, %d rounds per step, server + %d worker threads (oracle port of src/infer + FIFO job farm with %%f wire text)" % (
        stacks, w["T"], rounds_per_step, n_workers)
    _emit(json.dumps({
        "impl": "reference", "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config3: 20-D Gaussian factorised over 10 job types -- bounded sample: one server, %d stacks x %d "
                               "temperatures, %d worker threads (oracle port)" % (stacks, w["T"], n_workers),
                   "stacks_total": stacks, "temperatures": w["T"], "dims": w["d"], "job_types": w["J"], "swap_interval": 10},
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": n_workers + 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bench_config(T, d, J, S):
    return {"workload": "config3: 20-D Gaussian factorised over 10 job types, %d stacks x %d temperatures per GPU" % (S, T),
            "stacks_per_gpu": S, "temperatures": T, "dims": d, "job_types": J, "swap_interval": 10,
            "rounds_per_step": ROUNDS_PER_STEP, "rng": "philox4x32-10",
            "l2": "chain state per GPU (~140 MB, L factors 110 MB) exceeds the 126 MB L2 and every step rewrites it: no explicit flush"}


_json_out = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints its version line on
    stdout when NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the run and the JSON line goes
    to a private copy of the original stdout."""
    global _json_out
    if _json_out is None:
        sys.stdout.flush()
        _json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _json_out or sys.stdout
    out.write(line + "\n")
    out.flush()


def plugin_digest():
    """Source hash of the built likelihood plugin (stateline_b200/build.py): ties a stored ncu traffic figure to the
    kernels it was measured on."""
    try:
        from stateline_b200 import build
        return open(build.TARGETS_SO + ".srchash").read().strip()[:16]
    except Exception:
        return None


def timed_steps(e, K, R, per_launch=True):
    """K steps of R rounds on engine e: (ms, step-kernel ms, step-kernel launches, accept rate of the region per tier,
    kernel launches).  per_launch=False times the region with one event pair only: the per-launch pairs sit between the
    kernels and keep the step kernel from starting while k_tier_adapt still runs (programmatic dependent launch), so the
    headline figure is taken without them and the step kernel's own time in a second pass with them."""
    import numpy as np
    c0 = e.counters()
    l0 = e.launch_count
    e.timer_start(per_launch=per_launch)
    for _ in range(K):
        e.step_rounds(R)
    ms = e.timer_stop()
    kms, kn = e.timer_step_kernel()
    e.sync()
    c1 = e.counters()
    T = e.T
    acc = (c1["accepts"].astype(np.float64) - c0["accepts"]).reshape(-1, T).mean(0) / (K * R)
    return ms, kms, kn, acc, e.launch_count - l0


def side_config(name, S, local_rank, peaks, fp64_peak, steps=12, warm_rounds=199):
    """A parity-test configuration of BASELINE.json (#4, #5) timed on the device: not the bench line, but the only place
    the driver sees the d > 32 kernel."""
    import numpy as np
    from stateline_b200 import Engine, configs
    w = configs.workload(name, S=S)
    d, T, R = w["d"], w["T"], ROUNDS_PER_STEP
    cfg = configs.engine_config(w, sink="device", device=local_rank, seed=1234)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(warm_rounds)
        e.sync()
        ms, kms, kn, acc, _ = timed_steps(e, steps, R)
    a = float(np.mean(acc))
    chain_steps = w["S"] * T * steps * R
    launch_ms = kms / max(kn, 1)
    per_launch = w["S"] * T * R
    bpcs = streamed_bytes_per_chain_step(d, T, a) if d > 32 else algorithmic_bytes_per_chain_step(d, T, R)
    gbs = bpcs * per_launch / (launch_ms * 1e-3) / 1e9
    fl = flops_per_chain_step(d, T, F_LIK[name], a)
    tf = fl * per_launch / (launch_ms * 1e-3) / 1e12
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"workload": "%s: d=%d, %d stacks x %d temperatures" % (name, d, w["S"], T), "value": chain_steps / (ms * 1e-3),
            "unit": "chain-steps/s", "steps": steps, "rounds_after_init": warm_rounds, "ms_per_step": ms / steps,
            "kernel": "pt_step_warp_kernel" if d > 32 else "pt_step_fast_kernel", "launch_ms": launch_ms, "accept_rate_mean": a,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "bytes_per_chain_step": bpcs,
                         "bytes_model": "nL*8*(1+2a) + 2d*8 + (d+5)*8/T (factors streamed every round)" if d > 32 else "SURVEY 8(d), factors once per launch",
                         "fp64": {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "flops_per_chain_step": fl}}}


def run_group(args):
    """python bench.py --gpus N WITHOUT torchrun: one process drives the N devices through slgpu_group_* (the C++ host's
    own multi-GPU path, include/slgpu.h) -- the shape of one `stateline -c cfg` server process.  Same workload per GPU as
    the torchrun line; `value` from the engines' own CUDA-event timers (max over devices), `e2e` by the wall clock with
    the delta stream taken per device."""
    import numpy as np
    import torch
    from stateline_b200 import EngineGroup, build, configs
    build.build_all()
    if torch.cuda.device_count() < args.gpus:
        raise SystemExit("bench.py --gpus %d: only %d CUDA devices" % (args.gpus, torch.cuda.device_count()))
    N, K, W, R = args.gpus, args.steps, max(args.warmup, 3), ROUNDS_PER_STEP
    w = configs.workload(WORKLOAD, S=args.stacks * N)
    T, d, J = w["T"], w["d"], w["J"]
    chain_steps = w["S"] * T * K * R
    cfg = configs.engine_config(w, sink="device", seed=1234)
    with EngineGroup(cfg, plugin=w["plugin"], payload=w["payload"], devices=N) as g:
        g.init_chains()
        g.step_rounds(W * R - 1)
        g.sync()
        for e in g.engines:
            e.timer_start()
        for _ in range(K):
            g.step_rounds(R)
        ms = max(e.timer_stop() for e in g.engines)
        g.sync()
        rhat = g.rhat()
    cfg = configs.engine_config(w, sink="host", seed=1234, hostRingRounds=16 * R)
    with EngineGroup(cfg, plugin=w["plugin"], payload=w["payload"], devices=N) as g:
        g.init_chains()
        left = W * R - 1
        while left:
            n = min(left, 8 * R)
            g.step_rounds(n)
            left -= n
            g.sync()
            for e in g.engines:
                while e.next_segment(block=True) is not None:
                    e.release_segment(keep_rows_state=False)
        b0 = sum(e.d2h_bytes for e in g.engines)
        seen = 0
        t0 = time.perf_counter()
        for k in range(K):
            g.step_rounds(R)
            for e in g.engines:
                while True:
                    sg = e.next_segment(block=(k % 8 == 7))
                    if sg is None:
                        break
                    seen += int(np.count_nonzero(sg["flags"] & 1))
                    e.release_segment(keep_rows_state=False)
        g.sync()
        for e in g.engines:
            while True:
                sg = e.next_segment(block=True)
                if sg is None:
                    break
                seen += int(np.count_nonzero(sg["flags"] & 1))
                e.release_segment(keep_rows_state=False)
        e2e_s = time.perf_counter() - t0
        d2h = (sum(e.d2h_bytes for e in g.engines) - b0) // K
    _emit(json.dumps({
        "metric": "chain_steps_per_sec", "value": chain_steps / (ms * 1e-3), "unit": "chain-steps/s", "n_gpus": N, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(T, d, J, args.stacks), "launcher": "single process, slgpu_group (no torchrun)",
        "e2e": {"value": chain_steps / e2e_s, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h),
                "accepted_flags_seen": seen},
        "diagnostics": {"rhat_mean": float(np.mean(rhat)), "rhat_max": float(np.max(rhat))}}))


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stacks", type=int, default=8192, help="stacks per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the steady-state arm, configs #4 / #5 and the strong-scaling arm")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port (server + N worker threads) even when oracle/_ref exists")
    ap.add_argument("--ref-seconds", type=float, default=60.0, help="--impl reference: CPU seconds the whole run is sized for")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        return run_group(args)  # not under torchrun: the single-process multi-GPU host

    import numpy as np
    import torch
    from stateline_b200 import Engine, build, configs, measure_fp64_peak
    build.build_all()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    bound_cpus = None
    if world > 1:
        import torch.distributed as dist
        from stateline_b200.dist import bind_to_gpu_cpus
        bound_cpus = bind_to_gpu_cpus(local_rank)  # pinned row ring on the GPU's own NUMA node
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    w = configs.workload(WORKLOAD, S=args.stacks)
    S, T, d, J = w["S"], w["T"], w["d"], w["J"]
    K, W, R = args.steps, args.warmup, ROUNDS_PER_STEP
    chain_steps = world * S * T * K * R
    per_launch = S * T * R
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    def roofline_of(launch_ms, a_mean, fp64_peak):
        fl = flops_per_chain_step(d, T, F_LIK[WORKLOAD], a_mean)
        tf = fl * per_launch / (launch_ms * 1e-3) / 1e12
        alg = algorithmic_bytes_per_chain_step(d, T, R) * per_launch
        gbs = alg / (launch_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "launch_ms": launch_ms, "algorithmic_bytes_per_launch": alg, "chain_steps_per_launch": per_launch,
                "bytes_per_chain_step": algorithmic_bytes_per_chain_step(d, T, R),
                "fp64": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                         "flops_per_chain_step": fl, "accept_rate_used": a_mean,
                         "peak_source": "measured live: register-resident DFMA loop (slgpu_measure_fp64_peak)"}}

    # ---- device-resident run: rows go to the HBM ring only
    steady = None
    cfg = configs.engine_config(w, sink="device", device=local_rank, stackOffset=rank * S, seed=1234)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(W * R - 1)  # chain length starts at 1: after W*R-1 rounds the next launch starts a swap interval
        e.sync()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms, _, _, acc_tiers, launches = timed_steps(e, K, R, per_launch=False)
        barrier()
        clocks = sampler.finish()
        ms = max_over_ranks(ms)
        # the dominant kernel's own duration: a second pass with an event pair around every launch
        Kk = max(10, min(K, 40))
        ms_k, step_kernel_ms, step_kernel_n, _, _ = timed_steps(e, Kk, R, per_launch=True)
        summary = e.summary() if rank == 0 else None
        rhat_local = e.rhat_stage1()
        # end-of-run diagnostics: the only collective of the run (R-hat over all GPUs' stacks)
        sm, ns, nn = rhat_local
        if dist:
            t = torch.tensor(np.concatenate([sm, [ns]]), dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            sm, ns = t[:-1].cpu().numpy(), float(t[-1].item())
        b_part, w_part = e.rhat_stage2(sm / ns)
        if dist:
            t = torch.tensor(np.concatenate([b_part, w_part]), dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            b_part, w_part = t[:d].cpu().numpy(), t[d:].cpu().numpy()
        from stateline_b200 import rhat_finish
        rhat = rhat_finish(b_part, w_part, ns, nn)
        # ---- steady state: the same engine STEADY_ROUNDS rounds later (rank 0 reports; every rank runs it so that the
        # GPUs of a multi-GPU line stay in step)
        if not args.no_side:
            done = e.rounds_done
            e.step_rounds(STEADY_ROUNDS + (R - (done + STEADY_ROUNDS + 1) % R) % R)  # next launch starts a swap interval
            e.sync()
            barrier()
            Ks = max(10, min(K, 50))
            s_ms, s_kms, s_kn, s_acc, _ = timed_steps(e, Ks, R)
            s_ms = max_over_ranks(s_ms)
            steady = (s_ms, s_kms, s_kn, s_acc, Ks, e.rounds_done)
    value = chain_steps / (ms * 1e-3)

    # ---- end to end: same rounds through the C-ABI with the cold rows delivered to pinned host memory as the delta
    # stream (gpu.wire "delta") and taken by the host segment by segment
    cfg = configs.engine_config(w, sink="host", device=local_rank, stackOffset=rank * S, seed=1234, hostRingRounds=16 * R)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        left = W * R - 1
        while left:  # warm-up: its rows are taken and dropped, the ring holds 16 steps
            n = min(left, 8 * R)
            e.step_rounds(n)
            left -= n
            e.sync()
            while e.next_segment(block=True) is not None:
                e.release_segment(keep_rows_state=False)
        barrier()
        b0 = e.d2h_bytes
        rows0 = e.rows_emitted
        n_accept = 0
        n_changed = 0
        n_rows_seen = 0
        last_energy = 0.0
        t0 = time.perf_counter()
        in_flight = 0
        for _ in range(K):
            e.step_rounds(R)
            in_flight += 1
            while True:  # take what has landed; never run more than 8 steps ahead of the rows (the ring holds 16)
                sg = e.next_segment(block=in_flight >= 8)
                if sg is None:
                    break
                n_accept += int(np.count_nonzero(sg["flags"] & 1))
                n_changed += sg["count"]
                n_rows_seen += sg["flags"].size
                e.release_segment(keep_rows_state=False)
                in_flight -= 1
        t_enq = time.perf_counter()
        e.sync()                  # includes the D2H copies of the last segments
        while True:
            sg = e.next_segment(block=True)
            if sg is None:
                break
            n_accept += int(np.count_nonzero(sg["flags"] & 1))
            n_changed += sg["count"]
            n_rows_seen += sg["flags"].size
            last_energy = float(sg["payload"][:, d].sum())  # the result of the last step, read on the host
            e.release_segment(keep_rows_state=False)
        t1 = time.perf_counter()
        barrier()
        e2e_s = max_over_ranks(t1 - t0)
        rows = e.rows_emitted - rows0
        d2h_per_step = (e.d2h_bytes - b0) // K
        assert n_rows_seen == rows, (n_rows_seen, rows)
    e2e_value = chain_steps / e2e_s

    # ---- strong scaling (N > 1): the SAME total number of stacks as the 1-GPU line, split over the ranks
    strong = None
    if world > 1 and not args.no_side:
        from stateline_b200.dist import shard_stacks
        off, n_local = shard_stacks(args.stacks, rank, world)
        ws = configs.workload(WORKLOAD, S=n_local)
        cfg = configs.engine_config(ws, sink="device", device=local_rank, stackOffset=off, seed=1234)
        with Engine(cfg, plugin=ws["plugin"], payload=ws["payload"]) as e:
            e.init_chains()
            e.step_rounds(W * R - 1)
            e.sync()
            barrier()
            s2_ms, _, _, _, _ = timed_steps(e, K, R)
            barrier()
            s2_ms = max_over_ranks(s2_ms)
        strong = {"value": args.stacks * T * K * R / (s2_ms * 1e-3), "unit": "chain-steps/s", "stacks_total": args.stacks,
                  "stacks_per_gpu": n_local, "ms_per_step": s2_ms / K, "scaling": "strong"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (pt_step): HBM view (algorithmic bytes) and the fp64 pipe view
    fp64_peak = measure_fp64_peak(local_rank)
    step_launch_ms = step_kernel_ms / max(step_kernel_n, 1)
    a_mean = float(np.mean(acc_tiers))
    roofline = roofline_of(step_launch_ms, a_mean, fp64_peak)
    traffic, traffic_note = None, "no ncu capture of this build under profiles/"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "pt_step_traffic.json")))
        if tj.get("plugin_digest") == plugin_digest():
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = "ncu --set full of this build (profiles/pt_step_traffic.json: %s)" % tj.get("source", "")
        else:
            traffic_note = "profiles/pt_step_traffic.json was measured on another build of the kernels (digest %s, this build %s): not reported" % (
                tj.get("plugin_digest"), plugin_digest())
    except Exception:
        pass
    roofline.update({
        "traffic": traffic, "traffic_source": traffic_note, "peak_source": hbm_src, "kernel": "pt_step_fast_kernel<GaussFact,20,128>",
        "launches_timed": int(step_kernel_n), "share_of_step": step_kernel_ms / ms_k,
        "timing": "launch_ms from a second pass with CUDA-event pairs around every launch (%d launches); the headline region has "
                  "one pair only, so that programmatic dependent launch can overlap k_tier_adapt" % int(step_kernel_n),
        "note": "latency-bound: fp64 dependent-issue latency (~16 cycles) times the serial chains of the step; neither roof is close (DESIGN.md section 5, profiles/README.md)"})

    out = {
        "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(T, d, J, S), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h_per_step),
                "full_row_bytes_per_step": int(rows * (d + 5) * 8 // K),
                "note": "inputs are the resident chain state and on-device Philox draws: nothing to upload per step; the cold rows "
                        "reach pinned host memory as the delta stream (flag byte per row + the d+3 doubles of the rows whose cold "
                        "state changed; the host can rebuild every row exactly) inside the timed region; the host reads every flag "
                        "byte of every segment and the energy column of the last segment, and hands the ring space back "
                        "(slgpu_discard_segment: it does not rebuild rows)",
                "rows": int(rows), "rows_changed": int(n_changed), "accepted_flags_seen": int(n_accept), "checksum": last_energy,
                "host_loop_ms": {"enqueue_and_take": 1e3 * (t_enq - t0), "final_sync_and_take": 1e3 * (t1 - t_enq)},
                "host_cpus_bound": (len(bound_cpus) if bound_cpus else None)},
        "gpu_launches": int(launches), "roofline": roofline,
        "job_evals_per_sec": value * J, "cold_samples_per_sec": value / T,
        "diagnostics": {"rhat_mean": float(np.mean(rhat)), "rhat_max": float(np.max(rhat)),
                        "accept_rate": summary["accept_rate"], "swap_rate": summary["swap_rate"],
                        "accept_rate_timed_region": [float(v) for v in acc_tiers]},
        "phase": "transient: the timed region starts %d rounds after initialisation (hot tiers still at the sigma cap)" % (W * R),
    }
    if steady:
        s_ms, s_kms, s_kn, s_acc, Ks, s_rounds = steady
        out["steady_state"] = {
            "value": world * S * T * Ks * R / (s_ms * 1e-3), "unit": "chain-steps/s", "steps": Ks, "ms_per_step": s_ms / Ks,
            "rounds_after_init": int(s_rounds), "accept_rate": [float(v) for v in s_acc],
            "roofline": roofline_of(s_kms / max(s_kn, 1), float(np.mean(s_acc)), fp64_peak)}
    if strong:
        out["strong_scaling"] = strong
    if not args.no_side and world == 1:
        out["configs"] = {}
        for name, Sside in (("config4", 1024), ("config5", 8192)):
            try:
                out["configs"][name] = side_config(name, Sside, local_rank, peaks, fp64_peak)
            except Exception as ex:
                sys.stderr.write("side config %s failed (%r)\n" % (name, ex))
    if not args.no_side and world == 1:
        # the opt-in non-reference shaper (north star's "Welford + periodic re-Cholesky"): same workload, same timer, for context
        try:
            out["modes"] = {}
            for key, every in (("welford", 1), ("welford_rechol_every_3", 3)):
                cfg_w = configs.engine_config(w, sink="device", device=local_rank, seed=1234, shaper="welford", recholEvery=every)
                with Engine(cfg_w, plugin=w["plugin"], payload=w["payload"]) as e:
                    e.init_chains()
                    e.step_rounds(W * R - 1)
                    e.sync()
                    w_ms, _, _, w_acc, _ = timed_steps(e, K, R, per_launch=False)
                out["modes"][key] = {
                    "value": S * T * K * R / (w_ms * 1e-3), "unit": "chain-steps/s", "ms_per_step": w_ms / K, "rechol_every_adapt_points": every,
                    "accept_rate_timed_region": [float(v) for v in w_acc],
                    "ess_per_cold_sample": None if args.no_ess else ess_per_cold_sample(local_rank, shaper="welford", recholEvery=every),
                    "note": "gpu.shaper \"welford\": accepted jumps are logged, folded into a second-moment matrix and re-factorised at every "
                            "n-th adapt point (rechol_kernel); NOT the reference's arithmetic (no bit parity; tests/test_welford_gpu.py "
                            "validates it statistically); the headline above is the reference-exact mode"}
        except Exception as ex:
            sys.stderr.write("welford mode arm failed (%r)\n" % (ex,))
    if not args.no_ess:
        eps = ess_per_cold_sample(local_rank)
        out["ess"] = {"ess_per_cold_sample": eps, "ess_per_sec": eps * value / T,
                      "estimator": "Geyer initial positive sequence, min over dims, mean over 128 stacks of a side run"}
    if not args.no_cpu_baseline and world == 1:  # the CPU legs run at N=1 only
        cores = os.cpu_count() or 1
        nwk = max(1, cores - 1)
        base = None
        if reference_available():
            # the reference's own object code (oracle/_ref), one independent server per host core
            try:
                v, ncores, sample = cpu_baseline_reference(args.cpu_budget)
                base = {"value": v, "unit": "chain-steps/s", "cores": ncores, "kind": "reference", "sample": sample}
                if not args.no_ess:
                    eps_ref = ess_per_cold_sample_reference()
                    base["ess_per_cold_sample"] = eps_ref
                    base["ess_per_sec"] = eps_ref * v / T
            except Exception as e:
                sys.stderr.write("oracle/_ref baseline failed (%r): using the oracle port\n" % (e,))
        # the oracle port in the reference's deployment shape: one server, N worker threads, FIFO job farm, %f wire text
        vp, sp = cpu_baseline(nwk, args.cpu_budget if base is None else 6.0)
        port = {"value": vp, "cores": nwk + 1, "kind": "port", "sample": sp}
        if base is None:
            base = dict(port, unit="chain-steps/s")
        else:
            base["port_server_workers"] = port
        out["cpu_baseline"] = base
        try:
            full = cpu_baseline_full()
            if full:
                out["cpu_baseline"]["full_server_workers"] = {"value": full[0], "cores": full[1], "kind": "reference", "sample": full[2]}
        except Exception as e:
            sys.stderr.write("reference server + workers leg failed (%r)\n" % (e,))
        vi, si = cpu_baseline(0, 4.0)
        out["cpu_baseline"]["inline_1_thread"] = {"value": vi, "kind": "port", "sample": si}
    _emit(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
