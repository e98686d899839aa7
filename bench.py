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
# algorithmic fp64 flops per chain step, SURVEY.md 8(d) closed form at d=20, J=10, a=0.234, T=8
FLOPS_PER_CHAIN_STEP = 1330.0


def algorithmic_bytes_per_chain_step(d, T, K):
    """state (d + d(d+1)/2 + 6) doubles loaded+stored once per launch of K rounds, plus
    the cold row (d+5 doubles) written once per cold-chain step."""
    state = (d + d * (d + 1) // 2 + 6) * 8 * 2 / K
    row = (d + 5) * 8 / T
    return state + row


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


def cpu_baseline_full(budget_s=8.0, stacks=8, timeout_s=20.0):
    """The reference AS SHIPPED: stateline::ServerWrapper + (nproc - 1) stateline::WorkerWrapper over real ZeroMQ sockets
    (oracle/_ref/ref_full: every reference source unmodified, oracle/Makefile.full), on `stacks` stacks x 8 temperatures
    of config #3.  Returns (chain-steps/s, cores, sample) or None when the binary is absent.  The reference occasionally
    stalls at start-up (the workers' hello is never registered: about one run in ten here; not investigated further): a run that exceeds its time limit is killed and retried once; the leg
    gives up (raises) after that, so it can add at most about a minute to a bench run."""
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

    def run(rounds, limit):
        cfg = {"nStacks": w["S"], "nTemperatures": w["T"], "nSamplesTotal": w["S"] * rounds, "swapInterval": w["swapInterval"],
               "loggingRateSec": 1e6, "nJobTypes": w["J"], "outputPath": tmp, "optimalAcceptRate": w["optAccept"],
               "optimalSwapRate": w["optSwap"], "useInitial": False, "min": list(w["mn"]), "max": list(w["mx"])}
        with open(os.path.join(tmp, "cfg.json"), "w") as f:
            json.dump(cfg, f)
        for _ in range(2):
            with socket.socket() as sk:  # a free port for the delegator
                sk.bind(("127.0.0.1", 0))
                port = sk.getsockname()[1]
            p = subprocess.Popen([exe, os.path.join(tmp, "cfg.json"), str(port), str(n_workers), str(int(w["oracle_target"])),
                                  str(w["d"]), os.path.join(tmp, "params.bin")], stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True, cwd=tmp)
            try:
                out, _ = p.communicate(timeout=limit)
                secs = float(out.split()[0])
                return secs
            except Exception:
                p.kill()
                p.wait()
        raise RuntimeError("the reference's server + workers did not finish")

    startup = 0.3 + 0.05 * n_workers  # the harness's own sleeps while it brings the workers up (oracle/ref_full.cpp)
    t_small = run(20, timeout_s)
    per_round = max((t_small - startup) / 20, 1e-4)
    rounds = int(max(20, min(2000, budget_s / per_round)))
    secs = run(rounds, max(timeout_s, 3.0 * budget_s)) - startup
    steps = w["S"] * w["T"] * rounds
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)
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


def ess_per_cold_sample(device, stacks=128, warm=1000, rounds=4000):
    """ESS per cold-chain sample (min over dims, mean over stacks) from a side run of the same workload
    with fewer stacks; multiplied by cold samples/s it gives ESS/s."""
    import numpy as np
    from stateline_b200 import Engine, configs
    w = configs.workload(WORKLOAD, S=stacks)
    cfg = configs.engine_config(w, sink="host", device=device, hostRingRounds=rounds + 16, seed=99)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(warm)
        e.sync()
        e.drain()
        e.step_rounds(rounds)
        rows = e.drain().reshape(rounds, stacks, w["d"] + 5)
    per_stack = [min(ess_geyer(rows[:, s, i]) for i in range(w["d"])) for s in range(stacks)]
    return float(np.mean(per_stack)) / rounds


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
    _emit(json.dumps({
        "impl": "reference", "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(w["T"], w["d"], w["J"], 8192),
        "cpu_baseline": dict({"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "reference", "sample": sample}, **extra),
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
        "config": bench_config(w["T"], w["d"], w["J"], 8192),
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

    w = configs.workload(WORKLOAD, S=args.stacks)
    S, T, d, J = w["S"], w["T"], w["d"], w["J"]
    K, W, R = args.steps, args.warmup, ROUNDS_PER_STEP
    chain_steps = world * S * T * K * R

    # ---- device-resident run: rows go to the HBM ring only
    cfg = configs.engine_config(w, sink="device", device=local_rank, stackOffset=rank * S, seed=1234)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(W * R - 1)  # chain length starts at 1: after W*R-1 rounds the next launch starts a swap interval
        e.sync()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = e.launch_count
        e.timer_start()
        for _ in range(K):
            e.step_rounds(R)
        ms = e.timer_stop()
        step_kernel_ms, step_kernel_n = e.timer_step_kernel()
        e.sync()
        barrier()
        clocks = sampler.finish()
        launches = e.launch_count - l0
        ms = max_over_ranks(ms)
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
    value = chain_steps / (ms * 1e-3)

    # ---- end to end: same rounds through the C-ABI with every cold row copied to pinned host memory
    cfg = configs.engine_config(w, sink="host", overwrite=True, device=local_rank, stackOffset=rank * S, seed=1234,
                                hostRingRounds=16 * R)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        e.step_rounds(W * R - 1)
        e.sync()
        barrier()
        t0 = time.perf_counter()
        rows0 = e.rows_emitted
        for _ in range(K):
            e.step_rounds(R)
        e.sync()                  # includes the D2H copies of the last segments
        spans = e.peek()          # the result of the last step, read on the host
        last = spans[-1][-S:] if spans else np.zeros((0, d + 5))  # the rows of the last round: the result the host reads
        checksum = float(last[:, d].sum())
        t1 = time.perf_counter()
        barrier()
        e2e_s = max_over_ranks(t1 - t0)
        rows = e.rows_emitted - rows0
    e2e_value = chain_steps / e2e_s
    d2h_per_step = rows * (d + 5) * 8 // K

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (pt_step): fp64 pipe, plus the HBM view
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak = measure_fp64_peak(local_rank)
    # the dominant kernel is pt_step (one launch per step); its launches carry their own event pairs
    step_launch_ms = step_kernel_ms / max(step_kernel_n, 1)
    per_gpu_steps_per_launch = S * T * R
    achieved_tf = FLOPS_PER_CHAIN_STEP * per_gpu_steps_per_launch / (step_launch_ms * 1e-3) / 1e12
    alg_bytes = algorithmic_bytes_per_chain_step(d, T, R) * per_gpu_steps_per_launch
    achieved_gbs = alg_bytes / (step_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "pt_step_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
        "traffic": traffic, "peak_source": hbm_src, "kernel": "pt_step_fast_kernel<GaussFact,20,128>",
        "launch_ms": step_launch_ms, "launches_timed": int(step_kernel_n), "share_of_step": step_kernel_ms / ms,
        "algorithmic_bytes_per_launch": alg_bytes, "chain_steps_per_launch": per_gpu_steps_per_launch,
        "bytes_per_chain_step": algorithmic_bytes_per_chain_step(d, T, R),
        "note": "latency-bound in round 1: neither roof is close (DESIGN.md section 5, profiles/README.md)",
        "fp64": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                 "flops_per_chain_step": FLOPS_PER_CHAIN_STEP,
                 "peak_source": "measured live: register-resident DFMA loop (slgpu_measure_fp64_peak)"},
    }

    out = {
        "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(T, d, J, S), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h_per_step),
                "note": "inputs are the resident chain state and on-device Philox draws: nothing to upload per step; every cold-chain row is copied D2H to pinned memory inside the timed region",
                "checksum": checksum, "host_cpus_bound": (len(bound_cpus) if bound_cpus else None)},
        "gpu_launches": int(launches), "roofline": roofline,
        "job_evals_per_sec": value * J, "cold_samples_per_sec": value / T,
        "diagnostics": {"rhat_mean": float(np.mean(rhat)), "rhat_max": float(np.max(rhat)),
                        "accept_rate": summary["accept_rate"], "swap_rate": summary["swap_rate"]},
    }
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
