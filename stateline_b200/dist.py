"""Multi-GPU host logic: one process per GPU, stacks sharded contiguously, no collective in the step loop.

Stacks never exchange state (src/infer/sampler.cpp:173 swaps only id <-> id+1 inside a stack), so GPU g owns
stacks [g*S/G, (g+1)*S/G) with its own tier models -- the same thing as G stateline servers each running
S/G stacks.  The only collectives are the end-of-run diagnostics below (two small all-reduces for R-hat, one
for the counters), issued through torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_cpus(device_index):
    """Pin this process (and the threads it spawns later: the engine's drain thread) to the host cores next to
    GPU `device_index`, so that the pinned row ring is first-touched on the GPU's own NUMA node and the D2H
    row stream does not cross the socket interconnect.  Call it before creating the engine.  Returns the
    core set applied, or None when the topology cannot be read (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            tok = visible.split(",")[device_index].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(tok))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:], rest)
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_stacks(total_stacks, rank, world):
    """Contiguous shard of the stack range: returns (stack_offset, n_stacks) for this rank."""
    base, rem = divmod(int(total_stacks), int(world))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def _allreduce_sum(vec, group=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(vec, dtype=np.float64)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def allreduce_rhat(engine, group=None):
    """EPSRDiagnostic::rHat (src/infer/diagnostics.cpp:49-71) over the stacks of every rank.
    `engine` provides rhat_stage1() -> (sum_m[d], n_stacks, n_samples) and rhat_stage2(mean) -> (sum_dm2[d], sum_s[d])."""
    from .engine import rhat_finish
    sum_m, n_stacks, n_samples = engine.rhat_stage1()
    red = _allreduce_sum(np.concatenate([sum_m, [n_stacks]]), group)
    sum_m, m_total = red[:-1], float(red[-1])
    b, w = engine.rhat_stage2(sum_m / m_total)
    d = len(b)
    red = _allreduce_sum(np.concatenate([b, w]), group)
    return rhat_finish(red[:d], red[d:], m_total, n_samples)


def allreduce_rates(engine, group=None):
    """Per-tier global accept and swap rates (GlbAcptRt / GlbSwapRt of src/infer/logging.cpp:61-87) over all ranks."""
    c = engine.counters()
    T = engine.T
    acc = c["accepts"].reshape(-1, T).sum(0).astype(np.float64)
    length = c["length"].reshape(-1, T).sum(0).astype(np.float64)
    sw = c["swaps"].reshape(-1, T).sum(0).astype(np.float64)
    att = c["swap_attempts"].reshape(-1, T).sum(0).astype(np.float64)
    red = _allreduce_sum(np.concatenate([acc, length, sw, att]), group).reshape(4, T)
    return red[0] / red[1], red[2] / np.maximum(red[3], 1.0)
