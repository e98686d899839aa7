"""Where the end-to-end step loop spends its time on N ranks (development aid; run under torchrun).  Per rank:
  copy    pinned D2H bandwidth with every rank copying at once (5 MB and 16 MB buffers)
  device  the step loop with sink "device" (no rows leave the GPU)
  host    the step loop with sink "host" and the delta wire, segments released as they land, per-step host timings"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stateline_b200 import Engine, configs

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
out = {"rank": rank}
for mb in (5, 16):
    n = mb * (1 << 20) // 8
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    for _ in range(3):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        host.copy_(dev, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out["d2h_%dMB_GBps" % mb] = round(100 * mb * (1 << 20) / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
w = configs.workload("config3", S=8192)
K, R = 100, 10
for sink in ("device", "host"):
    cfg = configs.engine_config(w, sink=sink, device=lr, stackOffset=rank * 8192, seed=1234, hostRingRounds=16 * R)
    with Engine(cfg, plugin=w["plugin"], payload=w["payload"]) as e:
        e.init_chains()
        for _ in range(20):
            e.step_rounds(R)
            e.sync()
            while sink == "host" and e.next_segment(block=True) is not None:
                e.release_segment(keep_rows_state=False)
        if world > 1:
            dist.barrier()
        t_step = t_take = 0.0
        t0 = time.perf_counter()
        e.timer_start(per_launch=False)
        in_flight = 0
        for _ in range(K):
            a = time.perf_counter()
            e.step_rounds(R)
            b = time.perf_counter()
            in_flight += 1
            while sink == "host":
                sg = e.next_segment(block=in_flight >= 8)
                if sg is None:
                    break
                e.release_segment(keep_rows_state=False)
                in_flight -= 1
            c = time.perf_counter()
            t_step += b - a
            t_take += c - b
        ms = e.timer_stop()
        e.sync()
        t1 = time.perf_counter()
        out[sink] = {"device_ms_per_step": ms / K, "wall_ms_per_step": 1e3 * (t1 - t0) / K, "host_step_rounds_ms": 1e3 * t_step / K,
                     "host_take_ms": 1e3 * t_take / K}
print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
