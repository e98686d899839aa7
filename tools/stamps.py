"""Development: per-block clock64 stamps of one round of pt_step_tile_kernel (plugin built with -DSLGPU_STAMPS=<round>).
usage: SLGPU_TARGETS_SO=<stamped plugin> python tools/stamps.py [stacks]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (loads the CUDA runtime the copy below uses)
from stateline_b200 import Engine, configs

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = configs.workload("config3", S=S)
rt = C.CDLL("libcudart.so.12")
NAMES = ["draws", "gemv+bounce", "likelihood", "MH", "column sweep", "swap+tail"]
with Engine(configs.engine_config(w, sink="device", windowRates=True), plugin=w["plugin"], payload=w["payload"]) as e:
    e.init_chains()
    e.step_rounds(209)
    e.sync()
    e.step_rounds(10)  # one launch
    e.sync()
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "stateline_b200", "libslgpu.so"))
    lib.slgpu_debug_round_flags.restype = C.c_void_p
    lib.slgpu_debug_round_flags.argtypes = [C.c_void_p]
    dev = lib.slgpu_debug_round_flags(e.h)
    nblk = (S + 3) // 4
    host = np.zeros(nblk * 32, dtype=np.int64)
    rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(dev), C.c_size_t(host.nbytes), C.c_int(2))
st = host.reshape(nblk, 32).astype(np.float64)
d = np.diff(st[:, :7], axis=1)
print("blocks", nblk, "cycles per phase of the stamped round: median / p10 / p90")
for i, n in enumerate(NAMES):
    print("  %-14s %8.0f %8.0f %8.0f" % (n, np.median(d[:, i]), np.percentile(d[:, i], 10), np.percentile(d[:, i], 90)))
tot = st[:, 6] - st[:, 0]
print("  %-14s %8.0f %8.0f %8.0f" % ("round", np.median(tot), np.percentile(tot, 10), np.percentile(tot, 90)))
