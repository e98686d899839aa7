"""Development: per-warp clock64 stamps of one round of pt_step_warp_kernel (plugin built with -DSLGPU_WSTAMPS=<round>).
usage: python tools/wstamps.py <stamped plugin .so> [config4|config5] [stacks]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SLGPU_TARGETS_SO"] = os.path.abspath(sys.argv[1])
import torch  # noqa: F401  (loads the CUDA runtime the copy below uses)
from stateline_b200 import Engine, configs

w = configs.workload(sys.argv[2] if len(sys.argv) > 2 else "config4", S=int(sys.argv[3]) if len(sys.argv) > 3 else None)
rt = C.CDLL("libcudart.so.12")
NAMES = ["normals", "gemv", "bounce+likelihood", "MH+sigma", "shaper", "swap sweep"]
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
    nC = w["S"] * w["T"]
    host = np.zeros(nC + 16, dtype=np.int64)
    rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(dev), C.c_size_t(host.nbytes), C.c_int(2))
idx = np.arange(0, nC, 16)
st = np.stack([host[idx + k] for k in range(8)], axis=1).astype(np.float64)
acc = st[:, 7] > 0
d = np.diff(st[:, :7], axis=1)
print("stamped warps", len(idx), "accepted", int(acc.sum()), "; cycles per phase: median / p10 / p90  | median when accepted")
for i, n in enumerate(NAMES):
    print("  %-18s %8.0f %8.0f %8.0f | %8.0f" % (n, np.median(d[:, i]), np.percentile(d[:, i], 10), np.percentile(d[:, i], 90),
                                               np.median(d[acc, i]) if acc.any() else 0))
tot = st[:, 6] - st[:, 0]
print("  %-18s %8.0f %8.0f %8.0f | %8.0f" % ("round", np.median(tot), np.percentile(tot, 10), np.percentile(tot, 90), np.median(tot[acc]) if acc.any() else 0))
