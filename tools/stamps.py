"""Development: per-block phase stamps of pt_step_fast_kernel (plugin built with -DSLGPU_STAMPS)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (device memory access below goes through the runtime torch loads)
from stateline_b200 import Engine, configs

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = configs.workload("config3", S=S)
rt = C.CDLL("libcudart.so.12")
with Engine(configs.engine_config(w, sink="device", windowRates=True), plugin=w["plugin"], payload=w["payload"]) as e:
    e.init_chains()
    e.step_rounds(209)
    e.sync()
    e.step_rounds(10)  # one launch
    e.sync()
    # the flags pointer is not exported: find it through the logger buffer is not possible, so re-read via a C helper
    ptr = C.c_void_p()
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "stateline_b200", "libslgpu.so"))
    lib.slgpu_debug_round_flags.restype = C.c_void_p
    lib.slgpu_debug_round_flags.argtypes = [C.c_void_p]
    dev = lib.slgpu_debug_round_flags(e.h)
    nblk = S // 16
    host = np.zeros(nblk * 32, dtype=np.uint64)
    rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(dev), C.c_size_t(host.nbytes), C.c_int(2))
st = host.reshape(nblk, 32).astype(np.float64)
t0 = st[:, 0].min()
st = (st - t0) / 1e3  # us
print("blocks", nblk, "kernel span us", st[:, 31 if False else 4 + 3 * 9].max())
first = st[:, 0] < 5.0
for name, sel in (("first wave", first), ("second wave", ~first)):
    if not sel.any():
        continue
    b = st[sel]
    print(name, "n=%d" % sel.sum(), "start med %.1f" % np.median(b[:, 0]))
    print("  prologue       %.2f" % np.median(b[:, 1] - b[:, 0]))
    for it in range(10):
        prev = b[:, 1] if it == 0 else b[:, 4 + 3 * (it - 1)]
        print("  round %d: A %.2f  B %.2f  tail %.2f  total %.2f" % (
            it, np.median(b[:, 2 + 3 * it] - prev), np.median(b[:, 3 + 3 * it] - b[:, 2 + 3 * it]),
            np.median(b[:, 4 + 3 * it] - b[:, 3 + 3 * it]), np.median(b[:, 4 + 3 * it] - prev)))
