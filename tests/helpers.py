"""Shared fixtures: build the oracle and the engine for the same workload and the same draws."""
import numpy as np

from oracle import oracle as O
from stateline_b200 import Engine, configs


def make_replay(S, T, d, n_rounds, seed):
    rng = np.random.default_rng(seed)
    C = S * T
    z = rng.standard_normal((n_rounds, C, d))
    u_mh = rng.random((n_rounds, C))
    u_swap = rng.random((n_rounds, C))
    x_init = rng.uniform(-1.0, 1.0, (C, d))
    return z, u_mh, u_swap, x_init


def oracle_for(w, schedule=O.SCHED_BATCHED, rng_mode=O.RNG_PHILOX, seed=1234, stack_offset=0, adapt_interval=0, initial=None):
    return O.Sampler(S=w["S"], T=w["T"], J=w["J"], swap_interval=w["swapInterval"], opt_accept=w["optAccept"],
                     opt_swap=w["optSwap"], mn=w["mn"], mx=w["mx"], target=w["oracle_target"],
                     target_params=w["oracle_params"], schedule=schedule, rng_mode=rng_mode, seed=seed,
                     stack_offset=stack_offset, adapt_interval=adapt_interval, initial=initial)


def engine_for(w, rng="philox", seed=1234, sink="host", stack_offset=0, **gpu):
    if sink == "host":
        gpu.setdefault("hostRingRounds", 1024)  # the tests drain at the end
    cfg = configs.engine_config(w, rng=rng, seed=seed, sink=sink, stackOffset=stack_offset, **gpu)
    return Engine(cfg, plugin=w["plugin"], payload=w["payload"])


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / den
    e[(a == b)] = 0.0
    return float(np.max(e)) if e.size else 0.0


def scaled_err(a, b, scale):
    """max |a-b| / max(|b|, scale): relative error that does not blow up on entries that are small
    only through cancellation (a coordinate near 0 inside a [-10,10] box, an energy near its minimum)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), scale)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / den
    e[(a == b)] = 0.0
    return float(np.max(e)) if e.size else 0.0
