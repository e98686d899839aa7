"""binlog.expand_segment (the host-side reader of the delta stream) on a synthetic segment: slots of a chunk follow the stack
order, chunks may come in any order (the device hands out bases with an atomic add)."""
import numpy as np

from stateline_b200 import binlog


def test_expand_segment_synthetic():
    rng = np.random.default_rng(0)
    S, wp, R, chunk = 700, 5, 6, 256
    n_chunks = (S + chunk - 1) // chunk
    changed = {r: (rng.random(S) < 0.3 if r else np.ones(S, bool)) for r in range(R)}
    order = [(r, c) for r in range(R) for c in range(n_chunks)]
    rng.shuffle(order)
    base = np.zeros((R, n_chunks), np.uint32)
    slot = 0
    for r, c in order:
        base[r, c] = slot
        slot += int(changed[r][c * chunk:(c + 1) * chunk].sum())
    payload = np.zeros((slot, wp))
    flags = np.zeros((R, S), np.uint8)
    rows = np.zeros((R, S, wp + 2))
    state = np.zeros((S, wp))
    for r in range(R):
        vals = rng.standard_normal((S, wp))
        for c in range(n_chunks):
            ch = changed[r][c * chunk:(c + 1) * chunk]
            payload[base[r, c]:base[r, c] + ch.sum()] = vals[c * chunk:(c + 1) * chunk][ch]
        state[changed[r]] = vals[changed[r]]
        acc, sw = rng.integers(0, 2, S), rng.integers(0, 3, S)
        flags[r] = acc | (sw << 1) | (changed[r] * 8)
        rows[r, :, :wp], rows[r, :, wp], rows[r, :, wp + 1] = state, acc, sw
    prev = np.zeros((S, wp))
    assert np.array_equal(binlog.expand_segment(prev, R, base, flags, payload, chunk), rows)
    assert np.array_equal(prev, state)
