"""Reader of the binary chain file written by sink "bin" (<outputPath>/chains-<stackOffset>.slgpu): the delta stream of
the cold-chain rows as the GPU emitted it (include/slgpu.h: slgpu_segment).  expand() rebuilds the rows that
CSVChainArrayWriter would have printed (src/db/db.cpp:18-25): x_0..x_{d-1}, energy, sigma, beta, accepted, swapType."""
import struct

import numpy as np

FILE_HEADER = struct.Struct("<8s6I")   # magic, version, n_stacks, n_temps, n_dims, stack_offset, chunk
RECORD_HEADER = struct.Struct("<4sIQQQ")  # magic, n_rounds, row_begin, bytes, reserved


def _layout(n_rounds, S, chunk):
    n_chunks = (S + chunk - 1) // chunk
    off_flags = (16 + n_rounds * n_chunks * 4 + 15) // 16 * 16
    off_payload = (off_flags + n_rounds * S + 15) // 16 * 16
    return n_chunks, off_flags, off_payload


def expand_segment(prev, n_rounds, base, flags, payload, chunk=256):
    """Rows [n_rounds, S, d + 5] of one segment; prev [S, d + 3] is updated in place to the state after it."""
    S, wp = prev.shape
    rows = np.empty((n_rounds, S, wp + 2))
    for r in range(n_rounds):
        ch = (flags[r] & 8) != 0
        if ch.any():
            # slot of a changed row = base of its chunk + its rank among the chunk's changed rows
            rank = np.cumsum(ch) - 1
            chunk_of = np.arange(S) // chunk
            first_rank = np.concatenate([[0], np.cumsum(np.add.reduceat(ch.astype(np.int64), np.arange(0, S, chunk)))[:-1]])
            slot = base[r][chunk_of].astype(np.int64) + (rank - first_rank[chunk_of])
            prev[ch] = payload[slot[ch]]
        rows[r, :, :wp] = prev
        rows[r, :, wp] = flags[r] & 1
        rows[r, :, wp + 1] = (flags[r] >> 1) & 3
    return rows


def read(path):
    """-> (header dict, rows [n_rows_total / S, S, d + 5]) of a whole file."""
    with open(path, "rb") as f:
        magic, version, S, T, d, stack_offset, chunk = FILE_HEADER.unpack(f.read(FILE_HEADER.size))
        if magic != b"SLGPUBIN" or version != 1:
            raise ValueError("%s is not a slgpu binary chain file" % path)
        wp = d + 3
        prev = np.zeros((S, wp))
        out = []
        while True:
            h = f.read(RECORD_HEADER.size)
            if not h:
                break
            m, n_rounds, row_begin, nbytes, _ = RECORD_HEADER.unpack(h)
            if m != b"SLGS":
                raise ValueError("corrupt record header in %s" % path)
            seg = f.read(nbytes)
            n_chunks, off_flags, off_payload = _layout(n_rounds, S, chunk)
            count = struct.unpack_from("<I", seg, 8)[0]
            base = np.frombuffer(seg, dtype=np.uint32, count=n_rounds * n_chunks, offset=16).reshape(n_rounds, n_chunks)
            flags = np.frombuffer(seg, dtype=np.uint8, count=n_rounds * S, offset=off_flags).reshape(n_rounds, S)
            payload = np.frombuffer(seg, dtype=np.float64, count=count * wp, offset=off_payload).reshape(count, wp)
            out.append(expand_segment(prev, n_rounds, base, flags, payload, chunk))
    hdr = dict(n_stacks=S, n_temps=T, n_dims=d, stack_offset=stack_offset)
    return hdr, (np.concatenate(out) if out else np.empty((0, S, wp + 2)))
