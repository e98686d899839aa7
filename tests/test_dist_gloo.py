"""world_size-2 gloo test of the multi-GPU host logic (stateline_b200/dist.py): sharding and the
two-stage all-reduced R-hat equal the single-process result over the union of the shards."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from stateline_b200 import dist as sd


def test_shard_stacks_partitions_contiguously():
    for total, world in ((8192, 8), (10, 4), (7, 2), (3, 8)):
        spans = [sd.shard_stacks(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1


class OracleShard:
    """The interface dist.allreduce_rhat needs, fed from an oracle run of this rank's stacks."""

    def __init__(self, rank, world, total_stacks, rounds):
        from oracle import oracle as O
        off, n = sd.shard_stacks(total_stacks, rank, world)
        self.T = 4
        self.o = O.Sampler(S=n, T=self.T, J=3, swap_interval=10, opt_accept=0.234, opt_swap=0.3874, mn=[-10, 0, -10, -10],
                           mx=[10, 10, 10, 2], target=O.TARGET_DEMO_GAUSS, schedule=O.SCHED_BATCHED, seed=5, stack_offset=off)
        self.o.init_chains()
        self.o.run_rounds(rounds)
        self.n, self.rounds = n, rounds
        cold = np.stack([self.o.cold_rows(s)[1:, :4] for s in range(n)])  # [stack][round][d], initial row excluded
        self.M = cold.mean(axis=1)
        self.S = ((cold - self.M[:, None, :]) ** 2).sum(axis=1)

    def rhat_stage1(self):
        return self.M.sum(axis=0), float(self.n), float(self.rounds)

    def rhat_stage2(self, mean):
        return ((self.M - mean) ** 2).sum(axis=0), self.S.sum(axis=0)

    def counters(self):
        return self.o.counters()


def _worker(rank, world, port, total, rounds, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = OracleShard(rank, world, total, rounds)
    rhat = sd.allreduce_rhat(shard)
    acc, swp = sd.allreduce_rates(shard)
    if rank == 0:
        np.save(out, np.concatenate([rhat, acc, swp]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_rhat_matches_union(tmp_path):
    total, rounds, world = 6, 400, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, port, total, rounds, out), nprocs=world, join=True)
    got = np.load(out)
    shards = [OracleShard(r, world, total, rounds) for r in range(world)]
    M = np.concatenate([s.M for s in shards])
    S = np.concatenate([s.S for s in shards])
    m, n = float(total), float(rounds)
    B = n / (m - 1.0) * ((M - M.mean(axis=0)) ** 2).sum(axis=0)   # diagnostics.cpp:58-70
    W = (1.0 / m) * (S / (n - 1.0)).sum(axis=0)
    want = np.sqrt(((n - 1.0) / n * W + B / n) / (W + 1e-30))
    assert np.allclose(got[:4], want, rtol=1e-12)
    acc = sum(s.counters()["accepts"].reshape(-1, 4).sum(0) for s in shards) / sum(s.counters()["length"].reshape(-1, 4).sum(0) for s in shards)
    assert np.allclose(got[4:8], acc, rtol=1e-12)


def test_cpulist_parse_and_bind_is_harmless_without_gpu():
    import os
    from stateline_b200.dist import _parse_cpulist, bind_to_gpu_cpus
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("\n") == set()
    before = os.sched_getaffinity(0)
    got = bind_to_gpu_cpus(0)
    assert got is None or got <= before
    if got is None:
        assert os.sched_getaffinity(0) == before
