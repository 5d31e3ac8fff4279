"""Data-parallel orchestration of the training path across GPUs.

One process per GPU (torchrun).  The corpus is cut into S = G * shards_per_gpu shards exactly as
the reference cuts it across threads (src/word2bits.cpp:377,414); rank g trains shards
[g*S/G, (g+1)*S/G) on a full replica of u/v; every `sync_every` steps the replicas are
all-reduce-averaged (NCCL inside libw2b: w2b_sync) and the global word counter that drives the
learning rate is made exact.  torch.distributed is only the plumbing: rendezvous, the broadcast of
the NCCL unique id, barriers and the reduction of the reported statistics."""
import numpy as np


def shard_range(rank, world, total_shards):
    """Contiguous block of shards owned by `rank` (the last rank takes the remainder)."""
    per = total_shards // world
    lo = rank * per
    hi = total_shards if rank == world - 1 else lo + per
    return lo, hi


def exchange_unique_id(dist, make_id, device=None):
    """Rank 0 creates the 128-byte NCCL id, everyone receives it through the default group."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


class DataParallel:
    """Drives one Trainer per rank: train_step everywhere, replica averaging every k steps,
    whole-job statistics (sum of words/positions, max of times) on demand."""

    def __init__(self, trainer, dist=None, sync_every=4, device=None):
        self.t, self.dist, self.k, self.device = trainer, dist, max(1, int(sync_every)), device
        self.world = dist.get_world_size() if dist else 1
        self.steps = 0
        self.syncs = 0
        self.sync_ms = 0.0  # device time spent in the replica average (CUDA events inside libw2b)

    def step(self, words_per_shard):
        st = self.t.train_step(words_per_shard)
        self.steps += 1
        if self.world > 1 and self.steps % self.k == 0:
            self.sync_ms += self.t.sync()
            self.syncs += 1
        return st

    def finish(self):
        """Replicas must agree before anything is exported."""
        if self.world > 1 and self.steps % self.k != 0:
            self.sync_ms += self.t.sync()
            self.syncs += 1

    def replicas_identical(self):
        """After a sync every rank holds the same bits: compare table fingerprints across ranks."""
        if not self.dist or self.world == 1:
            return True
        import torch
        a, b = self.t.table_checksum()
        mine = torch.tensor([a & 0x7FFFFFFFFFFFFFFF, b & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=self.device)
        allv = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allv, mine)
        return all(bool((x == allv[0]).all()) for x in allv)

    def reduce(self, sums=(), maxes=()):
        """All-reduce python floats: returns (summed list, maxed list)."""
        if not self.dist or self.world == 1:
            return list(sums), list(maxes)
        import torch
        s = torch.tensor(list(sums) or [0.0], dtype=torch.float64, device=self.device)
        m = torch.tensor(list(maxes) or [0.0], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(s, op=self.dist.ReduceOp.SUM)
        self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
        return s.tolist()[: len(sums)], m.tolist()[: len(maxes)]
