"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partition, sync cadence,
NCCL-id exchange plumbing and the reduction of job statistics."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from word2bits_b200.parallel import DataParallel, exchange_unique_id, shard_range


class FakeTrainer:
    def __init__(self, rank):
        self.rank, self.steps, self.syncs = rank, 0, 0

    def train_step(self, words):
        self.steps += 1
        return {"words": words * (self.rank + 1), "positions": 10 * (self.rank + 1), "kernel_ms": 5.0 + self.rank}

    def sync(self):
        dist.barrier()
        self.syncs += 1
        return 1.5  # device ms of the exchange, as Trainer.sync reports it

    def table_checksum(self):
        return (1 << 63) + 12345, 777  # equal on every rank after a sync (top bit: must survive the int64 transport)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(rank, world, 296 + 1)
        uid = exchange_unique_id(dist, lambda: bytes(range(128)))
        t = FakeTrainer(rank)
        dp = DataParallel(t, dist, sync_every=4)
        words = 0
        for _ in range(10):
            words += dp.step(1000)["words"]
        dp.finish()
        sums, maxes = dp.reduce(sums=[words], maxes=[5.0 + rank])
        assert dp.replicas_identical() and dp.sync_ms == 1.5 * t.syncs
        q.put((rank, lo, hi, uid == bytes(range(128)), t.steps, t.syncs, sums[0], maxes[0]))
    finally:
        dist.destroy_process_group()


def test_two_rank_orchestration():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ok0, st0, sy0, sum0, max0), (r1, lo1, hi1, ok1, st1, sy1, sum1, max1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 148, 148, 297)       # contiguous, remainder to the last rank
    assert ok0 and ok1                                        # id created on rank 0 reaches rank 1
    assert st0 == st1 == 10 and sy0 == sy1 == 3               # steps 4, 8 and the final flush
    assert sum0 == sum1 == 10 * 1000 * (1 + 2)                # whole-job words
    assert max0 == max1 == 6.0                                # max over ranks


def test_shard_range_covers_everything():
    for world in (1, 2, 4, 8):
        for total in (8, 148, 1184, 1187):
            got = [shard_range(r, world, total) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
