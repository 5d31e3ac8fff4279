"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): replica averaging through
libw2b's NCCL path (w2b_sync) must produce the arithmetic mean of the replicas on every rank and
the exact global word count."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import word2bits_b200 as w2b  # noqa: E402
from word2bits_b200.parallel import DataParallel, exchange_unique_id, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    path = sys.argv[1]
    c = w2b.Corpus(path, 5)
    S = 8 * world
    lo, hi = shard_range(rank, world, S)
    t = w2b.Trainer(c, size=64, window=5, negative=6, bitlevel=1, threads=S, shard_range=(lo, hi), iter=1, device=local)
    t.nccl_init(exchange_unique_id(dist, w2b.nccl_unique_id, device="cuda"), rank, world)
    dp = DataParallel(t, dist, sync_every=1000, device="cuda")
    words = 0
    for _ in range(3):
        words += dp.step(3000)["words"]
    u, v = t.download_raw()
    gu = [torch.empty_like(torch.from_numpy(u)).cuda() for _ in range(world)]
    gv = [torch.empty_like(torch.from_numpy(v)).cuda() for _ in range(world)]
    dist.all_gather(gu, torch.from_numpy(u).cuda())
    dist.all_gather(gv, torch.from_numpy(v).cuda())
    mean_u = torch.stack(gu).double().mean(0).float().cpu().numpy()
    mean_v = torch.stack(gv).double().mean(0).float().cpu().numpy()
    assert not np.array_equal(gu[0].cpu().numpy(), gu[-1].cpu().numpy())  # replicas really diverged
    t.sync()
    u2, v2 = t.download_raw()
    assert np.allclose(u2, mean_u, rtol=0, atol=2e-7) and np.allclose(v2, mean_v, rtol=0, atol=2e-7)
    ref = torch.from_numpy(u2).cuda()
    dist.broadcast(ref, 0)
    assert np.array_equal(ref.cpu().numpy(), u2)  # bit-identical replicas after the average
    assert dp.replicas_identical()                # ... which is what the bench's sync_check reports
    (tot,), _ = dp.reduce(sums=[words])
    _, wca = t.get_state()
    # every shard is still below the 10k-word cadence except those that crossed it: the synced
    # counter must equal the sum over ranks of what each rank had added (scaled back)
    assert wca <= tot and wca % 1 == 0
    for _ in range(2):
        dp.step(3000)
    dp.finish()
    t.close()
    # ---- sync_mode 1: every rank's updates since the last exchange are summed onto the common base
    t = w2b.Trainer(c, size=64, window=5, negative=6, bitlevel=1, threads=S, shard_range=(lo, hi), iter=1, device=local,
                    sync_mode=1)
    t.nccl_init(exchange_unique_id(dist, w2b.nccl_unique_id, device="cuda"), rank, world)
    u0, v0 = t.download_raw()
    dp = DataParallel(t, dist, sync_every=1000, device="cuda")
    for _ in range(3):
        dp.step(3000)
    u, v = t.download_raw()
    gu = [torch.empty_like(torch.from_numpy(u)).cuda() for _ in range(world)]
    dist.all_gather(gu, torch.from_numpy(u).cuda())
    want_u = (torch.from_numpy(u0).cuda().double() + sum(g.double() - torch.from_numpy(u0).cuda().double() for g in gu)).float().cpu().numpy()
    t.sync()
    u2, v2 = t.download_raw()
    d_want, d_local = float((np.abs(u2 - want_u) / (1.0 + np.abs(want_u))).max()), float(np.abs(u2 - u).max())
    assert d_want <= 1e-6, ("sum mode: result differs from base + sum of deltas (relative)", d_want, d_local)
    assert d_local > 1e-4, ("sum mode: the other rank's updates did not arrive", d_want, d_local)
    assert dp.replicas_identical()
    for _ in range(2):
        dp.step(3000)
    t.sync()
    assert dp.replicas_identical()
    if rank == 0:
        print("MGPU_OK world=%d words=%d wca=%d" % (world, int(tot), wca))
    t.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
