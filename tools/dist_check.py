#!/usr/bin/env python
"""Multi-GPU check of the only two cross-rank steps of the path (run under torchrun, NCCL):
  * gather_scores: the per-image score table all_gather;
  * quantile(distributed=True): the global threshold over the maps of every rank (section 8f1),
    checked on every rank against np.quantile of the gathered maps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import values_b200 as vb


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    for dtype in (torch.float32, torch.float64):
        maps = [(torch.rand((48, 50, 52), generator=g, device=dev, dtype=torch.float64) ** (2 + rank)).to(dtype)
                for _ in range(2 + rank)]       # ranks hold different numbers of maps
        maps[0][:3] = 0.0
        flat = torch.cat([m.reshape(-1) for m in maps])
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([flat.numel()], device=dev))
        pad = int(max(s.item() for s in sizes))
        buf = torch.zeros(pad, dtype=dtype, device=dev)
        buf[: flat.numel()] = flat
        allbuf = [torch.empty(pad, dtype=dtype, device=dev) for _ in range(world)]
        dist.all_gather(allbuf, buf)
        everything = np.concatenate([allbuf[r][: int(sizes[r].item())].cpu().numpy() for r in range(world)])
        for q in (0.0, 0.5, 0.98, 0.999, 1.0):
            got = vb.quantile(maps, q, distributed=True)
            want = np.quantile(everything, q)
            if not (got.dtype == want.dtype and got == want):
                ok = False
                print(f"rank {rank}: quantile mismatch dtype={dtype} q={q}: {got!r} vs {want!r}", flush=True)
    # score table gather: uneven shards
    n_items = 5 * world + 1
    lo, hi = vb.shard_range(n_items, rank, world)
    local_tab = torch.arange(lo, hi, device=dev, dtype=torch.float64).reshape(-1, 1).repeat(1, 21)
    full = vb.gather_scores(local_tab, n_items)
    if not torch.equal(full[:, 0], torch.arange(n_items, device=dev, dtype=torch.float64)):
        ok = False
        print(f"rank {rank}: gather_scores mismatch", flush=True)
    # the same gather on a side stream, several in flight, while the main stream keeps computing
    gat = vb.AsyncScoreGather(dev)
    outs = []
    for k in range(6):
        tab = local_tab + 1000.0 * k
        outs.append(gat.submit(tab, n_items))
        torch.rand(1 << 24, device=dev).sum()          # main-stream work the gathers overlap with
    gat.wait()
    for k, out in enumerate(outs):
        if not torch.equal(out[:, 0], torch.arange(n_items, device=dev, dtype=torch.float64) + 1000.0 * k):
            ok = False
            print(f"rank {rank}: AsyncScoreGather mismatch at {k}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_check", "OK" if flag.item() == 1 else "FAILED", f"world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
