"""Multi-rank host logic on CPU: volume sharding and the score-table gather (the only
collective on the path) with world_size 2 and 3 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_everything():
    from values_b200.sharding import shard_range, shard_sizes

    for n in (0, 1, 7, 8, 10_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from values_b200.sharding import gather_scores, shard_range

    lo, hi = shard_range(n_items, rank, world)
    # row i of the global table is [i, 10 i, 100 i] (fp64, like the score table)
    local = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1) * torch.tensor([1.0, 10.0, 100.0], dtype=torch.float64)
    full = gather_scores(local, n_items)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    from values_b200.sharding import AsyncScoreGather

    g = AsyncScoreGather(torch.device("cpu"))        # no CUDA: the synchronous gather behind the same API
    again = g.submit(local, n_items)
    g.wait()
    assert torch.equal(again, full)
    try:
        gather_scores(local[:-1] if hi > lo else torch.zeros(1, 3, dtype=torch.float64), n_items)
        raised = False
    except ValueError:
        raised = True
    torch.save(torch.tensor(raised), os.path.join(out_dir, f"e{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 7), (2, 8), (3, 10)])
def test_gather_scores_gloo(tmp_path, world, n_items):
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    want = torch.arange(n_items, dtype=torch.float64).unsqueeze(1) * torch.tensor([1.0, 10.0, 100.0], dtype=torch.float64)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(got, want), f"rank {r}"
        assert bool(torch.load(os.path.join(tmp_path, f"e{r}.pt")))  # wrong shard size is rejected


def test_gather_scores_single_process():
    from values_b200.sharding import gather_scores

    t = torch.rand(5, 21, dtype=torch.float64)
    assert gather_scores(t, 5) is t
    with pytest.raises(ValueError):
        gather_scores(t, 6)
