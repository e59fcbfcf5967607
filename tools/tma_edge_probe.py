#!/usr/bin/env python
"""Do the tensor-map copies of K3 / K2b / the bulk copies of K1 ever read past the END (or, K3, before the START) of their input?
Run with PYTORCH_NO_CUDA_MEMORY_CACHING=1: every input is placed so that it ends exactly at the end of its own
cudaMalloc'd 2 MB region, with the neighbouring regions freed (unmapped), and the kernels are run on it; an
over-read is then a hardware fault (illegal address), not a silent read of a neighbour's bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb

assert os.environ.get("PYTORCH_NO_CUDA_MEMORY_CACHING") == "1", "run with PYTORCH_NO_CUDA_MEMORY_CACHING=1"
dev = torch.device("cuda")
PAGE = 2 << 20


def at_region_end(nbytes, dtype):
    """A tensor of nbytes that ends exactly at the end of a freshly malloc'd region whose neighbours are free."""
    n_pages = (nbytes + PAGE - 1) // PAGE
    bufs = [torch.empty(n_pages * PAGE, dtype=torch.uint8, device=dev) for _ in range(5)]
    keep = bufs[2]
    ptrs = sorted(b.data_ptr() for b in bufs)
    del bufs
    view = keep[keep.numel() - nbytes:].view(dtype)
    return keep, view, ptrs


g = torch.Generator(device=dev).manual_seed(0)
for trial in range(6):
    # K3: patches end at the region end
    shape, p, ov = (24, 20, 16), 8, 0.5
    crops = vb.patch_grid(shape, p, ov)
    n = 3 * len(crops) * 2 * p * p * p
    keep, flat, ptrs = at_region_end(n * 4, torch.float32)
    flat.copy_(torch.rand(n, generator=g, device=dev))
    patches = flat.view(3, len(crops), 2, p, p, p)
    a = vb.stitch_volume(patches, crops, shape, path=0)
    torch.cuda.synchronize()
    b = vb.stitch_volume(patches.clone(), crops, shape, path=1)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # K3 again with the patches at the START of a region (windows hanging over a patch's low faces have negative
    # tensor coordinates: an under-read would hit the freed region in front)
    n_pages = (n * 4 + PAGE - 1) // PAGE
    bufs = [torch.empty(n_pages * PAGE, dtype=torch.uint8, device=dev) for _ in range(5)]
    keep0 = bufs[2]
    del bufs
    front = keep0[:n * 4].view(torch.float32)
    front.copy_(flat)
    c = vb.stitch_volume(front.view(3, len(crops), 2, p, p, p), crops, shape, path=0)
    torch.cuda.synchronize()
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    # K1 ring: the stack ends at the region end (ragged last tile: V not a multiple of the tile)
    for dtype, N, C, V in ((torch.float32, 8, 3, 4 * 333), (torch.float64, 5, 2, 2 * 257), (torch.bfloat16, 4, 3, 8 * 77)):
        es = torch.empty((), dtype=dtype).element_size()
        keep1, flat1, _ = at_region_end(2 * N * C * V * es, dtype)
        flat1.copy_(torch.softmax(torch.randn(2, N, C, V, generator=g, device=dev), dim=2).to(dtype).reshape(-1))
        x = flat1.view(2, N, C, V)
        r = vb.uncertainty_fused(x, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.1))
        torch.cuda.synchronize()
        r2 = vb.uncertainty_fused(x.clone(), mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.1))
        torch.cuda.synchronize()
        assert torch.equal(r.pred_entropy, r2.pred_entropy) and torch.equal(r.scores, r2.scores)
    # K2b strip filter: the maps end at the region end
    M, s3 = 3, (24, 41, 52)
    keep2, flat2, _ = at_region_end(M * s3[0] * s3[1] * s3[2] * 4, torch.float32)
    flat2.copy_(torch.rand(flat2.numel(), generator=g, device=dev))
    maps = flat2.view(M, *s3)
    for m in range(M):
        d0 = vb.patch_level_aggregation(maps[m], 10)
        d1 = vb.patch_level_aggregation(maps[m].clone(), 10)
        assert d0 == d1, (d0, d1)
    torch.cuda.synchronize()
    print("trial", trial, "ok: inputs ended at", hex(keep.data_ptr() + keep.numel()), flush=True)
print("tma_edge_probe: no over-read")
