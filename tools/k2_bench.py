#!/usr/bin/env python
"""Micro-benchmark of K2b (values_patch_max) on the GPU box: M fp32 maps resident in L2/HBM,
CUDA-event timing per implementation path.

    python tools/k2_bench.py [--shape 128,128,128] [--maps 3,12,24] [--paths 0,5] [--patch 10]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb
from values_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="128,128,128")
    ap.add_argument("--maps", default="3,12")
    ap.add_argument("--paths", default="0,5")
    ap.add_argument("--patch", type=int, default=10)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sparse", action="store_true", help="95 %% zeros (real maps are mostly background)")
    args = ap.parse_args()
    shape = tuple(int(v) for v in args.shape.split(","))
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    for M in [int(v) for v in args.maps.split(",")]:
        maps = torch.rand((M,) + shape, generator=g, device=dev)
        if args.sparse:
            maps = torch.where(maps > 0.95, maps, torch.zeros_like(maps))
        ref = None
        for path in [int(v) for v in args.paths.split(",")]:
            score, bbox = vb.patch_max(maps, args.patch, path=path)
            torch.cuda.synchronize()
            if ref is None:
                ref = (score.clone(), bbox.clone())
            rel = ((ref[0] - score).abs() / ref[0].abs().clamp_min(1e-300)).max().item()
            same = f"{torch.equal(ref[1], bbox)} (score rel diff {rel:.1e})"
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                vb.patch_max(maps, args.patch, path=path)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            vox = M * maps[0].numel()
            print(f"shape={shape} M={M} path={path}: {best * 1e3:8.1f} us  {best * 1e3 / M:7.2f} us/map "
                  f"{vox / best / 1e6:7.2f} Gvox/s  ({vox * 4 / best / 1e6:7.1f} GB/s of map bytes) "
                  f"bbox_same_as_first={same}", flush=True)


if __name__ == "__main__":
    main()
