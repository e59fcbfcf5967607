#!/usr/bin/env python
"""K2b / K3 on shapes whose rows are not 16-byte aligned (odd innermost extents) next to the aligned shapes:
which kernel runs and what it costs.  K1's counterpart is tools/k1_bench.py --shape odd*.

    python tools/odd_shape_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import values_b200 as vb


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak_gbs()
    g = torch.Generator(device=dev).manual_seed(0)
    for shape, m in (((128, 128, 128), 96), ((127, 127, 127), 96), ((128, 128, 126), 96), ((1024, 2048), 18), ((1023, 2047), 18)):
        maps = torch.rand((m,) + shape, generator=g, device=dev)
        ms = timed(lambda: vb.patch_max(maps, 10))
        gbs = maps.numel() * 4 / ms / 1e6
        print(f"K2b patch_max {m} maps of {shape}: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s = {gbs / peak:.2f} of peak", flush=True)
        del maps
    for vol, p in (((256, 256, 256), 64), ((255, 255, 255), 64), ((256, 256, 254), 64)):
        crops = vb.patch_grid(vol, p, 0.5)
        patches = torch.rand((8, len(crops), 2, p, p, p), generator=g, device=dev)
        lo = vb.stitching.crops_to_lo(crops, dev)
        out = torch.empty((8, 2) + vol, dtype=torch.float64, device=dev)
        cnt = torch.empty(vol, dtype=torch.float64, device=dev)
        nbytes = patches.numel() * 4 + out.numel() * 8 + cnt.numel() * 8
        ms = timed(lambda: vb.stitch_accumulate(patches, lo, out, cnt, accumulate=False))
        gbs = nbytes / ms / 1e6
        print(f"K3 stitch {len(crops)} patches of {p}^3 into {vol} fp64: {ms:8.3f} ms  {gbs:7.1f} GB/s = {gbs / peak:.2f} of peak", flush=True)
        del patches, out, cnt


if __name__ == "__main__":
    main()
