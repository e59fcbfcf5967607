#!/usr/bin/env python
"""Throughput of the path's entry points on shapes / sample counts next to the BASELINE ones (which kernel a
call lands on is decided by divisibility and dtype): algorithmic GB/s against the measured HBM peak.

    python tools/cliff_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import values_b200 as vb

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak_gbs()
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps=4):
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


def line(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:70s} {ms:8.3f} ms {gbs:7.1f} GB/s = {gbs / peak:.2f} of peak", flush=True)


for N, C, sp, dt, B in ((3, 4, (128, 128, 128), torch.float32, 16), (7, 4, (128, 128, 128), torch.float32, 12),
                        (17, 4, (128, 128, 128), torch.float32, 6), (6, 150, (512, 512), torch.float32, 2),
                        (3, 2, (128, 128, 128), torch.float64, 16), (7, 2, (128, 128, 128), torch.float64, 8),
                        (12, 2, (128, 128, 128), torch.float64, 6), (10, 20, (1024, 2048), torch.bfloat16, 2)):
    x = torch.softmax(torch.randn((B, N, C) + sp, generator=g, device=dev), dim=2).to(dt)
    V = x[0, 0, 0].numel()
    ms = timed(lambda: vb.uncertainty_fused(x, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05)))
    line(f"K1 N={N} C={C} {sp} {str(dt)[6:]} maps + arg-max + scores", ms, B * V * (N * C * x.element_size() + 13))
    del x
for C, sp, dt in ((4, (256, 256, 256), torch.float32), (2, (256, 256, 256), torch.float64), (20, (2048, 2048), torch.float32)):
    x = torch.softmax(torch.randn((C,) + sp, generator=g, device=dev), dim=0).to(dt)
    ms = timed(lambda: vb.calculate_one_minus_msr(x))
    line(f"a2 one_minus_msr C={C} {sp} {str(dt)[6:]}", ms, x.numel() * x.element_size() + x[0].numel() * x.element_size())
    del x
maps = torch.rand((96, 128, 128, 128), generator=g, device=dev)
ms = timed(lambda: [vb.image_level_aggregation(maps[i]) for i in range(4)])
line("a8 image_level_aggregation, 4 calls on 128^3 fp32 maps (host dicts)", ms, 4 * maps[0].numel() * 4)
ms = timed(lambda: [vb.threshold_aggregation(maps[i], threshold=0.7) for i in range(4)])
line("a9 threshold_aggregation, 4 calls on 128^3 fp32 maps (host dicts)", ms, 4 * maps[0].numel() * 4)
ms = timed(lambda: [vb.patch_level_aggregation(maps[i], 10) for i in range(4)])
line("a7 patch_level_aggregation, 4 calls on 128^3 fp32 maps (host dicts)", ms, 4 * maps[0].numel() * 4)
m64 = maps[:24].double()
ms = timed(lambda: vb.patch_max(m64, 10))
line("a7 patch_max 24 fp64 maps of 128^3 (the NIfTI dtype)", ms, m64.numel() * 8)
ms = timed(lambda: vb.patch_max(maps[:48], 8))
line("a7 patch_max 48 fp32 maps of 128^3, patch 8", ms, 48 * maps[0].numel() * 4)
