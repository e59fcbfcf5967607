#!/usr/bin/env python
"""K1 calls with arg-max outputs (mean / per-sample, with and without the maps) on the BASELINE shapes:
milliseconds and algorithmic GB/s against the measured HBM peak.

    python tools/argmax_probe.py
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import values_b200 as vb
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak_gbs()
g = torch.Generator(device=dev).manual_seed(0)
def timed(fn, reps=4):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for N, C, sp, dt, B in ((16, 4, (128, 128, 128), torch.float32, 8), (8, 2, (256, 256, 256), torch.float64, 1), (5, 2, (64, 64, 64), torch.float64, 64), (10, 20, (1024, 2048), torch.float32, 2)):
    x = torch.softmax(torch.randn((B, N, C) + sp, generator=g, device=dev), dim=2).to(dt)
    V = x[0, 0, 0].numel(); es = x.element_size()
    for kw, ob in ((dict(mean_argmax=True), 13), (dict(mean_argmax=True, sample_argmax=True), 13 + N), (dict(maps=False, sample_argmax=True), N), (dict(maps=False, mean_argmax=True), 1)):
        ms = timed(lambda: vb.uncertainty_fused(x, **kw))
        gbs = B * V * (N * C * es + ob) / ms / 1e6
        print(f"N={N} C={C} {sp} {str(dt)[6:]} {kw}: {ms:8.3f} ms {gbs:7.1f} GB/s = {gbs / peak:.2f} of peak", flush=True)
    del x
