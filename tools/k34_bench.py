#!/usr/bin/env python
"""Micro-benchmarks of K3 (stitch accumulator, BASELINE cfg3) and the K4 statistics kernels on
one GPU: CUDA-event timing, algorithmic GB/s against the measured HBM peak.

    python tools/k34_bench.py [--reps 5] [--json gpurun_out/k34.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
import values_b200 as vb


def timed(fn, reps):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak, _ = bench.measured_peak_gbs()
    rows = []

    def report(name, ms, nbytes, extra=""):
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / peak})
        print(f"{name:58s} {ms:9.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of peak {extra}", flush=True)

    # ---------------- K3: cfg3 stitch, 256^3 volume, patch 64, overlap 0.5 -> 343 patches, C=2
    g = torch.Generator(device=dev).manual_seed(0)
    for vol, p, ov, N in (((256, 256, 256), 64, 0.5, 8), ((256, 256, 256), 64, 1.0, 8), ((250, 250, 200), 64, 0.5, 4)):
        crops = vb.patch_grid(vol, p, ov)
        patches = torch.rand((N, len(crops), 2, p, p, p), generator=g, device=dev, dtype=torch.float32)
        lo = vb.stitching.crops_to_lo(crops, dev)
        for out_dtype, es in ((torch.float64, 8), (torch.float32, 4)):
            out = torch.empty((N, 2) + vol, dtype=out_dtype, device=dev)
            cnt = torch.empty(vol, dtype=torch.float64, device=dev)
            nbytes = patches.numel() * 4 + out.numel() * es + cnt.numel() * 8
            ms = timed(lambda: vb.stitch_accumulate(patches, lo, out, cnt, accumulate=False), args.reps)
            report(f"K3 stitch vol={vol} p={p} ov={ov} N={N} {len(crops)} patches -> {str(out_dtype)[6:]}", ms, nbytes)
        w = vb.gaussian_importance_map((p, p, p), device=dev)
        ms = timed(lambda: vb.stitch_accumulate(patches, lo, out, cnt, accumulate=False, weight=w), args.reps)
        report(f"K3 stitch (gaussian weight map) vol={vol} ov={ov} -> float32", ms, nbytes)
        fac = vb.gaussian_importance_factors((p, p, p), device=dev)
        ms = timed(lambda: vb.stitch_accumulate(patches, lo, out, cnt, accumulate=False, weight=fac), args.reps)
        report(f"K3 stitch (gaussian weight, separable factors) vol={vol} ov={ov} -> float32", ms, nbytes)
        out64 = torch.empty((N, 2) + vol, dtype=torch.float64, device=dev)
        ms = timed(lambda: vb.stitch_accumulate(patches, lo, out64, cnt, accumulate=False, weight=fac), args.reps)
        report(f"K3 stitch (gaussian weight, separable factors) vol={vol} ov={ov} -> float64", ms,
               patches.numel() * 4 + out64.numel() * 8 + cnt.numel() * 8)
        del out64
        del patches, out, cnt

    # ---------------- K4: statistics over 32 x 128^3 maps
    maps32 = [torch.rand((128, 128, 128), generator=g, device=dev) ** 2 for _ in range(32)]
    n = sum(m.numel() for m in maps32)
    ms = timed(lambda: vb.quantile(maps32, 0.98), args.reps)
    report("K4 quantile(0.98) fp32, 32 x 128^3 (3 digit passes, digit selection on the device, one readback)", ms, 4 * n * 4, "(4 sweeps)")
    maps64 = [m.double() for m in maps32[:16]]
    n64 = sum(m.numel() for m in maps64)
    ms = timed(lambda: vb.quantile(maps64, 0.98), args.reps)
    report("K4 quantile(0.98) fp64, 16 x 128^3 (6 digit passes + NaN check)", ms, 7 * n64 * 8, "(7 sweeps)")
    hist = torch.zeros(2048, dtype=torch.int64, device=dev)
    big = torch.cat([m.reshape(-1) for m in maps32])

    def one_hist():
        vb._lib.check(vb._lib.lib.values_radix_histogram(big.data_ptr(), vb._lib.F32, big.numel(), 0, 0, 11,
                                                         hist.data_ptr(), vb._lib.stream_ptr(dev)))
    ms = timed(one_hist, args.reps)
    report("K4 radix_histogram fp32 top digit (every element counts), 268 MB", ms, big.numel() * 4)
    seg = (torch.rand(big.numel(), generator=g, device=dev) < 0.1).to(torch.uint8)
    ms = timed(lambda: vb.count_nonzero(seg), args.reps)
    report("K4 count_nonzero u8, 67 M voxels", ms, seg.numel())
    a = torch.stack(maps32[:16])
    b = torch.stack(maps32[16:])
    ms = timed(lambda: vb.ncc_batched(a, b), args.reps)
    report("K4 ncc_batched fp32, 16 pairs of 128^3 (2 sweeps)", ms, 2 * (a.numel() + b.numel()) * 4)
    unc = maps32[0]
    pred = (torch.rand(unc.shape, generator=g, device=dev) < 0.3).to(torch.uint8)
    refs = torch.stack([pred ^ (torch.rand(unc.shape, generator=g, device=dev) < 0.1).to(torch.uint8) for _ in range(4)])
    ms = timed(lambda: vb.metrics.calib_bins_fused(unc, pred, refs, 4.0, -1.0), args.reps)
    report("K4 calib_bins_fused fp32 map + 4 raters u8, 128^3", ms, unc.numel() * (4 + 1 + 4))
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"hbm_peak_gbs": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
