#!/usr/bin/env python
"""Micro-benchmark of K1 (values_uncertainty_fused) on the GPU box: CUDA-event timing of the
kernel alone over a pool larger than L2, for kernel variants / tiles-per-CTA settings.

    python tools/k1_bench.py [--shape cfg5|cfg4|cfg2|cfg4bf16] [--variants 0,1,2] [--iters 0,1,2,4] [--batch 4]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb
from values_b200 import _lib

SHAPES = {
    "cfg5": (16, 4, (128, 128, 128), torch.float32, 24),
    "cfg4": (10, 20, (1024, 2048), torch.float32, 4),
    "cfg4bf16": (10, 20, (1024, 2048), torch.bfloat16, 4),
    "cfg2": (5, 2, (64, 64, 64), torch.float32, 512),
    "cfg3": (16, 2, (256, 256, 256), torch.float64, 2),
    "cfg5f64": (16, 4, (128, 128, 128), torch.float64, 8),
    "cfg1": (5, 2, (64, 64, 64), torch.float64, 160),
    "cfg3n8": (8, 2, (256, 256, 256), torch.float64, 2),
    # volumes whose voxel count is not a multiple of the 16-byte vector: rows start at every phase
    "odd32": (16, 4, (127, 127, 127), torch.float32, 24),
    "odd64": (8, 2, (255, 255, 255), torch.float64, 2),
    "oddbf16": (10, 20, (1023, 2047), torch.bfloat16, 4),
    "odd2": (5, 2, (63, 63, 63), torch.float32, 512),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="cfg5")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--iters", default="0")
    ap.add_argument("--batch", type=int, default=0, help="volumes per launch (0 = whole pool)")
    ap.add_argument("--scores", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    N, C, spatial, dtype, pool = SHAPES[args.shape]
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.empty((pool, N, C) + spatial, dtype=dtype, device=dev)
    for i in range(pool):
        x[i] = torch.softmax(torch.randn((N, C) + spatial, generator=g, device=dev) * 3, dim=1).to(dtype)
    V = x[0, 0, 0].numel()
    es = x.element_size()
    bpv = N * C * es + 13
    batch = args.batch or pool
    out = torch.empty((3, batch) + spatial, dtype=torch.float32, device=dev)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    ref = vb.uncertainty_fused(x[:batch], mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05))
    ref = (ref.pred_entropy.clone(), ref.expected_entropy.clone(), ref.mutual_information.clone(),
           ref.mean_argmax.clone(), ref.scores.clone())
    for variant in [int(v) for v in args.variants.split(",")]:
        for it in [int(v) for v in args.iters.split(",")]:

            def run():
                for b0 in range(0, pool - batch + 1, batch):
                    vb.uncertainty_fused(x[b0:b0 + batch], mean_argmax=True, scores=bool(args.scores),
                                         thresholds=(0.5, 0.4, 0.05), out_maps=out, variant=variant, tiles_per_cta=it)

            chk = vb.uncertainty_fused(x[:batch], mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05),
                                       variant=variant, tiles_per_cta=it)
            same = all(torch.equal(a, b) for a, b in zip(ref, (chk.pred_entropy, chk.expected_entropy,
                                                               chk.mutual_information, chk.mean_argmax, chk.scores)))
            run()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            nvol = (pool // batch) * batch
            gbs = nvol * V * bpv / (best * 1e-3) / 1e9
            print(f"{args.shape} variant={variant} iter={it} batch={batch}: {best / nvol * 1e3:8.1f} us/vol "
                  f"{nvol * V / best / 1e6:8.2f} Gvox/s {gbs:8.1f} GB/s = {gbs / peak:.3f} of measured peak  "
                  f"bit-identical to variant 0: {same}", flush=True)


if __name__ == "__main__":
    main()
