#!/usr/bin/env python
"""Sweep the pipeline's chunk size and the K1/K2b two-stream overlap on one GPU (cfg5 pool).

    python tools/pipe_sweep.py [--workload cfg5] [--steps 10]
Prints one line per (chunk_bytes, overlap): ms per step, Gvox/s, mean K1 launch ms.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
import values_b200 as vb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--chunks", default="64,128,256,512")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(1)
    stack = bench.make_stack(gen, wl["pool"], wl, dev, bench.DTYPES[wl["dtype"]])
    V = int(np.prod(wl["spatial"]))
    thr = (0.5, 0.4, 0.05)
    ref_scores = None
    for mb in [int(c) for c in args.chunks.split(",")]:
        for overlap in (False, True):
            cfg = vb.AggregationConfig(patch_size=wl["patch"], thresholds=thr, chunk_bytes=mb << 20,
                                       overlap=overlap)
            pipe = vb.UncertaintyPipeline(cfg)
            ev = []
            pipe.k1_timer = ev
            for _ in range(3):
                res = pipe.run(stack, mean_argmax=True)
            torch.cuda.synchronize()
            if ref_scores is None:
                ref_scores = res.scores.clone()
            same = bool(torch.equal(res.scores, ref_scores))
            ev.clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                pipe.run(stack, mean_argmax=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            k1 = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
            print(f"chunk {mb:4d} MB overlap {int(overlap)}: {ms:7.3f} ms/step  "
                  f"{wl['pool'] * V / ms / 1e6:7.2f} Gvox/s  K1 total {k1:6.3f} ms  scores identical {same}",
                  flush=True)


if __name__ == "__main__":
    main()
