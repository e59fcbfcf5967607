#!/usr/bin/env python
"""Micro-benchmark of the f4 device pieces on the GPU box: values_reverse_axes against its HBM
roofline (2 x element bytes per element), and the two ways of getting a medpy-style [x, y, z]
view onto the device (strided host copy + upload, as numpy / the reference would, vs upload in
memory order + reversal on the GPU).

    python tools/formats_bench.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import values_b200 as vb


def gpu_time(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {}
    peak = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except Exception:
        pass
    for shape in [(128, 128, 128), (256, 256, 256), (512, 512, 512), (1024, 2048)]:
        for dt in (torch.uint8, torch.float32, torch.float64):
            x = (torch.rand(shape, device="cuda") * 100).to(dt)
            ms = gpu_time(lambda: vb.reverse_axes(x))
            gbs = 2 * x.numel() * x.element_size() / ms / 1e6
            key = f"reverse_axes {'x'.join(map(str, shape))} {str(dt).split('.')[1]}"
            out[key] = {"us": round(ms * 1e3, 1), "GB/s": round(gbs, 1),
                        "frac_of_measured_hbm_peak": round(gbs / peak, 3) if peak else None}
            print(key, out[key], flush=True)
    # host view -> device: the reference's way vs ours
    a = np.asfortranarray(np.random.default_rng(0).random((256, 256, 256)))      # what medpy.io.load returns
    t0 = time.perf_counter()
    for _ in range(3):
        ref = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        torch.cuda.synchronize()
    t_host = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3):
        mine = vb.reverse_axes(torch.from_numpy(a.T).cuda())
        torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 3
    assert torch.equal(ref, mine)
    out["fortran view 256^3 f64 -> device"] = {"strided host copy + upload ms": round(t_host * 1e3, 1),
                                               "upload + GPU reversal ms": round(t_dev * 1e3, 1)}
    print(out["fortran view 256^3 f64 -> device"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
