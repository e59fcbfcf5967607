#!/usr/bin/env python
"""Why is K2b slow on a workload's own maps?  Times patch_max per map of one bench stack and counts the
windows / tiles near the maximum (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import values_b200 as vb

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1234 + 1000 * wl["cfg"])
x = bench.make_stack(gen, 2, wl, dev, bench.DTYPES[wl["dtype"]])
res = vb.uncertainty_fused(x, volume_major=True)
maps = torch.stack([res.pred_entropy, res.expected_entropy, res.mutual_information], 1).reshape((-1,) + tuple(wl["spatial"])).contiguous()
for i in range(maps.shape[0]):
    m = maps[i:i + 1]
    ts = []
    for path in (0, 5):
        vb.patch_max(m, 10, path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, b = vb.patch_max(m, 10, path=path)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    nd = m.dim() - 1
    pool = torch.nn.functional.avg_pool2d if nd == 2 else torch.nn.functional.avg_pool3d
    sums = pool(m.double().unsqueeze(0), 10, stride=1) * (10 ** nd)
    g = sums.max().item()
    print(f"map {i}: min {m.min().item():.4g} max {m.max().item():.4g} nan {int(torch.isnan(m).sum())} | path0 {ts[0]:.0f} us path5 {ts[1]:.0f} us | "
          f"max box sum {g:.6f} score {s.item():.6f} | windows within 1e-2 / 1e-1 of max: {int((sums >= g - 1e-2).sum())} / {int((sums >= g - 1e-1).sum())}")
