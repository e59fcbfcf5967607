#!/usr/bin/env python
"""DataCarrier3D.concat_data's calling pattern (data_carrier_3D.py:99-179): the patches of a volume arrive a
batch at a time and are accumulated on top of the sums -- many small stitch calls with accumulate=True
against one call with every patch.

    python tools/concat_probe.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import values_b200 as vb
dev = torch.device("cuda", 0); g = torch.Generator(device=dev).manual_seed(0)
vol, p = (256, 256, 256), 64
crops = vb.patch_grid(vol, p, 0.5)
patches = torch.rand((1, len(crops), 2, p, p, p), generator=g, device=dev)
lo = vb.stitching.crops_to_lo(crops, dev)
def once():
    out = torch.zeros((1, 2) + vol, dtype=torch.float64, device=dev); cnt = torch.zeros(vol, dtype=torch.float64, device=dev)
    vb.stitch_accumulate(patches, lo, out, cnt, accumulate=True)
    return out, cnt
def batched(bs):
    out = torch.zeros((1, 2) + vol, dtype=torch.float64, device=dev); cnt = torch.zeros(vol, dtype=torch.float64, device=dev)
    for i in range(0, len(crops), bs):
        idx = torch.arange(i, min(i + bs, len(crops)), dtype=torch.int32, device=dev)
        vb.stitch_accumulate(patches, lo[i:i + bs], out, cnt, patch_index=idx, accumulate=True)
    return out, cnt
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1), r
t1, r1 = timed(once)
for bs in (2, 8):
    tb, rb = timed(lambda: batched(bs))
    print(f"343 patches into 256^3 (N=1, C=2): one call {t1:.3f} ms; batches of {bs}: {tb:.3f} ms; identical: {torch.equal(r1[0], rb[0]) and torch.equal(r1[1], rb[1])}")
