#!/usr/bin/env python
"""Small K2b / formats run for compute-sanitizer (memcheck, racecheck) on the GPU box:

    compute-sanitizer --tool racecheck python tools/sanitize_k2.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb
from values_b200 import _lib

g = torch.Generator(device="cuda").manual_seed(1)
for shape in [(40, 64, 96), (23, 50, 76), (21, 45, 140)]:
    maps = torch.rand((3,) + shape, generator=g, device="cuda")
    ref = None
    for path in (5, 4, 0):
        s, b = vb.patch_max(maps, 10, path=path)
        torch.cuda.synchronize()
        if ref is None:
            ref = (s.clone(), b.clone())
        assert torch.equal(ref[1], b) and torch.allclose(ref[0], s, rtol=1e-13, atol=0), (shape, path)
for shape in [(23, 50, 77), (21, 30, 61), (70, 131)]:   # rows not 16-byte aligned: the pitched scratch copy
    maps = torch.rand((3,) + shape, generator=g, device="cuda") * 3.0
    s0, b0 = vb.patch_max(maps, 10, path=0)
    s5, b5 = vb.patch_max(maps, 10, path=5)
    torch.cuda.synchronize()
    assert torch.equal(b0, b5) and torch.allclose(s0, s5, rtol=1e-13, atol=0), shape
for shape in [(70, 132), (41, 300)]:     # 2-D images: the filter marches along y
    maps = torch.rand((5,) + shape, generator=g, device="cuda") * 3.0
    s0, b0 = vb.patch_max(maps, 10, path=0)
    s5, b5 = vb.patch_max(maps, 10, path=5)
    torch.cuda.synchronize()
    assert torch.equal(b0, b5) and torch.allclose(s0, s5, rtol=1e-13, atol=0), shape
x = torch.rand((33, 17, 70), generator=g, device="cuda")
assert torch.equal(vb.reverse_axes(x), x.permute(2, 1, 0).contiguous())
torch.cuda.synchronize()
print("sanitize_k2: ok")
