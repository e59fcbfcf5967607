#!/usr/bin/env python
"""One K2b call on a given shape (for compute-sanitizer / ncu on the GPU box):
    python tools/k2_case.py 1024,2048 [maps] [path]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb

shape = tuple(int(v) for v in sys.argv[1].split(","))
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(1)
maps = torch.rand((M,) + shape, generator=g, device="cuda")
s, b = vb.patch_max(maps, 10, path=path)
torch.cuda.synchronize()
s5, b5 = vb.patch_max(maps, 10, path=5)
torch.cuda.synchronize()
assert torch.equal(b, b5) and torch.allclose(s, s5, rtol=1e-13, atol=0), (s, s5, b, b5)
print("k2_case ok", shape, M, path, s.tolist()[:2], b.tolist()[:2])
