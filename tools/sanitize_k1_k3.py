#!/usr/bin/env python
"""Small K1 / K3 runs for compute-sanitizer (memcheck, racecheck, synccheck) on the GPU box:

    compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py

K1: every ring shape (rows per stage 8 / 4 / 5 / 2 / 1, fp32 / bf16 / fp64 with per-sample accumulators),
ragged last tiles, several tiles per CTA, and the register-stream / sample-outer variants the ring is
checked against.  K3: vector and scalar kernels, overlapping patches, a volume with an uncovered
remainder, the weighted form (importance map and separable factors)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import values_b200 as vb

g = torch.Generator(device="cuda").manual_seed(2)


def stack(b, n, c, spatial, dtype):
    x = torch.softmax(3 * torch.randn((b, n, c) + spatial, generator=g, device="cuda"), dim=2)
    return x.to(dtype)


cases = [(16, 4, (24, 40, 20), torch.float32), (12, 3, (20, 36), torch.float32), (10, 20, (24, 52), torch.float32),
         (6, 3, (33, 28), torch.float32), (7, 5, (18, 44), torch.float32), (8, 4, (16, 40), torch.bfloat16),
         (16, 2, (12, 20, 18), torch.float64), (5, 2, (16, 16, 12), torch.float64), (10, 3, (20, 22), torch.float64),
         (3, 4, (10, 14), torch.float64),
         # rows that are not 16-byte aligned: the ring in its element-strided mode (SH)
         (16, 4, (13, 21, 9), torch.float32), (5, 2, (37, 59), torch.float32), (6, 3, (33, 45), torch.float32),
         (8, 4, (17, 39), torch.bfloat16), (8, 2, (11, 13, 9), torch.float64), (5, 2, (35, 37), torch.float64),
         # ragged last stage (N = 7, 11), aligned and element-strided
         (11, 3, (24, 40), torch.float32), (7, 2, (21, 39), torch.float32), (7, 3, (16, 40), torch.bfloat16)]
for n, c, spatial, dtype in cases:
    x = stack(2, n, c, spatial, dtype)
    ref = None
    for variant, tiles in [(0, 0), (0, 2), (1, 0), (2, 0), (3, 0), (4, 0)]:
        r = vb.uncertainty_fused(x, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05),
                                 variant=variant, tiles_per_cta=tiles)
        torch.cuda.synchronize()
        out = (r.pred_entropy, r.expected_entropy, r.mutual_information, r.mean_argmax, r.scores)
        if ref is None:
            ref = tuple(t.clone() for t in out)
        assert all(torch.equal(a, b) for a, b in zip(ref[:4], out[:4])), (n, c, spatial, dtype, variant, tiles)
        assert torch.allclose(ref[4], out[4], rtol=1e-13, atol=0), (n, c, spatial, dtype, variant, tiles)
    vb.uncertainty_fused(x, maps=False, sample_argmax=True)
    vb.uncertainty_fused(x, maps=False, mean_argmax=True, sample_argmax=True, variant=4)
    vb.uncertainty_fused(x, maps=False, mean_argmax=True)
print("sanitize K1: ok")

for shape, p, ov in [((24, 20, 16), 8, 0.5), ((21, 19, 18), 8, 0.5), ((16, 16, 16), 8, 1.0)]:
    crops = vb.patch_grid(shape, p, ov)
    patches = torch.rand((3, len(crops), 2, p, p, p), generator=g, device="cuda")
    a = vb.stitch_volume(patches, crops, shape, path=0)
    b = vb.stitch_volume(patches, crops, shape, path=1)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), shape
    w = vb.gaussian_importance_map((p, p, p), device="cuda")
    wa = vb.stitch_volume(patches, crops, shape, weight=w)
    fa = vb.stitch_volume(patches, crops, shape, weight=vb.gaussian_importance_factors((p, p, p), device="cuda"))
    torch.cuda.synchronize()
    assert torch.equal(wa[0], fa[0]) and torch.equal(wa[1], fa[1]), ("separable weights", shape)
    carrier_sum = torch.zeros((3, 2) + shape, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(shape, dtype=torch.float64, device="cuda")
    vb.stitch_accumulate(patches, vb.stitching.crops_to_lo(crops, "cuda"), carrier_sum, cnt, accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(carrier_sum, a[0])
    # the DataCarrier3D form: one sample, a patch_index into a larger batch (row count unknown to the library)
    sel = torch.arange(len(crops) - 1, -1, -1, dtype=torch.int32, device="cuda")
    one = torch.zeros((1, 2) + shape, dtype=torch.float64, device="cuda")
    vb.stitch_accumulate(patches[1:2], vb.stitching.crops_to_lo(crops, "cuda")[sel.long()], one, None,
                         patch_index=sel, accumulate=True)
    torch.cuda.synchronize()
    assert torch.allclose(one[0], a[0][1], rtol=1e-12, atol=0)
print("sanitize K3: ok")
