import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import values_b200 as vb
from values_b200 import _lib
torch.manual_seed(0)
x = torch.softmax(torch.randn(8, 16, 4, 128, 128, 128, device="cuda") * 3, dim=2)
def run(v, it):
    _lib.lib.values_debug_set_k1_variant(v); _lib.lib.values_debug_set_k1_iter(it)
    r = vb.uncertainty_fused(x, mean_argmax=True, scores=True, thresholds=(0.5, 0.4, 0.05))
    torch.cuda.synchronize()
    return r
ref = run(0, 0)
for v, it in [(0, 1), (7, 1), (7, 4), (7, 2), (10, 4), (13, 1)]:
    r = run(v, it)
    for name in ("pred_entropy", "expected_entropy", "mutual_information", "mean_argmax", "scores"):
        a, b = getattr(ref, name), getattr(r, name)
        ne = (a != b)
        if ne.any():
            idx = ne.nonzero()[0].tolist()
            print(v, it, name, "mismatches", int(ne.sum()), "first at", idx, a[tuple(idx)].item(), b[tuple(idx)].item(),
                  "max abs diff", (a.double() - b.double()).abs().max().item())
    print(v, it, "done")
