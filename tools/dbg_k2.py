import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import values_b200 as vb
from values_b200 import _lib
M, shape = 24, (128, 128, 128)
g = torch.Generator(device="cuda").manual_seed(3)
maps = torch.rand((M,) + shape, generator=g, device="cuda")
nbytes = vb.aggregation.patch_max_workspace_bytes(M, shape, 10)
ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
score, bbox = vb.patch_max(maps, 10, workspace=ws)
torch.cuda.synchronize()
wd = ws.view(torch.float64)
# ntiles from size: nbytes = (M*nt + 3M)*8 + M*65*4
nt = (nbytes - M * 65 * 4) // 8 // M - 3
print("ntiles", nt)
act = ws[(M * nt + 3 * M) * 8:].view(torch.int32).view(M, 65)
print("n_active per map:", act[:, 0].tolist())
tm = wd[:M * nt].view(M, nt)
print("tile max spread map0:", tm[0].min().item(), tm[0].max().item(), "gmax", wd[M * nt].item())
