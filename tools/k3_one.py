import sys; sys.path.insert(0, '/root/repo')
import torch, values_b200 as vb
dev = torch.device('cuda')
vol, p = (256,256,256), 64
crops = vb.patch_grid(vol, p, 0.5)
g = torch.Generator(device=dev).manual_seed(0)
patches = torch.rand((8, len(crops), 2, p, p, p), generator=g, device=dev)
lo = vb.stitching.crops_to_lo(crops, dev)
out = torch.empty((8, 2) + vol, dtype=torch.float64, device=dev)
cnt = torch.empty(vol, dtype=torch.float64, device=dev)
for _ in range(3):
    vb.stitch_accumulate(patches, lo, out, cnt, accumulate=False)
torch.cuda.synchronize()
