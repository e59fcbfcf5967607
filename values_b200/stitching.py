"""Sliding-window stitching on B200 (kernel K3) and the patch grid that feeds it.

    patch_grid(image_shape, patch_size, patch_overlap)
        crop-index loop of uncertainty_modeling/lidc_idri_datamodule_3D.py:719-736
    stitch_accumulate / stitch_volume
        the `+=` slab updates of DataCarrier3D.concat_data (data_carrier_3D.py:154-179)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

Crop = Tuple[Tuple[int, int], Tuple[int, int], Tuple[int, int]]


def patch_grid(image_shape: Sequence[int], patch_size: int, patch_overlap: float) -> List[Crop]:
    """Crop tuples ((x0,x1),(y0,y1),(z0,z1)), z outer -> y -> x inner, stride
    int(patch_size*patch_overlap); the trailing remainder of each axis is never covered
    (lidc_idri_datamodule_3D.py:719-736, toy_datamodule_3D.py:637-654)."""
    stride = int(patch_size * patch_overlap)
    if stride <= 0:
        raise ValueError("int(patch_size * patch_overlap) must be >= 1")
    xs = range(0, image_shape[0] - patch_size + 1, stride)
    ys = range(0, image_shape[1] - patch_size + 1, stride)
    zs = range(0, image_shape[2] - patch_size + 1, stride)
    return [((x, x + patch_size), (y, y + patch_size), (z, z + patch_size))
            for z in zs for y in ys for x in xs]


def crops_to_lo(crops: Sequence[Crop], device) -> torch.Tensor:
    lo = np.asarray([[c[0][0], c[1][0], c[2][0]] for c in crops], dtype=np.int32).reshape(-1, 3)
    return torch.from_numpy(lo).to(device)


def stitch_accumulate(
    patches: torch.Tensor,          # [N, P, C, p0, p1, p2] CUDA (N, P axes may be strided)
    crop_lo: torch.Tensor,          # int32 [n_sel, 3] CUDA
    out_sum: torch.Tensor,          # [N, C, X, Y, Z] CUDA fp64/fp32, contiguous
    out_count: Optional[torch.Tensor] = None,  # fp64 [X, Y, Z]
    patch_index: Optional[torch.Tensor] = None,  # int32 [n_sel] or None (identity)
    accumulate: bool = True,
    weight: Optional[torch.Tensor] = None,     # fp64 [p0, p1, p2] importance map (None = uniform)
    path: int = 0,                             # 0 automatic (box kernel), 1 scalar, 2 vector kernel (tests; same results)
) -> None:
    if patches.device.type != "cuda":
        raise RuntimeError("stitch_accumulate expects CUDA tensors (no CPU fallback)")
    if patches.dim() != 6 or out_sum.dim() != 5:
        raise ValueError("patches must be [N,P,C,p0,p1,p2] and out_sum [N,C,X,Y,Z]")
    N, P, Cn = patches.shape[:3]
    if not patches[0, 0].is_contiguous():
        patches = patches.contiguous()
    if out_sum.shape[:2] != (N, Cn) or not out_sum.is_contiguous():
        raise ValueError("out_sum must be contiguous [N, C, X, Y, Z] matching the patches")
    n_sel = crop_lo.shape[0]
    if crop_lo.dtype != torch.int32 or not crop_lo.is_contiguous():
        crop_lo = crop_lo.to(torch.int32).contiguous()
    if patch_index is not None:
        patch_index = patch_index.to(torch.int32).contiguous()
        if patch_index.numel() != n_sel:
            raise ValueError("patch_index and crop_lo disagree")
    elif n_sel != P:
        raise ValueError("crop_lo must have one row per patch when patch_index is None")
    if out_count is not None and (out_count.dtype != torch.float64 or not out_count.is_contiguous()
                                  or tuple(out_count.shape) != tuple(out_sum.shape[2:])):
        raise ValueError("out_count must be contiguous fp64 [X, Y, Z]")
    dev = patches.device
    if weight is not None:
        if tuple(weight.shape) != tuple(patches.shape[3:]):
            raise ValueError("weight must have the patch shape [p0, p1, p2]")
        weight = weight.to(device=dev, dtype=torch.float64).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib.values_stitch_accumulate_weighted(
            patches.data_ptr(), _lib.dtype_code(patches.dtype), patches.stride()[0],
            patches.stride()[1], _lib.ptr(patch_index), crop_lo.data_ptr(), _lib.ptr(weight), n_sel, N, Cn,
            _lib.i64x3(patches.shape[3:]), _lib.i64x3(out_sum.shape[2:]), out_sum.data_ptr(),
            _lib.dtype_code(out_sum.dtype), _lib.ptr(out_count), int(accumulate), int(path),
            _lib.stream_ptr(dev))
    _lib.check(rc)


def gaussian_importance_map(patch_shape: Sequence[int], sigma_scale: float = 0.125,
                            device=None) -> torch.Tensor:
    """Separable Gaussian patch weight, centre 1, sigma = sigma_scale * extent per axis, zeros
    lifted to the smallest positive weight (the usual sliding-window-inference importance map).
    The reference accumulates with UNIFORM weights; this is the opt-in
    Gaussian-weighted variant BASELINE.json's north_star names.  fp64 [p0, p1, p2]."""
    axes = []
    for n in patch_shape:
        c = (n - 1) / 2.0
        x = torch.arange(n, dtype=torch.float64)
        axes.append(torch.exp(-0.5 * ((x - c) / (sigma_scale * n)) ** 2))
    w = axes[0][:, None, None] * axes[1][None, :, None] * axes[2][None, None, :]
    w = w / w.max()
    w = torch.clamp(w, min=float(w[w > 0].min()))
    return w.to(device) if device is not None else w


def stitch_volume(patches: torch.Tensor, crops, vol_shape: Sequence[int],
                  out_dtype: torch.dtype = torch.float64,
                  weight: Optional[torch.Tensor] = None, path: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """All patches of one volume at once: patches [N, P, C, p,p,p] + P crops ->
    (raw sum [N, C, X, Y, Z], count fp64 [X, Y, Z]); every output voxel is written exactly
    once (uncovered voxels are 0, as in the reference).  With `weight` [p,p,p] the sum is
    weighted and `count` is the sum of weights (divide to normalise)."""
    dev = patches.device
    crop_lo = crops if isinstance(crops, torch.Tensor) else crops_to_lo(crops, dev)
    N, _, Cn = patches.shape[:3]
    out = torch.empty((N, Cn) + tuple(vol_shape), dtype=out_dtype, device=dev)
    cnt = torch.empty(tuple(vol_shape), dtype=torch.float64, device=dev)
    stitch_accumulate(patches, crop_lo, out, cnt, accumulate=False, weight=weight, path=path)
    return out, cnt
