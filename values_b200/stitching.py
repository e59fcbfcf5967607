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
    weight=None,                               # fp64 [p0, p1, p2] importance map, or its three 1-D factors
                                               # (wx, wy, wz) -- see gaussian_importance_factors; None = uniform
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
    if isinstance(weight, (tuple, list)):
        # separable map: the factors go to the box kernel's shared memory (no weight traffic); whatever
        # that kernel cannot take gets the materialised map below -- bit-identical either way
        if len(weight) != 3 or any(w.dim() != 1 or w.numel() != n for w, n in zip(weight, patches.shape[3:])):
            raise ValueError("weight factors must be three 1-D tensors of lengths p0, p1, p2")
        fac = [w.to(device=dev, dtype=torch.float64).contiguous() for w in weight]
        with torch.cuda.device(dev):
            rc = _lib.lib.values_stitch_accumulate_separable(
                patches.data_ptr(), _lib.dtype_code(patches.dtype), patches.stride()[0],
                patches.stride()[1], _lib.ptr(patch_index), crop_lo.data_ptr(), fac[0].data_ptr(),
                fac[1].data_ptr(), fac[2].data_ptr(), n_sel, N, Cn,
                _lib.i64x3(patches.shape[3:]), _lib.i64x3(out_sum.shape[2:]), out_sum.data_ptr(),
                _lib.dtype_code(out_sum.dtype), _lib.ptr(out_count), int(accumulate), int(path),
                _lib.stream_ptr(dev))
        if rc != _lib.ERR_UNSUPPORTED:
            _lib.check(rc)
            return
        weight = importance_map_from_factors(fac)
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


def gaussian_importance_factors(patch_shape: Sequence[int], sigma_scale: float = 0.125,
                                device=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The three 1-D factors of the Gaussian patch weight: exp(-((i - c) / sigma)^2 / 2) with c the centre of
    the axis and sigma = sigma_scale * extent, scaled to a maximum of 1, zeros lifted to the smallest positive
    value of the factor (the usual sliding-window-inference importance map, axis by axis).  Passed as
    `weight=(wx, wy, wz)` they reach the stitch kernel without a weight map in memory."""
    fac = []
    for n in patch_shape:
        x = torch.arange(n, dtype=torch.float64)
        g = torch.exp(-0.5 * ((x - (n - 1) / 2.0) / (sigma_scale * n)) ** 2)
        g = g / g.max()
        g = torch.clamp(g, min=float(g[g > 0].min()))
        fac.append(g.to(device) if device is not None else g)
    return tuple(fac)


def importance_map_from_factors(factors) -> torch.Tensor:
    """weight[x][y][z] = fl(fl(wx[x] * wy[y]) * wz[z]) -- the rounding order of the separable kernel path."""
    wx, wy, wz = factors
    return (wx[:, None, None] * wy[None, :, None]) * wz[None, None, :]


def gaussian_importance_map(patch_shape: Sequence[int], sigma_scale: float = 0.125,
                            device=None) -> torch.Tensor:
    """Gaussian patch weight fp64 [p0, p1, p2]: the outer product of gaussian_importance_factors (centre 1).
    The reference accumulates with UNIFORM weights; this is the opt-in
    Gaussian-weighted variant BASELINE.json's north_star names."""
    return importance_map_from_factors(gaussian_importance_factors(patch_shape, sigma_scale, device))


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
