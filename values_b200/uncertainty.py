"""C2 uncertainty measures on B200: drop-in mirrors of the reference's Python entry points
(uncertainty_modeling/test_3D.py:486-534) over the fused CUDA kernel K1.

    calculate_uncertainty(softmax_preds, ssn=False)            test_3D.py:486-518
    calculate_one_minus_msr(softmax_pred)                      test_3D.py:521-525
    caculcate_uncertainty_multiple_pred(test_datacarrier, ssn) test_3D.py:528-534 (sic)

plus the batched form `uncertainty_fused` the pipeline and the benchmark use.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib

MAP_KEYS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")


@dataclass
class FusedResult:
    """Outputs of one K1 sweep over a batch of stacks [B, N, C, *S]."""

    pred_entropy: Optional[torch.Tensor]      # [B, *S] fp32
    expected_entropy: Optional[torch.Tensor]  # [B, *S] fp32
    mutual_information: Optional[torch.Tensor]  # [B, *S] fp32
    mean_argmax: Optional[torch.Tensor]       # [B, *S] uint8
    sample_argmax: Optional[torch.Tensor]     # [B, N, *S] uint8
    scores: Optional[torch.Tensor]            # [B, 3, 3] fp64: map (pe, ee, mi) x {sum, thr_sum, thr_count}

    def as_dict(self, b: int, ssn: bool = False) -> Dict[str, torch.Tensor]:
        """Reference return layout for image b (key swap when `ssn`, test_3D.py:510-516)."""
        ee, mi = self.expected_entropy[b], self.mutual_information[b]
        return {
            "pred_entropy": self.pred_entropy[b],
            "aleatoric_uncertainty": mi if ssn else ee,
            "epistemic_uncertainty": ee if ssn else mi,
        }


def _to_device_tensor(x, device: torch.device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor or numpy array, got {type(x)}")
    if x.dtype in (torch.float16,):
        x = x.float()
    if x.dtype not in (torch.float32, torch.float64, torch.bfloat16):
        raise TypeError(f"softmax stack must be float32/float64/bfloat16, got {x.dtype}")
    if x.device.type != "cuda":
        x = x.to(device, non_blocking=True)
    return x


def _spatial_contiguous(x: torch.Tensor, lead: int) -> torch.Tensor:
    """Make the trailing spatial block contiguous (leading axes may keep any stride)."""
    expect = 1
    for size, stride in zip(reversed(x.shape[lead:]), reversed(x.stride()[lead:])):
        if size != 1 and stride != expect:
            return x.contiguous()
        expect *= size
    if any(s < 0 for s in x.stride()[:lead]):
        return x.contiguous()
    return x


def uncertainty_fused(
    probs: torch.Tensor,
    *,
    maps: bool = True,
    mean_argmax: bool = False,
    sample_argmax: bool = False,
    scores: bool = False,
    thresholds: Optional[Sequence[float]] = None,
    out_maps: Optional[torch.Tensor] = None,
    volume_major: bool = False,
    out_scores: Optional[torch.Tensor] = None,
    out_argmax: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    variant: int = 0,
    tiles_per_cta: int = 0,
) -> FusedResult:
    """One HBM sweep over `probs` [B, N, C, *S] (CUDA; B/N/C axes may be strided views, e.g.
    a permuted [N, B, C, H, W] stack as test_2D.py:317 builds it).

    thresholds: (pe, ee, mi) scalars for the fused threshold sums (aggregate_uncertainties.py:61-62).
    out_maps:   optional preallocated fp32 buffer to write the maps into: [3, B, *S] (pe, ee, mi
                planes, map-major) or, with volume_major=True, [B, 3, *S] -- the layout K2b
                consumes as 3B consecutive maps.
    out_scores / out_argmax / workspace: optional preallocated outputs ([B, 3, 3] fp64 -- or a
                [B, 3, W >= 3] fp64 row-strided view such as table[:, :, :3] of a score table, written in
                place --, [B, *S] uint8, uint8 scratch) so that a pipeline can run without allocating.
    """
    if probs.dim() < 3:
        raise ValueError("probs must be [B, N, C, *spatial]")
    dev = probs.device
    if dev.type != "cuda":
        raise RuntimeError("uncertainty_fused expects a CUDA tensor (no CPU fallback)")
    probs = _spatial_contiguous(probs, 3)
    B, N, C = probs.shape[:3]
    spatial = tuple(probs.shape[3:])
    V = int(np.prod(spatial)) if spatial else 1
    sb, sn, sc = probs.stride()[:3]
    pe = ee = mi = None
    map_stride_b = 0
    if maps:
        want = ((B, 3) if volume_major else (3, B)) + spatial
        if out_maps is None:
            out_maps = torch.empty(want, dtype=torch.float32, device=dev)
        elif (tuple(out_maps.shape) != want or out_maps.dtype != torch.float32
              or not out_maps.is_contiguous() or out_maps.device != dev):
            raise ValueError(f"out_maps must be a contiguous fp32 {want} CUDA tensor")
        if volume_major:
            pe, ee, mi = out_maps[:, 0], out_maps[:, 1], out_maps[:, 2]
            map_stride_b = 3 * V
        else:
            pe, ee, mi = out_maps[0], out_maps[1], out_maps[2]
    if mean_argmax:
        if out_argmax is None:
            out_argmax = torch.empty((B,) + spatial, dtype=torch.uint8, device=dev)
        elif tuple(out_argmax.shape) != (B,) + spatial or out_argmax.dtype != torch.uint8 \
                or not out_argmax.is_contiguous():
            raise ValueError("out_argmax must be a contiguous uint8 [B, *S] tensor")
    am = out_argmax if mean_argmax else None
    sam = torch.empty((B, N) + spatial, dtype=torch.uint8, device=dev) if sample_argmax else None
    sc_out = ws = None
    ws_bytes, score_stride = 0, 3
    thr = None
    if scores:
        if out_scores is None:
            out_scores = torch.empty((B, 3, 3), dtype=torch.float64, device=dev)
        elif tuple(out_scores.shape) != (B, 3, 3) or out_scores.dtype != torch.float64 \
                or out_scores.stride(2) != 1 or out_scores.stride(1) < 3 \
                or (B > 1 and out_scores.stride(0) != 3 * out_scores.stride(1)):
            raise ValueError("out_scores must be an fp64 [B, 3, 3] tensor, rows contiguous and evenly strided")
        sc_out = out_scores
        score_stride = out_scores.stride(1)
        ws_bytes = _lib.lib.values_uncertainty_workspace_bytes(B, V, _lib.dtype_code(probs.dtype))
        if workspace is not None and workspace.numel() * workspace.element_size() >= ws_bytes:
            ws = workspace
        else:
            ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
        if thresholds is not None:
            if len(thresholds) != 3:
                raise ValueError("thresholds must be (pe, ee, mi)")
            thr = _lib.dbl_array(thresholds)
    with torch.cuda.device(dev):
        rc = _lib.lib.values_uncertainty_fused(
            probs.data_ptr(), _lib.dtype_code(probs.dtype), B, N, C, V, sb, sn, sc,
            _lib.ptr(pe), _lib.ptr(ee), _lib.ptr(mi), map_stride_b, _lib.ptr(am), _lib.ptr(sam),
            _lib.ptr(sc_out), score_stride, thr, _lib.ptr(ws), ws_bytes, int(variant), int(tiles_per_cta),
            _lib.stream_ptr(dev))
    _lib.check(rc)
    return FusedResult(pe, ee, mi, am, sam, sc_out)


def calculate_uncertainty(softmax_preds: torch.Tensor, ssn: bool = False) -> Dict[str, torch.Tensor]:
    """Drop-in for uncertainty_modeling/test_3D.py:486-518.

    softmax_preds [N, C, *S] (torch tensor on any device, or numpy) -> dict of fp32 maps [*S]
    on the input's device: pred_entropy, aleatoric_uncertainty (expected entropy),
    epistemic_uncertainty (mutual information); the last two swapped when `ssn`.
    """
    dev = _lib.require_cuda()
    src_device = softmax_preds.device if isinstance(softmax_preds, torch.Tensor) else torch.device("cpu")
    x = _to_device_tensor(softmax_preds, dev)
    if x.dim() < 2:
        raise ValueError("softmax_preds must be [N, C, *spatial]")
    res = uncertainty_fused(x.unsqueeze(0))
    if ssn:
        print("mutual information is aleatoric unc")  # the reference prints this (test_3D.py:514)
    out = res.as_dict(0, ssn)
    if src_device.type != "cuda":
        out = {k: v.to(src_device) for k, v in out.items()}
    return out


def calculate_one_minus_msr(softmax_pred: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Drop-in for test_3D.py:521-525: {"pred_entropy": 1 - max_c p_c}, input dtype/device."""
    dev = _lib.require_cuda()
    src_device = softmax_pred.device if isinstance(softmax_pred, torch.Tensor) else torch.device("cpu")
    x = _spatial_contiguous(_to_device_tensor(softmax_pred, dev), 1)
    Cn = x.shape[0]
    spatial = tuple(x.shape[1:])
    V = int(np.prod(spatial)) if spatial else 1
    out = torch.empty(spatial, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib.values_one_minus_msr(x.data_ptr(), _lib.dtype_code(x.dtype), 1, Cn, V,
                                           0, x.stride()[0], out.data_ptr(),
                                           _lib.stream_ptr(x.device))
    _lib.check(rc)
    if src_device.type != "cuda":
        out = out.to(src_device)
    return {"pred_entropy": out}


def caculcate_uncertainty_multiple_pred(test_datacarrier, ssn: bool = False) -> None:
    """Drop-in for test_3D.py:528-534 (name misspelt in the reference; kept).  For every image
    in `test_datacarrier.data` the RAW accumulated `softmax_pred` (not count-normalised, as the
    reference does) goes through K1 and the three maps are stored in place."""
    for _, value in test_datacarrier.data.items():
        sp = value["softmax_pred"]
        value.update(calculate_uncertainty(sp, ssn))


calculate_uncertainty_multiple_pred = caculcate_uncertainty_multiple_pred  # correctly spelt alias
