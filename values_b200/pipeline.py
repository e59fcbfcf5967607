"""Fused C2 -> C3 pipeline: softmax stacks in, per-image score table out, maps optional.

The reference hands C2 -> C3 through image files (NIfTI / tif) and re-loads every map once
per aggregation.  Here K1 writes the three maps of a chunk of volumes
into one HBM buffer sized to stay L2-resident, the image-level and threshold numerators
come out of the same sweep, and K2b (patch max) reads the maps back while they are hot.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .aggregation import ISCLOSE_ATOL, ISCLOSE_RTOL, patch_max
from .uncertainty import MAP_KEYS, uncertainty_fused

# column layout of the score table [B, 3 maps, N_COLS]
COL_SUM, COL_THR_SUM, COL_THR_COUNT, COL_PATCH_MAX, COL_BBOX = 0, 1, 2, 3, 4
N_COLS = 7


@dataclass
class AggregationConfig:
    """Parameters the reference passes through hydra (evaluation/configs/tasks/*.yaml)."""

    patch_size: Optional[object] = 10          # int or per-axis list; None = skip patch level
    patch_mean: bool = False
    thresholds: Optional[Sequence[float]] = None  # (pred_entropy, aleatoric, epistemic)
    threshold_mean: bool = True
    l2_budget_bytes: int = 80 << 20            # maps + K2b intermediate of one chunk stay L2-resident (126 MB L2)


@dataclass
class PipelineResult:
    scores: torch.Tensor                       # [B, 3, N_COLS] fp64, device; map order pe, ee, mi
    maps: Optional[torch.Tensor] = None        # [3, B, *S] fp32 if kept
    mean_argmax: Optional[torch.Tensor] = None  # [B, *S] uint8
    ssn: bool = False
    patch_size: Optional[List[int]] = None
    thresholds: Optional[Sequence[float]] = None
    threshold_mean: bool = True

    def map_index(self, unc_type: str) -> int:
        """Row of the score table for a reference uncertainty key (SSN swaps EE and MI)."""
        k = {"predictive_uncertainty": 0}.get(unc_type, None)
        if k is None:
            k = MAP_KEYS.index(unc_type)
        if self.ssn and k > 0:
            k = 3 - k
        return k

    def to_dicts(self, image_ids: Sequence[str]) -> Dict[str, Dict[str, Dict]]:
        """{unc_type: {image_id: {aggregation: result dict}}} -- the schema of
        aggregated_<unc>.json (aggregate_uncertainties.py:75, 87-95)."""
        tab = self.scores.cpu().numpy()
        out: Dict[str, Dict[str, Dict]] = {}
        for unc in MAP_KEYS:
            k = self.map_index(unc)
            per_image = {}
            for b, image_id in enumerate(image_ids):
                row = tab[b, k]
                entry = {"image_level": {"max_score": float(row[COL_SUM])}}
                if self.patch_size is not None:
                    nd = len(self.patch_size)
                    lo = row[COL_BBOX + 3 - nd:COL_BBOX + 3].astype(np.int64)
                    entry["patch_level"] = {
                        "max_score": float(row[COL_PATCH_MAX]),
                        "bounding_box": [(int(i), int(i + p)) for i, p in zip(lo, self.patch_size)],
                    }
                if self.thresholds is not None:
                    s, n = float(row[COL_THR_SUM]), float(row[COL_THR_COUNT])
                    score = s / n if (self.threshold_mean and n > 0) else s
                    entry["threshold"] = {"max_score": score,
                                          "threshold": float(self.thresholds[MAP_KEYS.index(unc)])}
                per_image[image_id] = entry
            out[unc] = per_image
        return out


class UncertaintyPipeline:
    def __init__(self, cfg: Optional[AggregationConfig] = None):
        self.cfg = cfg or AggregationConfig()
        self._maps_buf: Optional[torch.Tensor] = None
        # optional list; when set, (start_event, end_event, n_volumes) is appended per K1 launch
        # (CUDA events on the launching stream -- bench.py's roofline measurement)
        self.k1_timer: Optional[list] = None

    def _chunk(self, B: int, V: int) -> int:
        per_volume = 3 * V * (4 + 8)  # three fp32 maps + their fp64 z/x box sums (K2b workspace)
        return max(1, min(B, self.cfg.l2_budget_bytes // max(per_volume, 1)))

    def run(self, probs: torch.Tensor, ssn: bool = False, keep_maps: bool = False,
            mean_argmax: bool = False) -> PipelineResult:
        """probs [B, N, C, *S] on CUDA (B/N/C may be strided views).  No host sync."""
        cfg = self.cfg
        B = probs.shape[0]
        spatial = tuple(probs.shape[3:])
        V = int(np.prod(spatial))
        dev = probs.device
        nd = len(spatial)
        patch = cfg.patch_size
        if patch is not None and isinstance(patch, (int, np.integer)):
            patch = nd * [int(patch)]
        thr = cfg.thresholds  # given per reference key; K1 wants (pe, ee, mi) order
        if thr is not None and ssn:
            thr = (thr[0], thr[2], thr[1])
        scores = torch.zeros((B, 3, N_COLS), dtype=torch.float64, device=dev)
        am = torch.empty((B,) + spatial, dtype=torch.uint8, device=dev) if mean_argmax else None
        all_maps = torch.empty((3, B) + spatial, dtype=torch.float32, device=dev) if keep_maps else None
        cb = B if keep_maps and patch is None else self._chunk(B, V)
        if not keep_maps:
            if (self._maps_buf is None or self._maps_buf.device != dev
                    or self._maps_buf.numel() < 3 * cb * V):
                self._maps_buf = torch.empty(3 * cb * V, dtype=torch.float32, device=dev)
        for b0 in range(0, B, cb):
            b1 = min(b0 + cb, B)
            nb = b1 - b0
            if keep_maps:
                # [3, nb, *S] slice of the kept buffer is not contiguous across maps: K1 writes
                # through a contiguous scratch only when chunked; with keep_maps write per chunk
                buf = torch.empty((3, nb) + spatial, dtype=torch.float32, device=dev) if cb < B else all_maps
            else:
                buf = self._maps_buf[:3 * nb * V].view((3, nb) + spatial)
            if self.k1_timer is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            res = uncertainty_fused(probs[b0:b1], maps=True, mean_argmax=mean_argmax, scores=True,
                                    thresholds=thr, out_maps=buf)
            if self.k1_timer is not None:
                ev1.record()
                self.k1_timer.append((ev0, ev1, nb))
            scores[b0:b1, :, :3] = res.scores
            if mean_argmax:
                am[b0:b1] = res.mean_argmax
            if patch is not None:
                ps, bbox = patch_max(buf.view((3 * nb,) + spatial), patch, mean=cfg.patch_mean,
                                     rtol=ISCLOSE_RTOL, atol=ISCLOSE_ATOL)
                scores[b0:b1, :, COL_PATCH_MAX] = ps.view(3, nb).t()
                scores[b0:b1, :, COL_BBOX + 3 - nd:COL_BBOX + 3] = bbox.reshape(3, nb, nd).permute(1, 0, 2).to(torch.float64)
            if keep_maps and cb < B:
                all_maps[:, b0:b1] = buf
        return PipelineResult(scores=scores, maps=all_maps, mean_argmax=am, ssn=ssn, patch_size=patch,
                              thresholds=cfg.thresholds, threshold_mean=cfg.threshold_mean)
