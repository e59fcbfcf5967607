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

from . import _lib
from .aggregation import ISCLOSE_ATOL, ISCLOSE_RTOL, patch_max, patch_max_workspace_bytes
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
    chunk_bytes: int = 1 << 30                 # fp32 map scratch per chunk (32 128^3 volumes = 96 maps:
                                               # K2b's 32x64x(z-chunk) tiles need ~100 maps per launch
                                               # to fill 148 SMs x 2 CTAs for several waves; measured
                                               # 4.28 ms/step at 256 MB -> 4.05 ms at 1 GB on cfg5)
    overlap: bool = False                      # K2b and the score assembly run on a second stream, under
                                               # K1 of the next chunk -- also of the NEXT run() (K1 is
                                               # HBM-bound, K2b latency-bound); costs a second map scratch
                                               # buffer.  The result's `scores` wait for that stream when
                                               # they are first touched (PipelineResult.wait)


@dataclass
class PipelineResult:
    _scores: torch.Tensor                      # [B, 3, N_COLS] fp64, device; map order pe, ee, mi
    maps: Optional[torch.Tensor] = None        # [3, B, *S] fp32 if kept
    mean_argmax: Optional[torch.Tensor] = None  # [B, *S] uint8
    ssn: bool = False
    patch_size: Optional[List[int]] = None
    thresholds: Optional[Sequence[float]] = None
    threshold_mean: bool = True
    ready: Optional[torch.cuda.Event] = None   # overlap mode: recorded on the side stream once the table is complete

    def wait(self) -> "PipelineResult":
        """Make the current stream wait for the score table (a no-op outside overlap mode, where the
        table is produced in stream order).  Never blocks the host."""
        if self.ready is not None:
            torch.cuda.current_stream(self._scores.device).wait_event(self.ready)
        return self

    @property
    def scores(self) -> torch.Tensor:
        return self.wait()._scores

    def table_async(self):
        """(score table, event or None): the table WITHOUT making the current stream wait for it -- for
        consumers that order themselves behind the event (AsyncScoreGather.submit(..., after=event))."""
        return self._scores, self.ready

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
    """K1 -> K2b over chunks of volumes with every buffer preallocated and reused: per chunk
    the host issues two C-ABI calls (one K1 launch + a 4-byte memset, three K2b launches) and
    nothing else; both write their columns of the run's score table in place.  Never synchronises the host.
    With cfg.overlap the K2b launches and the assembly go to a side stream and run() returns with the
    main stream free for the next run's K1."""

    def __init__(self, cfg: Optional[AggregationConfig] = None):
        self.cfg = cfg or AggregationConfig()
        self._bufs: Dict[tuple, torch.Tensor] = {}
        # optional list; when set, (start_event, end_event, n_volumes) is appended per K1 launch
        # (CUDA events on the launching stream -- bench.py's roofline measurement)
        self.k1_timer: Optional[list] = None
        self._side: Dict[torch.device, dict] = {}

    def _chunk(self, B: int, V: int) -> int:
        """Volumes per chunk: as many as the map scratch holds, in EQUAL chunks (1024 volumes at 341 per
        chunk would leave a last chunk of one volume -- a K1 and three K2b launches for 3 maps)."""
        per_volume = 3 * V * 4  # three fp32 maps
        cap = max(1, min(B, self.cfg.chunk_bytes // max(per_volume, 1)))
        n_chunks = -(-B // cap) if B else 1
        return max(1, -(-B // n_chunks))

    def _buf(self, name: str, shape, dtype, dev) -> torch.Tensor:
        """Grow-only scratch tensors keyed by (name, dtype, device)."""
        key = (name, dtype, dev)
        n = int(np.prod(shape)) if len(shape) else 1
        t = self._bufs.get(key)
        if t is None or t.numel() < n:
            if t is not None and dev in self._side:   # the side stream may still be using the old scratch
                torch.cuda.current_stream(dev).wait_stream(self._side[dev]["stream"])
            t = torch.empty(max(n, 1), dtype=dtype, device=dev)
            self._bufs[key] = t
        return t[:n].view(shape)

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
        cb = self._chunk(B, V)
        overlap = cfg.overlap and patch is not None
        if overlap:
            main = torch.cuda.current_stream(dev)
            st = self._side.get(dev)
            if st is None:   # side stream, the K2b-done event of each scratch buffer, the buffer to use next
                st = self._side[dev] = {"stream": torch.cuda.Stream(dev, priority=-1), "done": [None, None], "next": 0}
            side = st["stream"]
        if keep_maps:   # K1 writes straight into the kept [B, 3, *S] buffer, chunk by chunk
            maps_all = torch.empty((B, 3) + spatial, dtype=torch.float32, device=dev)
        else:
            maps_buf = self._buf("maps", (cb, 3) + spatial, torch.float32, dev)
            maps_buf2 = self._buf("maps2", (cb, 3) + spatial, torch.float32, dev) if overlap else maps_buf
        # the score table of this run: K1 writes columns 0..2 of every row in place, K2b columns 3..6
        # (max, corner as doubles) -- nothing is assembled afterwards.  Columns of unused leading axes
        # (2-D images) and of a skipped patch level stay 0.
        scores = torch.zeros((B, 3, N_COLS), dtype=torch.float64, device=dev)
        k1_scores = scores[:, :, :3]
        flat = scores.view(B * 3, N_COLS)
        am = torch.empty((B,) + spatial, dtype=torch.uint8, device=dev) if mean_argmax else None
        k1_ws_bytes = _lib.lib.values_uncertainty_workspace_bytes(cb, V, _lib.dtype_code(probs.dtype))
        k1_ws = self._buf("k1_ws", (max(k1_ws_bytes, 8),), torch.uint8, dev)
        if patch is not None:
            ps, bb = flat[:, COL_PATCH_MAX], flat[:, COL_BBOX:COL_BBOX + 3]
            k2_ws = self._buf("k2_ws", (max(patch_max_workspace_bytes(cb * 3, spatial, patch), 8),),
                              torch.uint8, dev)   # one workspace: K2b launches are serialised on one stream
        for ci, b0 in enumerate(range(0, B, cb)):
            b1 = min(b0 + cb, B)
            nb = b1 - b0
            if overlap:
                slot = st["next"]
                st["next"] ^= 1
                if st["done"][slot] is not None and not keep_maps:
                    main.wait_event(st["done"][slot])   # the K2b that last read this scratch has finished
            else:
                slot = 0
            buf = maps_all[b0:b1] if keep_maps else (maps_buf2 if slot else maps_buf)[:nb]
            if self.k1_timer is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            uncertainty_fused(probs[b0:b1], maps=True, mean_argmax=mean_argmax, scores=True,
                              thresholds=thr, out_maps=buf, volume_major=True,
                              out_scores=k1_scores[b0:b1], out_argmax=am[b0:b1] if mean_argmax else None,
                              workspace=k1_ws)
            if self.k1_timer is not None:
                ev1.record()
                self.k1_timer.append((ev0, ev1, nb))
            if patch is not None and not overlap:
                patch_max(buf.view((3 * nb,) + spatial), patch, mean=cfg.patch_mean,
                          rtol=ISCLOSE_RTOL, atol=ISCLOSE_ATOL, out_score=ps[3 * b0:3 * b1],
                          out_bbox=bb[3 * b0:3 * b1], workspace=k2_ws)
            elif overlap:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    patch_max(buf.view((3 * nb,) + spatial), patch, mean=cfg.patch_mean,
                              rtol=ISCLOSE_RTOL, atol=ISCLOSE_ATOL, out_score=ps[3 * b0:3 * b1],
                              out_bbox=bb[3 * b0:3 * b1], workspace=k2_ws)
                    st["done"][slot] = torch.cuda.Event()
                    st["done"][slot].record(side)

        ready = None
        if overlap:   # the table is complete behind the last K2b, on the side stream
            ready = torch.cuda.Event()
            ready.record(side)
            for t in (scores,) + ((maps_all,) if keep_maps else ()):
                t.record_stream(side)     # allocated on the main stream, last written on the side stream
        return PipelineResult(_scores=scores, ready=ready, maps=maps_all.permute(1, 0, *range(2, 2 + nd)) if keep_maps else None,
                              mean_argmax=am, ssn=ssn, patch_size=patch,
                              thresholds=cfg.thresholds, threshold_mean=cfg.threshold_mean)
