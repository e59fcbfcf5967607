"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (values_b200/).

Quantifies how far the CUDA path's uncertainty maps are from the reference / oracle maps where the
north star asks for bit-exact DECISIONS (arg-max and threshold masks, aggregate_uncertainties.py:61-62
`image >= threshold`): the fp32 maps carry a float tolerance (different summation order and log
implementation), so a voxel whose reference value lies within that tolerance of the threshold can
fall on the other side.  `parity_counts` counts those voxels and checks that every one of them is
explained by the tolerance; the tests assert it and bench.py prints the counts."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

MAPS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units of fp32 representable numbers (order-preserving integer keys)."""
    def key(x):
        i = np.ascontiguousarray(x, dtype=np.float32).view(np.int32).astype(np.int64)
        return np.where(i < 0, np.int64(-2**31) - i, i)
    return np.abs(key(a) - key(b))


def parity_counts(got: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], thresholds: Sequence[float],
                  rtol: float = 1e-5, atol: float = 1e-6, got_argmax: Optional[np.ndarray] = None,
                  ref_argmax: Optional[np.ndarray] = None, class_means: Optional[np.ndarray] = None) -> Dict:
    """got / ref: the three fp32 maps; thresholds: one per map.  Returns the counts and `explained`
    (True iff every threshold-mask flip has |ref - thr| <= atol + rtol |ref|, i.e. lies inside the
    map tolerance, and every arg-max mismatch is a tie of the two largest class means within 2 ulp)."""
    out = {"voxels": int(np.asarray(ref[MAPS[0]]).size), "explained": True}
    for k, thr in zip(MAPS, thresholds):
        a = np.asarray(got[k], dtype=np.float32)
        b = np.asarray(ref[k], dtype=np.float32)
        diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
        t32 = np.float32(thr)                      # numpy compares an fp32 image with a Python float in fp32
        flips = (a >= t32) != (b >= t32)
        n_flip = int(flips.sum())
        ok = bool(np.all(np.abs(b[flips].astype(np.float64) - float(t32)) <= atol + rtol * np.abs(b[flips])))
        out[k] = {
            "differing_voxels": int((a != b).sum()),
            "max_abs_diff": float(diff.max()) if diff.size else 0.0,
            "max_ulp": int(ulp_distance(a, b).max()) if diff.size else 0,
            "beyond_tolerance": int((diff > atol + rtol * np.abs(b)).sum()),
            "threshold": float(thr),
            "mask_flips": n_flip,
            "mask_size_ref": int((b >= t32).sum()),
        }
        out["explained"] = out["explained"] and ok and out[k]["beyond_tolerance"] == 0
    if got_argmax is not None and ref_argmax is not None:
        bad = np.asarray(got_argmax) != np.asarray(ref_argmax)
        out["argmax_mismatches"] = int(bad.sum())
        if bad.any():
            if class_means is None:
                out["explained"] = False
            else:
                m = np.sort(np.asarray(class_means, dtype=np.float64), axis=0)
                gap = (m[-1] - m[-2])[bad]
                out["explained"] = out["explained"] and bool(np.all(gap <= 4e-7 * np.abs(m[-1][bad])))
    return out
