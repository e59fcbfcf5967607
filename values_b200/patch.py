"""Install the B200 path into the reference's own module namespaces, so existing callers
(`from uncertainty_modeling.test_3D import calculate_uncertainty`, hydra `_target_` strings
like evaluation.uncertainty_aggregation.aggregate_uncertainties.patch_level_aggregation) pick
up the CUDA implementations without source changes.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

from . import aggregation, data_carrier, metrics, segmetrics, threshold, uncertainty

_PATCHES = {
    "uncertainty_modeling.test_3D": {
        "calculate_uncertainty": uncertainty.calculate_uncertainty,
        "calculate_one_minus_msr": uncertainty.calculate_one_minus_msr,
        "caculcate_uncertainty_multiple_pred": uncertainty.caculcate_uncertainty_multiple_pred,
        "calculate_ged": segmetrics.calculate_ged,
        "DataCarrier3D": data_carrier.DataCarrier3D,
    },
    "uncertainty_modeling.test_2D": {
        "calculate_uncertainty": uncertainty.calculate_uncertainty,
        "calculate_one_minus_msr": uncertainty.calculate_one_minus_msr,
    },
    "uncertainty_modeling.data_carrier_3D": {"DataCarrier3D": data_carrier.DataCarrier3D},
    "evaluation.uncertainty_aggregation.aggregate_uncertainties": {
        "patch_level_aggregation": aggregation.patch_level_aggregation,
        "image_level_aggregation": aggregation.image_level_aggregation,
        "threshold_aggregation": aggregation.threshold_aggregation,
        "aggregate_uncertainties": aggregation.aggregate_uncertainties,
    },
    "evaluation.uncertainty_aggregation.find_threshold": {
        "calculate_foreground_quantile_image": threshold.calculate_foreground_quantile_image,
        "get_foreground_quantile": threshold.get_foreground_quantile,
        "calculate_threshold_image": threshold.calculate_threshold_image,
        "find_threshold": threshold.find_threshold,
    },
    "evaluation.metrics.ncc": {"compute_ncc": metrics.compute_ncc, "main": metrics.ncc_main},
    "evaluation.metrics.ace": {
        "calib_stats": metrics.calib_stats,
        "calc_ace": metrics.calc_ace,
        "calibration_error": metrics.calibration_error,
    },
}


def install(import_missing: bool = False) -> dict:
    """Rebind the hot-path names in every reference module that is already imported (or, with
    import_missing=True, importable).  Returns {module: {name: original}} for `uninstall`."""
    saved = {}
    for mod_name, names in _PATCHES.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                mod = None
        if mod is None:
            continue
        saved[mod_name] = {}
        for name, fn in names.items():
            if hasattr(mod, name):
                saved[mod_name][name] = getattr(mod, name)
                setattr(mod, name, fn)
    return saved


def uninstall(saved: dict) -> None:
    for mod_name, names in saved.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for name, fn in names.items():
            setattr(mod, name, fn)
