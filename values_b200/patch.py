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
        "calculate_test_metrics": segmetrics.calculate_test_metrics,
        "calculate_metrics": data_carrier.calculate_metrics,
        "DataCarrier3D": data_carrier.DataCarrier3D,
    },
    "uncertainty_modeling.test_2D": {
        "calculate_uncertainty": uncertainty.calculate_uncertainty,
        "calculate_one_minus_msr": uncertainty.calculate_one_minus_msr,
    },
    "uncertainty_modeling.data_carrier_3D": {"DataCarrier3D": data_carrier.DataCarrier3D},
    # lightning_experiment.py:23 binds the class at import time (`from data_carrier_3D import DataCarrier3D`)
    # and instantiates it at :79; its test_step / on_test_end (:399, 405) only call concat_data / save_data
    "uncertainty_modeling.lightning_experiment": {"DataCarrier3D": data_carrier.DataCarrier3D},
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


def _file_io_patches() -> dict:
    """SURVEY 8 f4, opt-in: the reference's medpy.io.load / save names and its ExperimentDataloader.
    (Our ExperimentDataloader does not instantiate hydra datamodule configs; experiments that set
    `datamodule_config` should keep the reference class and only take the load / save rebinding.)"""
    from . import experiment_dataloader, formats

    io = {"load": formats.load, "save": formats.save}
    return {
        "uncertainty_modeling.data_carrier_3D": dict(io),
        "evaluation.experiment_dataloader": dict(io, ExperimentDataloader=experiment_dataloader.ExperimentDataloader),
        "evaluation.uncertainty_aggregation.aggregate_uncertainties": {"load": formats.load},
        "evaluation.uncertainty_aggregation.find_threshold": {"load": formats.load},
    }


def install(import_missing: bool = False, file_io: bool = False) -> dict:
    """Rebind the hot-path names in every reference module that is already imported (or, with
    import_missing=True, importable); file_io=True also rebinds the medpy.io names and
    ExperimentDataloader.  Returns {module: {name: original}} for `uninstall`."""
    saved = {}
    patches = {k: dict(v) for k, v in _PATCHES.items()}
    if file_io:
        for mod_name, names in _file_io_patches().items():
            patches.setdefault(mod_name, {}).update(names)
    for mod_name, names in patches.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                mod = None
        # The reference runs with uncertainty_modeling/ itself on sys.path and imports its own files
        # under their BARE names too (`from data_carrier_3D import ...`, lightning_experiment.py:23;
        # `from loss_modules import ...`, `from main import set_seed`, test_3D.py:23-24): a module can
        # therefore exist twice, as `uncertainty_modeling.x` and as `x`.  Both copies are rebound.
        bare = sys.modules.get(mod_name.rpartition(".")[2]) if mod_name.startswith("uncertainty_modeling.") else None
        for key, target in ((mod_name, mod), (mod_name.rpartition(".")[2], bare)):
            if target is None or (key != mod_name and target is mod):
                continue
            saved[key] = {}
            for name, fn in names.items():
                if hasattr(target, name):
                    original = getattr(target, name)
                    saved[key][name] = original
                    if name == "calculate_metrics":   # the numpy carrier of the reference still goes to its own code
                        target._values_b200_original_calculate_metrics = original
                    setattr(target, name, fn)
    return saved


def uninstall(saved: dict) -> None:
    for mod_name, names in saved.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for name, fn in names.items():
            setattr(mod, name, fn)
