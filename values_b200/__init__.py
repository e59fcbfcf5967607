"""values_b200 -- B200-native (sm_100a) implementation of the ValUES C2+C3 uncertainty hot
path: fused PE/EE/MI maps, the evaluation/uncertainty_aggregation strategies and the
sliding-window stitch accumulator, behind the reference's own Python entry points.

Importing this package loads values_b200/lib/libvalues_b200.so and fails loudly if it is
missing (there is no CPU fallback).  Build it with `python -m values_b200.build`.
"""
from . import _lib  # noqa: F401  (raises ValuesExtensionMissing when the .so is absent)
from .aggregation import (aggregate_uncertainties, image_level_aggregation, map_reduce,
                          normalize_maps, patch_level_aggregation, patch_max,
                          threshold_aggregation)
from . import formats, metrics, segmetrics, tester2d, threshold
from .experiment_dataloader import ExperimentDataloader
from .formats import load_to_device, reverse_axes, save_from_device
from .data_carrier import DataCarrier3D
from .metrics import (calc_ace, calib_stats, calibration_error, calibration_error_image,
                      compute_ncc, ncc_batched, ncc_main, platt_scale_confid)
from .pipeline import AggregationConfig, PipelineResult, UncertaintyPipeline
from .segmetrics import (calculate_ged, calculate_test_metrics, confusion_counts, dice_from_confusion,
                         mean_prediction_dice, seg_loss_terms)
from .sharding import AsyncScoreGather, gather_scores, shard_range, shard_sizes
from .stitching import gaussian_importance_factors, gaussian_importance_map, importance_map_from_factors, patch_grid, stitch_accumulate, stitch_volume
from .threshold import (calculate_foreground_quantile_image, calculate_threshold_image,
                        count_nonzero, find_threshold, get_foreground_quantile, quantile,
                        save_foreground_quantiles)
from .uncertainty import (FusedResult, caculcate_uncertainty_multiple_pred,
                          calculate_one_minus_msr, calculate_uncertainty,
                          calculate_uncertainty_multiple_pred, uncertainty_fused)

__all__ = [
    "calculate_uncertainty", "calculate_one_minus_msr", "caculcate_uncertainty_multiple_pred",
    "calculate_uncertainty_multiple_pred", "uncertainty_fused", "FusedResult",
    "patch_level_aggregation", "image_level_aggregation", "threshold_aggregation",
    "aggregate_uncertainties", "patch_max", "map_reduce", "normalize_maps",
    "DataCarrier3D", "patch_grid", "stitch_accumulate", "stitch_volume", "gaussian_importance_map", "gaussian_importance_factors", "importance_map_from_factors",
    "UncertaintyPipeline", "AggregationConfig", "PipelineResult",
    "shard_range", "shard_sizes", "gather_scores", "AsyncScoreGather",
    "calculate_foreground_quantile_image", "get_foreground_quantile", "save_foreground_quantiles",
    "calculate_threshold_image", "find_threshold", "quantile", "count_nonzero",
    "compute_ncc", "ncc_batched", "ncc_main", "calib_stats", "calc_ace", "platt_scale_confid",
    "calibration_error_image", "calibration_error",
    "calculate_ged", "calculate_test_metrics", "seg_loss_terms", "confusion_counts", "dice_from_confusion", "mean_prediction_dice",
    "formats", "ExperimentDataloader", "load_to_device", "save_from_device", "reverse_axes",
]
