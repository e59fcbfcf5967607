"""CPU tests of the hand-off file formats: values_b200.formats against
NIfTI-1 files built field by field from the specification (independently of the writer), against
cv2 for the 2D formats, and round trips; the ExperimentDataloader mirror on a results directory.
No kernel runs here (host I/O only)."""
import gzip
import os
import struct
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from values_b200 import formats
from values_b200.experiment_dataloader import ExperimentDataloader

CODES = {np.uint8: 2, np.int16: 4, np.int32: 8, np.float32: 16, np.float64: 64}


def spec_nifti(arr_xyz, endian="<", slope=1.0, inter=0.0, pixdim=(1.0, 1.0, 1.0), sform=True):
    """A NIfTI-1 single file written field by field at the offsets of nifti1.h."""
    dt = np.dtype(arr_xyz.dtype)
    h = bytearray(352)
    e = endian
    struct.pack_into(e + "i", h, 0, 348)
    nd = arr_xyz.ndim
    dim = [nd] + list(arr_xyz.shape) + [1] * (7 - nd)
    struct.pack_into(e + "8h", h, 40, *dim)
    struct.pack_into(e + "hh", h, 70, CODES[dt.type], dt.itemsize * 8)
    struct.pack_into(e + "8f", h, 76, 1.0, *pixdim, *([1.0] * (7 - len(pixdim))))
    struct.pack_into(e + "3f", h, 108, 352.0, slope, inter)
    if sform:
        struct.pack_into(e + "hh", h, 252, 0, 1)
        struct.pack_into(e + "12f", h, 280, -pixdim[0], 0, 0, -3.0, 0, -pixdim[1], 0, -4.0, 0, 0, pixdim[2], 5.0)
    h[344:348] = b"n+1\x00"
    payload = np.asfortranarray(arr_xyz).astype(dt.newbyteorder(e)).tobytes(order="F")   # x fastest
    return bytes(h) + payload


def pattern(shape, dtype):
    idx = np.indices(shape)
    return sum(i * 7 ** k for k, i in enumerate(idx)).astype(dtype) % 120


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("endian", ["<", ">"])
@pytest.mark.parametrize("gz", [False, True])
def test_load_spec_built_nifti(tmp_path, dtype, endian, gz):
    a = pattern((5, 4, 3), dtype)
    raw = spec_nifti(a, endian, pixdim=(0.5, 0.75, 2.0))
    p = tmp_path / ("k.nii.gz" if gz else "k.nii")
    p.write_bytes(gzip.compress(raw) if gz else raw)
    arr, hdr = formats.load(p)
    assert arr.shape == (5, 4, 3) and arr.dtype == np.dtype(dtype)
    np.testing.assert_array_equal(arr, a)                      # arr[x, y, z] as medpy indexes it
    assert arr.flags.f_contiguous                               # a view of the x-fastest payload
    assert hdr.get_voxel_spacing() == (0.5, 0.75, 2.0)
    np.testing.assert_allclose(hdr.get_offset(), (3.0, 4.0, 5.0))   # RAS -> LPS flips x and y
    np.testing.assert_allclose(hdr.get_direction(), np.eye(3))


def test_load_applies_scaling_and_2d(tmp_path):
    a = pattern((6, 5), np.int16)
    p = tmp_path / "s.nii"
    p.write_bytes(spec_nifti(a, slope=0.5, inter=2.0, pixdim=(1.0, 1.0), sform=False))
    arr, hdr = formats.load(p)
    assert arr.shape == (6, 5)
    np.testing.assert_array_equal(arr, a * 0.5 + 2.0)
    with pytest.raises(FileNotFoundError):
        formats.load(tmp_path / "missing.nii.gz")
    (tmp_path / "bad.nii").write_bytes(b"\x00" * 400)
    with pytest.raises(ValueError):
        formats.load(tmp_path / "bad.nii")


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.float32, np.float64, np.bool_])
def test_save_writes_spec_layout(tmp_path, dtype):
    a = pattern((7, 3, 5), np.float64).astype(dtype)
    p = tmp_path / "w.nii.gz"
    formats.save(a, p)
    raw = gzip.decompress(p.read_bytes())
    stored = np.uint8 if dtype == np.bool_ else dtype
    assert struct.unpack_from("<i", raw, 0)[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack_from("<8h", raw, 40)[:4] == (3, 7, 3, 5)
    assert struct.unpack_from("<hh", raw, 70) == (CODES[stored], np.dtype(stored).itemsize * 8)
    assert struct.unpack_from("<3f", raw, 108) == (352.0, 1.0, 0.0)
    assert struct.unpack_from("<hh", raw, 252) == (1, 1)
    assert struct.unpack_from("<3f", raw, 256) == (0.0, 0.0, 1.0)          # identity LPS = 180 deg about z in RAS
    assert struct.unpack_from("<12f", raw, 280) == (-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0)
    assert len(raw) == 352 + a.size * np.dtype(stored).itemsize
    np.testing.assert_array_equal(np.frombuffer(raw, stored, offset=352).reshape(5, 3, 7), a.astype(stored).T)
    back, _ = formats.load(p)
    np.testing.assert_array_equal(back, a.astype(stored))
    first = p.read_bytes()
    formats.save(a, p)
    assert p.read_bytes() == first                                        # deterministic bytes (gzip mtime 0)


def test_header_round_trip(tmp_path):
    a = pattern((4, 5, 6), np.float32)
    hdr = formats.Header((0.7, 0.8, 2.5), (-10.0, 20.0, 30.0), np.eye(3))
    formats.save(a, tmp_path / "h.nii.gz", hdr)
    _, back = formats.load(tmp_path / "h.nii.gz")
    np.testing.assert_allclose(back.get_voxel_spacing(), (0.7, 0.8, 2.5), rtol=1e-6)
    np.testing.assert_allclose(back.get_offset(), (-10.0, 20.0, 30.0), rtol=1e-6)
    np.testing.assert_allclose(back.get_direction(), np.eye(3), atol=1e-6)


def test_cv2_formats_against_cv2(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    unc = rng.random((37, 53)).astype(np.float32)                       # [H, W] as test_2D.py:157 writes it
    cv2.imwrite(str(tmp_path / "u.tif"), unc)
    arr, _ = formats.load(tmp_path / "u.tif")
    assert arr.shape == (53, 37) and arr.dtype == np.float32            # medpy: [W, H] (ace.py:21-23)
    np.testing.assert_array_equal(arr, unc.T)
    formats.save(arr, tmp_path / "v.tif")
    np.testing.assert_array_equal(cv2.imread(str(tmp_path / "v.tif"), cv2.IMREAD_UNCHANGED), unc)
    seg = rng.integers(0, 19, (37, 53)).astype(np.uint8)
    formats.save(seg.T, tmp_path / "s.png")
    np.testing.assert_array_equal(cv2.imread(str(tmp_path / "s.png"), cv2.IMREAD_UNCHANGED), seg)


def make_results_dir(root: Path, pred_model="Dropout", n_classes=2):
    rng = np.random.default_rng(5)
    d = root / pred_model / "test_results" / "0" / "id"
    shape = (12, 11, 10)
    for sub in ("pred_seg", "pred_prob", "gt_seg", "pred_entropy", "aleatoric_uncertainty"):
        (d / sub).mkdir(parents=True, exist_ok=True)
    store = {}
    for image_id in ("case_b", "case_a"):
        for name in ("mean", "01", "02"):
            store[f"pred_seg/{image_id}_{name}"] = (rng.random(shape) < 0.3).astype(np.uint8)
        for r in range(2):
            store[f"gt_seg/{image_id}_{r:02d}"] = (rng.random(shape) < 0.3).astype(np.float64)
        p = rng.random((n_classes,) + shape)
        p /= p.sum(0)
        for c in range(n_classes):
            store[f"pred_prob/{image_id}_01_{c + 1:02d}"] = p[c]
        for unc in ("pred_entropy", "aleatoric_uncertainty"):
            store[f"{unc}/{image_id}"] = rng.random(shape)
    for k, v in store.items():
        formats.save(v, d / (k + ".nii.gz"))
    return d, store


def exp_version(root, pred_model="Dropout", **kw):
    v = dict(exp_path=root / pred_model / "test_results" / "0", version_params={"seed": 3}, pred_model=pred_model,
             n_classes=2, unc_types=["predictive_uncertainty", "aleatoric_uncertainty"], image_ending=".nii.gz",
             unc_ending=".nii.gz", n_reference_segs=2, datamodule_config=None, gt_unc_map_loading=None,
             pred_seg_loading=None)
    v.update(kw)
    return SimpleNamespace(**v)


def test_experiment_dataloader_paths_and_arrays(tmp_path):
    d, store = make_results_dir(tmp_path)
    dl = ExperimentDataloader(exp_version(tmp_path), "id")
    assert dl.image_ids == ["case_a", "case_b"] and dl.dataset_path == d
    assert dl.unc_path_dict == {"predictive_uncertainty": d / "pred_entropy",
                                "aleatoric_uncertainty": d / "aleatoric_uncertainty"}
    np.testing.assert_array_equal(dl.get_unc_map("case_a", "predictive_uncertainty"), store["pred_entropy/case_a"])
    np.testing.assert_array_equal(dl.get_mean_pred_seg("case_b"), store["pred_seg/case_b_mean"])
    assert len(dl.get_pred_segs("case_a")) == 3
    refs = dl.get_reference_segs("case_a")
    assert refs.shape == (2, 12, 11, 10)
    np.testing.assert_array_equal(dl.get_gt_unc_map("case_a"), np.var(refs, axis=0))
    assert dl.ref_seg_dir == d / "gt_seg" and dl.dataloader is None
    (d / "aggregated_predictive_uncertainty.json").write_text("{}")
    assert list(dl.get_aggregated_unc_files_dict()) == ["predictive_uncertainty"]


def test_experiment_dataloader_softmax_writes_pred_entropy(tmp_path):
    d, store = make_results_dir(tmp_path, pred_model="Softmax")
    import shutil

    shutil.rmtree(d / "pred_entropy")
    dl = ExperimentDataloader(exp_version(tmp_path, pred_model="Softmax", unc_types=["predictive_uncertainty"]), "id")
    want = 1 - np.maximum(store["pred_prob/case_a_01_01"], store["pred_prob/case_a_01_02"])
    np.testing.assert_array_equal(dl.get_unc_map("case_a", "predictive_uncertainty"), want)
    np.testing.assert_array_equal(dl.get_mean_pred_seg("case_a"), store["pred_seg/case_a_01"])   # Softmax -> _01
