"""CPU: the oracle restatement against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Bit-exact for C2 maps / argmax / bbox / stitch sums;
1e-12 relative for patch scores (reference = fp64 FFT, oracle = direct sums)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import values_oracle as vo

C2_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "c2_*.npz")))


def test_golden_files_present():
    assert len(C2_CASES) >= 6


@pytest.mark.parametrize("name", C2_CASES)
def test_c2_maps_bit_exact(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = torch.from_numpy(g["softmax"])
    d = vo.calculate_uncertainty(x, ssn=bool(g["ssn"]))
    for k in ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty"):
        assert d[k].dtype == torch.float32
        np.testing.assert_array_equal(d[k].numpy(), g[k], err_msg=k)
    np.testing.assert_array_equal(vo.mean_argmax(x).numpy(), g["mean_argmax_torch"])
    np.testing.assert_array_equal(vo.sample_argmax(x).numpy(), g["sample_argmax"])


@pytest.mark.parametrize("name", ["msr_f32", "msr_f64"])
def test_msr(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = vo.calculate_one_minus_msr(torch.from_numpy(g["softmax"]))
    np.testing.assert_array_equal(d["pred_entropy"].numpy(), g["pred_entropy"])


def test_c3_aggregations(golden_dir):
    g = np.load(os.path.join(golden_dir, "c3_aggregations.npz"))
    for i in range(int(g["n_patch_cases"])):
        for mean in (0, 1):
            key = f"patch_{i}_{mean}"
            m = g["map_" + str(g[key + "_map"])]
            p = g[key + "_patch"].tolist()
            p = p[0] if len(p) == 1 else p
            for method in ("direct", "fft"):
                r = vo.patch_level_aggregation(m, p, mean=bool(mean), method=method)
                # fp32 images: scipy runs the image FFT in complex64 -> ~1e-8 noise in
                # the reference itself; fp64 images: ~1e-13
                rtol = 1e-12 if (m.dtype == np.float64 or method == "fft") else 1e-6
                np.testing.assert_allclose(r["max_score"], float(g[key + "_score"]),
                                           rtol=rtol, atol=1e-300)
                assert [list(b) for b in r["bounding_box"]] == g[key + "_bbox"].tolist(), (key, method)
    for mname in ("m3d_f64", "m2d_f32", "m3d_zero", "m3d_planted"):
        m = g["map_" + mname]
        assert vo.image_level_aggregation(m)["max_score"] == float(g["image_sum_" + mname])
        assert vo.image_level_aggregation(m, mean=True) == float(g["image_mean_" + mname])
        for j, thr in enumerate(g["thresholds"].tolist()):
            for mean in (1, 0):
                r = vo.threshold_aggregation(m, threshold=thr, mean=bool(mean))
                assert float(r["max_score"]) == float(g[f"thr_{mname}_{j}_{mean}"])
                assert r["threshold"] == thr


def test_planted_isclose_rule(golden_dir):
    g = np.load(os.path.join(golden_dir, "c3_aggregations.npz"))
    r = vo.patch_level_aggregation(g["map_m3d_planted"], 4)
    # the block 5e-6 below the max (earlier in C order) wins; the one 5e-5 below does not
    assert r["bounding_box"] == [(10, 14), (2, 6), (2, 6)]
    assert r["max_score"] == 64.0


def test_threshold_errors():
    img = np.ones((3, 3))
    with pytest.raises(Exception, match="A threshold needs to be provided"):
        vo.threshold_aggregation(img)
    with pytest.raises(ValueError):
        vo.patch_level_aggregation(img, 4)


def test_stitch(golden_dir):
    g = np.load(os.path.join(golden_dir, "stitch_3d.npz"))
    shape, p = tuple(g["shape"].tolist()), int(g["patch"])
    crops = vo.patch_grid(shape, p, float(g["overlap"]))
    assert np.array(crops).tolist() == g["crops"].tolist()
    patches = torch.from_numpy(g["patches"])
    st = vo.StitchOracle()
    n_pred = patches.shape[0]
    for pred_idx in range(n_pred):
        for s in range(0, len(crops), 7):  # different batching than the generator
            idx = list(range(s, min(s + 7, len(crops))))
            batch = {"image_paths": ["v"] * len(idx), "org_image_size": [shape] * len(idx),
                     "crop_idx": [crops[i] for i in idx]}
            st.concat_data(batch, patches[pred_idx, idx], n_pred=n_pred, pred_idx=pred_idx)
    v = st.data["v"]
    np.testing.assert_array_equal(v["softmax_pred"], g["softmax_sum"])
    np.testing.assert_array_equal(v["num_predictions"], g["num_predictions"])
    # remainder never covered (20 = 8 + 3*4 covered fully; 18 -> last 2 uncovered; 17 -> last 1)
    assert v["num_predictions"][:, :, 16:, :].max() == 0
    assert v["num_predictions"][:, :, :, 16:].max() == 0
    d = vo.calculate_uncertainty(torch.from_numpy(v["softmax_pred"]))
    for k in ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty"):
        np.testing.assert_array_equal(d[k].numpy(), g[k])
    np.testing.assert_array_equal(
        vo.normalize_map(d["pred_entropy"], v["num_predictions"]), g["pred_entropy_saved"])
    np.testing.assert_array_equal(
        vo.stitched_argmax(v["softmax_pred"], v["num_predictions"]), g["mean_seg"])


def test_known_answers():
    # uniform p = 1/C -> PE = EE = log C, MI = 0
    x = torch.full((4, 5, 3, 3), 0.2, dtype=torch.float32)
    d = vo.calculate_uncertainty(x)
    np.testing.assert_allclose(d["pred_entropy"].numpy(), np.log(5.0), rtol=1e-6)
    np.testing.assert_allclose(d["epistemic_uncertainty"].numpy(), 0.0, atol=2e-7)
    # N one-hot samples cycling over K classes -> PE = log K, EE = 0, MI = log K
    x = torch.zeros(4, 4, 2, 2)
    for n in range(4):
        x[n, n] = 1.0
    d = vo.calculate_uncertainty(x)
    np.testing.assert_allclose(d["pred_entropy"].numpy(), np.log(4.0), rtol=1e-6)
    np.testing.assert_array_equal(d["aleatoric_uncertainty"].numpy(), 0.0)
    # stitch count pattern p=8 stride 4 on a length-12 axis: 1 1 1 1 2 2 2 2 1 1 1 1
    st = vo.StitchOracle()
    crops = vo.patch_grid((12, 8, 8), 8, 0.5)
    assert len(crops) == 2
    for c in crops:
        st.concat_data({"image_paths": ["a"], "org_image_size": [(12, 8, 8)], "crop_idx": [c]},
                       torch.ones(1, 2, 8, 8, 8))
    assert st.data["a"]["num_predictions"][0, :, 0, 0].tolist() == [1] * 4 + [2] * 4 + [1] * 4


# ------------------------------------------------------------------ f1 / f3 (k4_stats.npz)
def test_k4_threshold_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "k4_stats.npz"))
    assert vo.calculate_foreground_quantile_image(g["fg_seg"]) == float(g["fg_quantile"])
    for j, q in enumerate(g["thr_q"].tolist()):
        assert vo.quantile_threshold(g["thr_maps64"], q) == g[f"thr64_{j}"]
        r32 = vo.quantile_threshold(g["thr_maps32"], q)
        assert r32.dtype == np.float32 and r32 == g[f"thr32_{j}"]


def test_k4_ncc_ace_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "k4_stats.npz"))
    for tag in ("3d", "2d"):
        ignore = int(g[f"{tag}_ignore"])
        ignore = None if ignore < 0 else ignore
        for k in range(3):
            a, b = g[f"{tag}_platt_{k}"].tolist()
            aces = []
            for i in range(3):
                unc, pred, refs = g[f"{tag}_unc_{k}_{i}"], g[f"{tag}_pred_{i}"], g[f"{tag}_refs_{i}"]
                ace = vo.calibration_error_image(unc, pred, refs, a, b, ignore)
                assert ace == float(g[f"{tag}_ace_{k}_{i}"])
                aces.append(ace)
                if tag == "3d":
                    assert vo.compute_ncc(g[f"{tag}_gt_{i}"], unc) == float(g[f"{tag}_ncc_{k}_{i}"])
            assert np.mean(np.array(aces)) == float(g[f"{tag}_ace_mean_{k}"])
    disc, tot, nz = vo.calib_stats(g["cs_correct"], g["cs_conf"])
    np.testing.assert_array_equal(disc, g["cs_disc"])
    np.testing.assert_array_equal(tot, g["cs_total"])
    assert nz == int(g["cs_nonzero"]) and vo.calc_ace(g["cs_correct"], g["cs_conf"]) == float(g["cs_ace"])
    assert vo.compute_ncc(g["ncc32_a"], g["ncc32_b"]) == g["ncc32"]


def test_a5_patch_grid_golden(golden_dir):
    """the fixture holds the crop tuples the reference's own datamodule loop produced
    (tests/golden/make_golden.py::patch_grid_cases); the oracle and the product's host-side
    patch_grid (pure integer arithmetic, no device) must both reproduce them, in order."""
    import json

    from values_b200.stitching import patch_grid

    cases = json.load(open(os.path.join(golden_dir, "patch_grid.json")))
    assert len(cases) >= 8
    for c in cases:
        want = [tuple(tuple(ax) for ax in crop) for crop in c["crops"]]
        assert vo.patch_grid(c["shape"], c["patch_size"], c["patch_overlap"]) == want
        assert patch_grid(c["shape"], c["patch_size"], c["patch_overlap"]) == want
