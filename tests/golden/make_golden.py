"""Generate tests/golden/*.npz by running the UNMODIFIED reference (IML-DKFZ/values) in
place from /root/reference (through oracle/ref_loader.py stubs).  Run in the build
container only:   python tests/golden/make_golden.py

The vectors pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py) on the GPU box, where the reference itself cannot travel.
Inputs are stored next to the reference's outputs, so nothing is re-derived at test time.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def softmax_stack(gen, n, c, spatial, dtype, sharp=3.0, shared=False):
    if shared:  # samples mostly agree -> small MI (the cancellation regime)
        base = torch.randn(1, c, *spatial, generator=gen, dtype=torch.float64) * sharp
        logits = base + 0.3 * torch.randn(n, c, *spatial, generator=gen, dtype=torch.float64)
    else:
        logits = torch.randn(n, c, *spatial, generator=gen, dtype=torch.float64) * sharp
    return torch.softmax(logits, dim=1).to(dtype)


def c2_case(ref, name, x, ssn=False):
    d = ref.calculate_uncertainty(x, ssn=ssn)
    mean = torch.mean(x, dim=0)
    out = {
        "softmax": x.numpy(),
        "ssn": np.array(ssn),
        "pred_entropy": d["pred_entropy"].numpy(),
        "aleatoric_uncertainty": d["aleatoric_uncertainty"].numpy(),
        "epistemic_uncertainty": d["epistemic_uncertainty"].numpy(),
        # argmax exactly as the reference's savers take it (data_carrier_3D.py:253-256)
        "mean_argmax": np.argmax(np.mean(x.numpy(), axis=0), axis=0).astype(np.uint8),
        "mean_argmax_torch": torch.argmax(mean, dim=0).to(torch.uint8).numpy(),
        "sample_argmax": np.stack([np.argmax(x[i].numpy(), axis=0) for i in range(x.shape[0])]
                                  ).astype(np.uint8),
    }
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, {k: (v.shape, v.dtype) for k, v in out.items()})


class FakeExpDataloader:
    """Duck-typed stand-in for evaluation.experiment_dataloader.ExperimentDataloader: the
    reference's metric loops only call these getters (ncc.py:28-47, ace.py:89-133)."""

    def __init__(self, root, unc_maps, pred_segs, ref_segs, gt_unc):
        import types
        from pathlib import Path

        self.dataset_path = Path(root)
        self.exp_version = types.SimpleNamespace(exp_path=Path(root), unc_types=sorted(unc_maps),
                                                 pred_model="Dropout", version_name="v0")
        self.image_ids = sorted(pred_segs)
        self._unc, self._pred, self._ref, self._gt = unc_maps, pred_segs, ref_segs, gt_unc

    def get_unc_map(self, image_id, unc_type):
        return self._unc[unc_type][image_id]

    def get_mean_pred_seg(self, image_id):
        return self._pred[image_id]

    def get_pred_segs(self, image_id):
        return [self._pred[image_id]]

    def get_reference_segs(self, image_id):
        return self._ref[image_id]

    def get_gt_unc_map(self, image_id):
        return self._gt[image_id]


def k4_cases(ref):
    """f1 / f3: threshold finding, NCC and ACE through the reference's own loops."""
    import json
    import tempfile

    rng = np.random.default_rng(20261018)
    out = {}
    # ---- foreground quantile + np.quantile threshold (find_threshold.py:11-13, 63-68)
    seg = (rng.random((9, 10, 11)) < 0.07).astype(np.uint8)
    out["fg_seg"] = seg
    out["fg_quantile"] = np.array(ref.calculate_foreground_quantile_image(seg))
    maps64 = rng.random((3, 12, 11, 10)).astype(np.float32).astype(np.float64) / np.array([1.0, 2.0, 3.0]).reshape(3, 1, 1, 1)
    maps64[1, :3] = 0.0          # duplicates (uncovered remainder)
    maps32 = rng.random((4, 33, 20)).astype(np.float32)
    out["thr_maps64"], out["thr_maps32"] = maps64, maps32
    qs = [0.0, 0.25, 0.9312345, 0.98, 1.0, float(out["fg_quantile"])]
    out["thr_q"] = np.array(qs)
    with tempfile.TemporaryDirectory() as tmp:
        for j, q in enumerate(qs):
            qp = os.path.join(tmp, f"q{j}.json")
            with open(qp, "w") as f:
                json.dump({"Dropout": q}, f)
            out[f"thr64_{j}"] = np.array(ref.calculate_threshold_image(qp, maps64, "Dropout"))
            out[f"thr32_{j}"] = np.array(ref.calculate_threshold_image(qp, maps32, "Dropout"))
    # ---- NCC + ACE through the reference loops with a duck-typed data loader
    ids = ["img_a", "img_b", "img_c"]
    shape3, shape2 = (10, 9, 8), (14, 21)
    for tag, shape, dt, ignore in (("3d", shape3, np.float64, None), ("2d", shape2, np.float32, 255)):
        unc = {u: {} for u in ("aleatoric_uncertainty", "epistemic_uncertainty", "pred_entropy")}
        pred, refs, gt = {}, {}, {}
        for i, image_id in enumerate(ids):
            R = 1 + i
            pred[image_id] = (rng.random(shape) < 0.3).astype(np.uint8)
            r = np.stack([np.where(rng.random(shape) < 0.85, pred[image_id], 1 - pred[image_id])
                          for _ in range(R)]).astype(np.uint8)
            if ignore is not None:
                r[rng.random(r.shape) < 0.1] = ignore
            if i == 2 and tag == "3d":
                r = np.repeat(pred[image_id][None], R, 0)   # all raters agree: single-label quirk
            refs[image_id] = r
            gt[image_id] = rng.random(shape).astype(dt)
            for u in unc:
                m = (rng.random(shape) ** 2 * 0.7).astype(dt)
                # the 2D maps are loaded as (W, H) (ace.py:101-102)
                unc[u][image_id] = np.ascontiguousarray(m.T) if tag == "2d" else m
        params = {u: {"a": 3.0 + k, "b": -1.25 + 0.5 * k} for k, u in enumerate(sorted(unc))}
        with tempfile.TemporaryDirectory() as tmp:
            dl = FakeExpDataloader(tmp, unc, pred, refs, gt)
            with open(os.path.join(tmp, "platt_scale_params.json"), "w") as f:
                json.dump(params, f)
            ref.calibration_error(dl, ignore_value=ignore)
            with open(os.path.join(tmp, "calibration.json")) as f:
                calib = json.load(f)
            if tag == "3d":     # ncc.py needs equal shapes
                ref.ncc_main(dl)
                with open(os.path.join(tmp, "ambiguity_modeling.json")) as f:
                    ncc = json.load(f)
        for k, u in enumerate(sorted(unc)):
            out[f"{tag}_platt_{k}"] = np.array([params[u]["a"], params[u]["b"]])
            out[f"{tag}_ace_mean_{k}"] = np.array(calib["mean"][u]["metrics"]["ace"])
            if tag == "3d":
                out[f"{tag}_ncc_mean_{k}"] = np.array(ncc["mean"][u]["metrics"]["ncc"])
            for i, image_id in enumerate(ids):
                out[f"{tag}_unc_{k}_{i}"] = unc[u][image_id]
                out[f"{tag}_ace_{k}_{i}"] = np.array(calib[image_id][u]["metrics"]["ace"])
                if tag == "3d":
                    out[f"{tag}_ncc_{k}_{i}"] = np.array(ncc[image_id][u]["metrics"]["ncc"])
        for i, image_id in enumerate(ids):
            out[f"{tag}_pred_{i}"], out[f"{tag}_refs_{i}"], out[f"{tag}_gt_{i}"] = pred[image_id], refs[image_id], gt[image_id]
        out[f"{tag}_ignore"] = np.array(-1 if ignore is None else ignore)
    out["unc_names"] = np.array(sorted(unc))
    # ---- calib_stats directly (ace.py:49-81) incl. fp32 confidences
    conf = rng.random(5000)
    corr = (rng.random(5000) < conf).astype(int)
    disc, tot, nz = ref.calib_stats(corr, conf)
    out["cs_conf"], out["cs_correct"] = conf, corr
    out["cs_disc"], out["cs_total"], out["cs_nonzero"] = disc, tot, np.array(nz)
    out["cs_ace"] = np.array(ref.calc_ace(corr, conf))
    # ncc on an fp32 pair
    a32, b32 = rng.random((17, 19)).astype(np.float32), rng.random((17, 19)).astype(np.float32)
    out["ncc32_a"], out["ncc32_b"], out["ncc32"] = a32, b32, np.array(ref.compute_ncc(a32, b32))
    np.savez_compressed(os.path.join(HERE, "k4_stats.npz"), **out)
    print("k4_stats", len(out), "arrays")


def formats_cases(ref, carrier):
    """the reference's own file-level hand-off, run unmodified on the stitch case above
    with medpy.io.save / load replaced by values_b200.formats.save / load (medpy is absent here):
      * DataCarrier3D.save_data -> every array it hands to save(), keyed by relative path;
      * ExperimentDataloader + aggregate_uncertainties on the directory it wrote -> the
        aggregated_<unc>.json contents (hydra.utils.instantiate replaced by a direct call of the
        reference's own aggregation functions, jsbeautifier by the identity)."""
    import importlib
    import json
    import tempfile
    from pathlib import Path

    from values_b200 import formats  # host I/O only; no kernel runs here

    dc = ref.modules["data_carrier_3D"]
    agg = ref.modules["aggregate_uncertainties"]
    edl = importlib.import_module("evaluation.experiment_dataloader")
    ev = importlib.import_module("evaluation.experiment_version")
    recorded = {}
    with tempfile.TemporaryDirectory() as tmp:
        def rec_save(arr, path, hdr=False):
            recorded[os.path.relpath(path, tmp)] = np.asarray(arr)
            formats.save(arr, path, hdr)

        dc.save = rec_save
        carrier.save_data(root_dir=tmp, exp_name="Dropout", version=0, org_data_path=None, test_split="id")
        edl.load, edl.save = formats.load, formats.save
        agg.load = formats.load

        def instantiate(cfg, **kwargs):
            cfg = dict(cfg)
            fn = getattr(agg, cfg.pop("_target_").rsplit(".", 1)[1])
            return fn(**cfg, **kwargs)

        agg.hydra.utils.instantiate = instantiate
        agg.jsbeautifier.beautify = lambda text, opts=None: text
        target = "evaluation.uncertainty_aggregation.aggregate_uncertainties."
        aggregations = {
            "patch_level": {"_target_": target + "patch_level_aggregation", "patch_size": 10},
            "image_level": {"_target_": target + "image_level_aggregation"},
            "threshold": {"_target_": target + "threshold_aggregation", "threshold": 0.3},
        }
        unc_types = ["predictive_uncertainty", "aleatoric_uncertainty", "epistemic_uncertainty"]
        version = ev.ExperimentVersion(
            base_path=Path(tmp), naming_scheme_version="{version}", pred_model="Dropout",
            image_ending=".nii.gz", unc_ending=".nii.gz", unc_types=unc_types, aggregations=aggregations,
            n_reference_segs=1, version=0, seed=123)
        loader = edl.ExperimentDataloader(version, "id")
        agg.aggregate_uncertainties(loader, aggregations)
        aggregated = {}
        for unc in unc_types:
            with open(loader.dataset_path / f"aggregated_{unc}.json") as f:
                aggregated[unc] = json.load(f)
        meta = {
            "image_ids": loader.image_ids,
            "unc_dirs": {k: os.path.relpath(v, tmp) for k, v in loader.unc_path_dict.items()},
            "pred_seg_files": sorted(os.path.relpath(p, tmp) for p in loader.get_pred_seg_paths(loader.image_ids[0])),
            "gt_unc_map_sum": float(np.sum(loader.get_gt_unc_map(loader.image_ids[0]))),
            "aggregated": aggregated, "aggregations": aggregations,
        }
    np.savez_compressed(os.path.join(HERE, "save_data_3d.npz"),
                        **{k.replace(os.sep, "|"): v for k, v in recorded.items()})
    with open(os.path.join(HERE, "save_data_3d.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("save_data_3d", len(recorded), "files")


A5_CASES = [((20, 17, 33), 8, 0.5), ((64, 64, 64), 16, 1), ((70, 45, 33), 16, 0.75), ((40, 40, 40), 10, 0.25),
            ((33, 20, 17), 8, 0.3), ((16, 16, 16), 16, 1), ((15, 40, 40), 16, 0.5), ((128, 128, 128), 64, 0.5),
            ((256, 256, 256), 64, 0.5), ((256, 256, 256), 64, 1)]


def patch_grid_cases():
    import json

    """SURVEY 8 row a5: the crop-index loop of the sliding-window path, produced by the reference's own
    get_val_test_data_samples (lidc_idri_datamodule_3D.py:719-736; toy_datamodule_3D.py:637-654 must
    agree) on temporary .npy volumes -> tests/golden/patch_grid.json."""
    import tempfile
    from pathlib import Path

    lidc, toy = ref_loader.load_datamodules()
    out = []
    for shape, p, overlap in A5_CASES:
        crops = {}
        for mod, is_toy in ((lidc, False), (toy, True)):
            with tempfile.TemporaryDirectory() as d:
                sub = "Tr" if is_toy else ""
                os.makedirs(Path(d) / f"images{sub}")
                os.makedirs(Path(d) / f"labels{sub}")
                np.save(Path(d) / f"images{sub}" / "vol.npy", np.zeros(shape, np.uint8))
                samples = mod.get_val_test_data_samples(d, subject_ids=["vol.npy"], num_raters=1,
                                                        patch_size=p, patch_overlap=overlap)
                crops[is_toy] = [s["crop_idx"] for s in samples]
        assert crops[False] == crops[True]
        out.append({"shape": list(shape), "patch_size": p, "patch_overlap": overlap,
                    "crops": [[list(ax) for ax in c] for c in crops[False]]})
    with open(os.path.join(HERE, "patch_grid.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("patch_grid", len(out), "cases,", sum(len(c["crops"]) for c in out), "crops")


def process_output_2d_case():
    """The 2-D caller: the reference's own Tester.process_output (uncertainty_modeling/test_2D.py:204-254) and
    the calculate_test_metrics / calculate_ged it calls, run UNMODIFIED on one synthetic test batch with a bare
    object standing in for the Tester (device, ignore_index, results_dict; the two file writers record their
    arguments instead of writing).  torchmetrics is absent: `dice` is bound to the oracle's dice_micro in both
    modules, so the Dice numbers are the oracle's while everything else (zero channel, relabelling, slicing,
    means, which maps go where) is the reference's."""
    import importlib
    import types

    from oracle import values_oracle as vo

    ref_loader.load()
    t2d = importlib.import_module("uncertainty_modeling.test_2D")
    t3d = importlib.import_module("uncertainty_modeling.test_3D")

    def micro(preds, target, ignore_index=None):
        p = torch.argmax(preds, dim=1) if preds.dim() == target.dim() + 1 else preds
        n_cls = preds.shape[1] if preds.dim() == target.dim() + 1 else int(max(p.max(), target.max())) + 1
        return torch.tensor(vo.dice_micro(p.numpy(), target.numpy(), n_cls, ignore_index), dtype=torch.float64)

    gen = torch.Generator().manual_seed(77)
    out = {}
    for name, n, b, c, hw, ssn in (("mc", 4, 3, 5, (24, 40), False), ("ssn", 3, 2, 4, (16, 24), True), ("single", 1, 2, 3, (12, 20), False)):
        base = 2.0 * torch.randn(1, b, c, *hw, generator=gen)
        sm = torch.softmax(base + torch.randn(n, b, c, *hw, generator=gen), dim=2)
        gt = torch.argmax(base[0], dim=1, keepdim=True)
        flip = torch.rand(gt.shape, generator=gen) < 0.15
        gt = torch.where(flip, (gt + 1) % c, gt)
        gt[torch.rand(gt.shape, generator=gen) < 0.1] = 255
        saved = {"pred": {}, "unc": {}}
        fake = types.SimpleNamespace(device="cpu", ignore_index=255, results_dict={})
        fake.calculate_test_metrics = types.MethodType(t2d.Tester.calculate_test_metrics, fake)
        fake.save_prediction = lambda image_id, image_preds, mean_pred, ign, _s=saved: _s["pred"].__setitem__(
            image_id, (torch.argmax(mean_pred, dim=0).numpy().astype(np.uint8),
                       torch.argmax(image_preds, dim=1).numpy().astype(np.uint8), ign[..., 0].astype(bool)))
        fake.save_uncertainty = lambda image_id, d, _s=saved: _s["unc"].__setitem__(
            image_id, {k: v.numpy() for k, v in d.items()})
        keep = (t2d.dice, t3d.dice)
        t2d.dice = t3d.dice = micro
        try:
            t2d.Tester.process_output(fake, {"softmax_pred": sm.clone(), "gt": gt.clone(),
                                             "image_id": [f"{name}_{i}" for i in range(b)],
                                             "dataset": ["synthetic"] * b}, ssn)
        finally:
            t2d.dice, t3d.dice = keep
        out[name + "/softmax_pred"], out[name + "/gt"] = sm.numpy(), gt.numpy().astype(np.int64)
        out[name + "/ssn"] = np.asarray(ssn)
        for i in range(b):
            key = f"{name}_{i}"
            for m, v in fake.results_dict[key]["metrics"].items():
                out[f"{name}/{i}/metrics/{m}"] = np.asarray(v, dtype=np.float64)
            out[f"{name}/{i}/mean_argmax"], out[f"{name}/{i}/sample_argmax"], out[f"{name}/{i}/ignore"] = saved["pred"][key]
            for k, v in saved["unc"][key].items():
                out[f"{name}/{i}/unc/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "process_output_2d.npz"), **out)
    print("process_output_2d", len(out), "arrays")


def main():
    ref = ref_loader.load()
    if "--only-2d" in sys.argv:
        process_output_2d_case()
        return
    if "--only-k4" in sys.argv:
        k4_cases(ref)
        return
    if "--only-a5" in sys.argv:
        patch_grid_cases()
        return
    gen = torch.Generator().manual_seed(20261017)

    # ---- C2: fp32 2D path incl. the appended all-zero channel (test_2D.py:208-218)
    x = softmax_stack(gen, 4, 5, (9, 13), torch.float32)
    x = torch.cat([x, torch.zeros(4, 1, 9, 13)], dim=1)
    c2_case(ref, "c2_f32_2d_zero_channel", x)
    # ---- C2: fp32 3D, cfg5-like class/sample counts, with exact zeros / one-hots planted
    x = softmax_stack(gen, 16, 4, (6, 7, 8), torch.float32)
    x[:, :, 0, 0, :] = 0.0
    x[:, 0, 0, 0, :] = 1.0  # one-hot -> NaN-skip on every other class
    x[3, :, 1, 1, 1] = torch.tensor([0.25, 0.25, 0.25, 0.25])
    c2_case(ref, "c2_f32_3d_n16_c4", x)
    # ---- C2: low-MI regime, C=2 N=5 (cfg1/2 shape family)
    x = softmax_stack(gen, 5, 2, (8, 8, 8), torch.float32, shared=True)
    c2_case(ref, "c2_f32_3d_lowmi", x)
    # ---- C2: fp64 as the 3D path feeds it, raw sums > 1 where patches overlap (H5)
    x = softmax_stack(gen, 5, 2, (8, 8, 8), torch.float64)
    x[:, :, 4:, :, :] *= 2.0
    x[:, :, :, :, 7] = 0.0  # uncovered remainder -> all zero
    c2_case(ref, "c2_f64_3d_rawsum", x)
    # ---- C2: ssn swap, C=25 like GTA (24 + zero channel)
    x = softmax_stack(gen, 10, 24, (5, 11), torch.float32)
    x = torch.cat([x, torch.zeros(10, 1, 5, 11)], dim=1)
    c2_case(ref, "c2_f32_2d_c25_ssn", x, ssn=True)
    # ---- C2: negative / >1 / nan inputs (log -> NaN is skipped, +inf is not)
    x = softmax_stack(gen, 3, 3, (4, 6), torch.float32)
    x[0, 0, 0, 0] = -0.25
    x[1, 1, 0, 1] = float("nan")
    x[2, 2, 0, 2] = 1.5
    c2_case(ref, "c2_f32_2d_pathological", x)

    # ---- 1 - MSR
    x = softmax_stack(gen, 1, 7, (5, 6, 7), torch.float32)[0]
    d = ref.calculate_one_minus_msr(x)
    np.savez_compressed(os.path.join(HERE, "msr_f32.npz"), softmax=x.numpy(),
                        pred_entropy=d["pred_entropy"].numpy())
    x = softmax_stack(gen, 1, 2, (6, 6, 6), torch.float64)[0]
    d = ref.calculate_one_minus_msr(x)
    np.savez_compressed(os.path.join(HERE, "msr_f64.npz"), softmax=x.numpy(),
                        pred_entropy=d["pred_entropy"].numpy())

    # ---- C3 aggregations
    rng = np.random.default_rng(7)
    agg = {}
    maps = {
        "m3d_f64": rng.random((16, 15, 14)).astype(np.float32).astype(np.float64),
        "m2d_f32": rng.random((31, 20)).astype(np.float32),
        "m3d_zero": np.zeros((12, 12, 12)),
    }
    # planted blocks: an EARLIER window 5e-6 (relative) below the max wins the bbox,
    # a window 5e-5 below does not (np.isclose rule, aggregate_uncertainties.py:20-23)
    planted = np.zeros((24, 24, 24))
    planted[2:6, 2:6, 2:6] = 1.0 - 5e-5
    planted[10:14, 2:6, 2:6] = 1.0 - 5e-6
    planted[18:22, 18:22, 18:22] = 1.0
    maps["m3d_planted"] = planted
    cases = [("m3d_f64", 4), ("m3d_f64", 10), ("m2d_f32", 10), ("m2d_f32", [3, 5]),
             ("m3d_zero", 10), ("m3d_planted", 4), ("m3d_f64", [16, 15, 14])]
    for i, (mname, p) in enumerate(cases):
        for mean in (False, True):
            r = ref.patch_level_aggregation(maps[mname], p, mean=mean)
            key = f"patch_{i}_{int(mean)}"
            agg[key + "_map"] = np.array(mname)
            agg[key + "_patch"] = np.atleast_1d(np.array(p))
            agg[key + "_score"] = np.array(r["max_score"], dtype=np.float64)
            agg[key + "_bbox"] = np.array(r["bounding_box"], dtype=np.int64)
    agg["n_patch_cases"] = np.array(len(cases))
    for mname, m in maps.items():
        agg["map_" + mname] = m
        agg["image_sum_" + mname] = np.array(ref.image_level_aggregation(m)["max_score"])
        agg["image_mean_" + mname] = np.array(ref.image_level_aggregation(m, mean=True))
        for j, thr in enumerate([0.0, 0.5, 0.9, 2.0]):
            for mean in (True, False):
                r = ref.threshold_aggregation(m, threshold=thr, mean=mean)
                agg[f"thr_{mname}_{j}_{int(mean)}"] = np.array(float(r["max_score"]))
    agg["thresholds"] = np.array([0.0, 0.5, 0.9, 2.0])
    np.savez_compressed(os.path.join(HERE, "c3_aggregations.npz"), **agg)
    print("c3_aggregations", len(agg), "arrays")

    patch_grid_cases()

    # ---- stitching through the reference DataCarrier3D (C=2 hardcoded there)
    from oracle import values_oracle as vo

    shape, p, overlap, n_pred = (20, 18, 17), 8, 0.5, 3
    crops = vo.patch_grid(shape, p, overlap)
    carrier = ref.DataCarrier3D()
    patches = softmax_stack(gen, n_pred * len(crops), 2, (p, p, p), torch.float64)
    patches = patches.reshape(n_pred, len(crops), 2, p, p, p)
    bs = 5
    for pred_idx in range(n_pred):
        for s in range(0, len(crops), bs):
            idx = list(range(s, min(s + bs, len(crops))))
            batch = {
                "image_paths": ["vol_a.npy"] * len(idx),
                "label_paths": [["lab_a.npy"]] * len(idx),
                "org_image_size": [shape] * len(idx),
                "crop_idx": [crops[i] for i in idx],
                "data": torch.zeros(len(idx), 1, p, p, p),
                "seg": torch.zeros(1, len(idx), p, p, p, dtype=torch.int32),
            }
            carrier.concat_data(batch, patches[pred_idx, idx], n_pred=n_pred, pred_idx=pred_idx)
    ref.caculcate_uncertainty_multiple_pred(carrier)
    v = carrier.data["vol_a.npy"]
    cnt = np.clip(v["num_predictions"], 1, None)
    sm = v["softmax_pred"] / cnt
    np.savez_compressed(
        os.path.join(HERE, "stitch_3d.npz"),
        shape=np.array(shape), patch=np.array(p), overlap=np.array(overlap),
        crops=np.array(crops, dtype=np.int64), patches=patches.numpy(),
        softmax_sum=v["softmax_pred"], num_predictions=v["num_predictions"],
        pred_entropy=v["pred_entropy"].numpy(),
        aleatoric_uncertainty=v["aleatoric_uncertainty"].numpy(),
        epistemic_uncertainty=v["epistemic_uncertainty"].numpy(),
        pred_entropy_saved=np.asarray(v["pred_entropy"] / cnt[0]),
        mean_seg=np.argmax(np.mean(sm, axis=0), axis=0).astype(np.uint8),
    )
    print("stitch_3d", len(crops), "patches")
    formats_cases(ref, carrier)
    k4_cases(ref)


if __name__ == "__main__":
    main()
