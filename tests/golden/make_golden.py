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


def main():
    ref = ref_loader.load()
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


if __name__ == "__main__":
    main()
