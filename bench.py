#!/usr/bin/env python
"""bench.py -- throughput of the ValUES C2+C3 uncertainty hot path on B200.

One "step" = one pass of the hot path over one pool of synthetic inputs per rank: for the softmax-stack
workloads (cfg1 / cfg2 / cfg4 / cfg5) K1 (PE / EE / MI + arg-max + image-level and threshold sums) and
K2b (patch max) writing the per-image score table; for cfg3 the stitch accumulator K3, K1 on the fp64
raw sums and the save-path normalisation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5]

The line carries a BURST measurement (exactly K timed steps -- `value`, `ms_per_step`, `roofline`) and
a SUSTAINED one (>= 1 s of back-to-back steps, where the 1000 W power cap engages -- `sustained`), each
with the SM clocks and throttle reasons sampled inside its own timed region.

Under torchrun (N > 1) every rank processes its own pool (volume-sharded, weak scaling); the only
collective is ONE all_gather of the per-image score tables of the whole timed region.
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU implementation of the
path on the host cores, on whole volumes of the same workload in the workload's dtype: the UNMODIFIED
reference functions imported from baseline/_ref (offline pip install, git-ignored, travels with the
snapshot) or, if that is absent, the oracle port (oracle/values_oracle.py).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[4] per-volume shape (the config BASELINE.md quotes the 70% target on)
    "cfg5": dict(name="cfg5: 128^3 volumes, N=16 samples, C=4 classes, fp32; PE/EE/MI + arg-max + "
                      "image-level + threshold + patch-level(10) aggregation",
                 N=16, C=4, spatial=(128, 128, 128), dtype="f32", pool=32, e2e_pool=4, patch=10, thr=True, cfg=5),
    # BASELINE.json configs[0]: the toy 3-D set (Case_1 test split: 20 volumes), fp64 as the reference's 3-D
    # path feeds calculate_uncertainty (test_3D.py:532), image-level aggregation only
    "cfg1": dict(name="cfg1: toy 3D 64^3 volumes, N=5 (MC dropout), C=2, fp64; PE/EE/MI + arg-max + "
                      "image-level aggregation (8 passes over the 20-volume test split per step)",
                 N=5, C=2, spatial=(64, 64, 64), dtype="f64", pool=160, e2e_pool=40, patch=None, thr=False, cfg=1),
    # BASELINE.json configs[1]
    "cfg2": dict(name="cfg2: LIDC 64^3 patches, N=5, C=2, fp32; patch-level(10) + threshold aggregation",
                 N=5, C=2, spatial=(64, 64, 64), dtype="f32", pool=1024, e2e_pool=256, patch=10, thr=True, cfg=2),
    # BASELINE.json configs[2]: sliding-window stitching of one 256^3 volume (patch 64, overlap 0.5 -> 343
    # patches, TTA N=8, C=2, fp32 patches -> fp64 raw sums as DataCarrier3D keeps them), MI maps on the raw
    # sums (test_3D.py:528-534), then the save path's / clip(count, 1) (data_carrier_3D.py:326-363)
    "cfg3": dict(name="cfg3: stitch 343 patches 64^3 (overlap 0.5) into a 256^3 volume, N=8 (TTA), C=2, fp32 "
                      "patches -> fp64 sums; K1 fp64 on the raw sums; normalised maps",
                 N=8, C=2, spatial=(256, 256, 256), dtype="f64", patch_dtype="f32", pool=1, e2e_pool=1,
                 stitch=dict(patch=64, overlap=0.5), patch=None, thr=False, cfg=3),
    # the same with the Gaussian patch weight BASELINE.json's configs[2] / north_star name (the reference itself
    # accumulates with uniform weights, SURVEY D1: its CPU arm runs the uniform concat_data): separable
    # factors in the stitch kernel's shared memory, sums normalised by the weight sum
    "cfg3gauss": dict(name="cfg3 (Gaussian-weighted): stitch 343 patches 64^3 (overlap 0.5) into a 256^3 volume with "
                           "the separable Gaussian importance map, N=8 (TTA), C=2, fp32 patches -> fp64 weighted "
                           "sums + weight sum; K1 fp64 on the raw sums; maps normalised by the weight sum",
                      N=8, C=2, spatial=(256, 256, 256), dtype="f64", patch_dtype="f32", pool=1, e2e_pool=1,
                      stitch=dict(patch=64, overlap=0.5, weight="gaussian"), patch=None, thr=False, cfg=3),
    # BASELINE.json configs[3] (19 classes + the zero channel test_2D appends)
    "cfg4": dict(name="cfg4: 1024x2048 images, N=10, C=19+1, fp32; all C3 aggregations",
                 N=10, C=20, spatial=(1024, 2048), dtype="f32", pool=6, e2e_pool=2, patch=10, thr=True, cfg=4,
                 batch_sweep=(1, 2, 4, 6, 12)),
    # not BASELINE configs: shapes whose rows are not 16-byte aligned (every kernel of the step takes its
    # unaligned form: K1 element-strided ring, K2b pitched copy) and the reference-true 2-D shape of SURVEY 8d
    "cfg5odd": dict(name="cfg5 on 127^3 volumes (rows at every 16-byte phase), N=16, C=4, fp32; same aggregations",
                    N=16, C=4, spatial=(127, 127, 127), dtype="f32", pool=32, e2e_pool=4, patch=10, thr=True, cfg=5),
    "cfg4true": dict(name="cfg4 at the reference-true GTA shape 256x478, N=10, C=24+1, fp32; all C3 aggregations",
                     N=10, C=25, spatial=(256, 478), dtype="f32", pool=48, e2e_pool=8, patch=10, thr=True, cfg=4),
    "cfg4bf16": dict(name="cfg4: 1024x2048 images, N=10, C=19+1, bf16; all C3 aggregations",
                     N=10, C=20, spatial=(1024, 2048), dtype="bf16", pool=6, e2e_pool=2, patch=10, thr=True, cfg=4,
                     batch_sweep=(1, 2, 4, 6, 12)),
}
DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f64": torch.float64}
ELEM = {"f32": 4, "bf16": 2, "f64": 8}
METRIC, UNIT = "uncertainty_voxels_per_sec", "voxels/s"
MAPS = ("pred_entropy", "aleatoric_uncertainty", "epistemic_uncertainty")
SUSTAINED_S = 1.1   # length of the second timed region


def measured_peaks():
    """(burst GB/s, sustained GB/s or None, source): the driver-written copy bandwidth of this pool's B200s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        sus = next((float(j[k]) for k in ("hbm_gbs_sustained", "hbm_sustained_gbs", "hbm_gbs_long") if k in j), None)
        return float(j["hbm_gbs"]), sus, "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, None, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_peak_gbs():
    burst, _, src = measured_peaks()
    return burst, src


class ClockSampler:
    """SM clocks and throttle reasons sampled DURING the timed region (NVML, ~2 ms period, in a
    thread of this process; nvidia-smi -lms cannot start fast enough for a 100 ms region)."""

    HW_SLOWDOWN, SW_POWER_CAP = 0x8, 0x4
    SW_THERMAL, HW_THERMAL = 0x20, 0x40

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.thread = index, [], False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first calls resolve the NVML entry points (slow): do that before the timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        self.rows, self.stop_flag = [], False
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = {self.HW_SLOWDOWN: "hw_slowdown", self.HW_THERMAL: "hw_thermal_slowdown",
                 self.SW_THERMAL: "sw_thermal_slowdown", self.SW_POWER_CAP: "sw_power_cap"}
        reasons = sorted({n for _, r in self.rows for bit, n in names.items() if r & bit})
        return {"sm_mhz": statistics.median(sm for sm, _ in self.rows), "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(self.rows)}


def make_stack(gen, n_vol, wl, device, dtype):
    """Synthetic softmax stacks [n_vol, N, C, *S]: logits ~ N(0, 3^2), softmax over classes."""
    shape = (wl["N"], wl["C"]) + tuple(wl["spatial"])
    out = torch.empty((n_vol,) + shape, dtype=dtype, device=device)
    for i in range(n_vol):
        logits = torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * 3.0
        out[i] = torch.softmax(logits.double() if dtype == torch.float64 else logits, dim=1).to(dtype)
        del logits
    return out


def algorithmic_bytes_per_voxel(wl):
    return wl["N"] * wl["C"] * ELEM[wl["dtype"]] + 3 * 4 + 1


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs of the GPU's NUMA node before it allocates pinned host memory
    (first touch places the pages there).  Returns a description; a no-op where sysfs says nothing."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) <= 1:
            return f"single NUMA node (gpu {index} numa_node={node}, {len(nodes)} node(s))"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to NUMA node {node} ({len(cpus)} cpus) of gpu {index}"
        return f"NUMA node {node} has no allowed cpus"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


# ------------------------------------------------------------------------------ CPU reference arm
# The reference's own path for one unit of work on the host, on WHOLE volumes / images of the workload in the
# workload's dtype: C2 maps, the arg-max of the mean, then the workload's C3 aggregations on each of the
# three maps (scipy FFT box sum exactly as the reference calls it); for cfg3 DataCarrier3D.concat_data over
# every patch, the maps on the raw sums and the save path's normalisation.
# Implementation: the UNMODIFIED reference imported from baseline/_ref (offline pip install,
# oracle/ref_loader.py) when present -- kind "reference" -- else the oracle port -- kind "port".
_CPU = {}


def _cpu_impl():
    if "impl" not in _CPU:
        from oracle import ref_loader

        if ref_loader.available():
            _CPU["impl"], _CPU["kind"] = ref_loader.load(), "reference"
        else:
            from oracle import values_oracle as vo

            _CPU["impl"], _CPU["kind"] = vo, "port"
    return _CPU["impl"], _CPU["kind"]


def cpu_unit_shape(wl):
    """The reference arm works on whole volumes; only cfg3's 256^3 volume (minutes per volume through the
    reference's per-patch host loop) is bounded to a 128^3 volume with the same patch size and overlap."""
    return (128, 128, 128) if "stitch" in wl else tuple(wl["spatial"])


def _cpu_worker_init(wl, threads, seed):
    torch.set_num_threads(threads)
    _cpu_impl()
    g = torch.Generator().manual_seed(seed + os.getpid() % 1000)
    dt = torch.float32 if wl["dtype"] == "bf16" else DTYPES[wl["dtype"]]
    if "stitch" in wl:
        from oracle import values_oracle as vo

        p, shape = wl["stitch"]["patch"], cpu_unit_shape(wl)
        crops = vo.patch_grid(shape, p, wl["stitch"]["overlap"])
        logits = torch.randn((wl["N"], len(crops), wl["C"], p, p, p), generator=g) * 3.0
        _CPU["patches"], _CPU["crops"] = torch.softmax(logits, dim=2), crops
    else:
        shape = (wl["N"], wl["C"]) + cpu_unit_shape(wl)
        _CPU["x"] = torch.softmax(torch.randn(shape, generator=g) * 3.0, dim=1).to(dt)
    _CPU["wl"] = wl


def _cpu_stitch_step(impl, wl):
    patches, crops, shape = _CPU["patches"], _CPU["crops"], cpu_unit_shape(wl)
    p = wl["stitch"]["patch"]
    if hasattr(impl, "DataCarrier3D"):
        carrier = impl.DataCarrier3D()
    else:
        carrier = impl.StitchOracle()
    n_pred = patches.shape[0]
    for pred_idx in range(n_pred):
        for i, crop in enumerate(crops):
            batch = {"image_paths": ["vol.npy"], "label_paths": [["lab.npy"]], "org_image_size": [shape],
                     "crop_idx": [crop], "data": torch.zeros(1, 1, p, p, p),
                     "seg": torch.zeros(1, 1, p, p, p, dtype=torch.int32)}
            carrier.concat_data(batch, patches[pred_idx, i:i + 1], n_pred=n_pred, pred_idx=pred_idx)
    v = carrier.data["vol.npy"]
    d = impl.calculate_uncertainty(torch.from_numpy(v["softmax_pred"]))       # test_3D.py:532-533
    cnt = np.clip(v["num_predictions"], 1, None)[0]
    out = [float((np.asarray(d[k]) / cnt).sum()) for k in MAPS]                # data_carrier_3D.py:326-363
    return int(np.prod(shape)), out


def _cpu_worker_step(_):
    impl, _k = _cpu_impl()
    wl = _CPU["wl"]
    if "stitch" in wl:
        return _cpu_stitch_step(impl, wl)
    x = _CPU["x"]
    thr = (0.5, 0.4, 0.05)
    d = impl.calculate_uncertainty(x)
    mean_seg = torch.argmax(torch.mean(x, dim=0), dim=0)  # data_carrier_3D.py:253-259 / test_2D.py:119-127
    out = []
    for k, key in enumerate(MAPS):
        m = d[key].numpy()
        row = [impl.image_level_aggregation(m)["max_score"]]
        if wl["patch"] is not None:
            row.append(impl.patch_level_aggregation(m, wl["patch"])["max_score"])
        if wl["thr"]:
            row.append(float(impl.threshold_aggregation(m, threshold=thr[k])["max_score"]))
        out.append(row)
    return int(mean_seg.numel()), out


class CpuReference:
    """Pool of worker processes, each running the reference path on its own whole volume with its
    share of the host threads; one step = one volume per worker."""

    def __init__(self, wl):
        import torch.multiprocessing as mp

        cores = os.cpu_count() or 1
        unit_bytes = wl["N"] * wl["C"] * int(np.prod(cpu_unit_shape(wl))) * 8
        self.workers = max(1, min(8, cores // 4, int(128e9 // (12 * unit_bytes)) or 1))   # ~12 temporaries per stack, 128 GB of host memory at most
        self.threads = max(1, cores // self.workers)
        self.cores = self.workers * self.threads
        self.wl = wl
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.workers, initializer=_cpu_worker_init, initargs=(wl, self.threads, 4321))
        self.kind = _cpu_impl()[1]

    def step(self):
        res = self.pool.map(_cpu_worker_step, range(self.workers))
        return sum(r[0] for r in res)

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        vox = 0
        for _ in range(steps):
            vox += self.step()
        dt = time.perf_counter() - t0
        return vox / dt, dt

    def describe(self, steps, dt):
        unit = cpu_unit_shape(self.wl)
        whole = "whole" if unit == tuple(self.wl["spatial"]) else "bounded (same patch size and overlap)"
        return (f"{steps} step(s) x {self.workers} {whole} volume(s) {unit} of the workload "
                f"(N={self.wl['N']}, C={self.wl['C']}, {self.wl['dtype'] if self.wl['dtype'] != 'bf16' else 'f32'}), "
                f"{self.workers} worker processes x {self.threads} torch threads, {dt:.1f} s, "
                f"os.cpu_count()={os.cpu_count()}; implementation: "
                + ("unmodified reference functions from baseline/_ref" if self.kind == "reference"
                   else "oracle/values_oracle.py port"))

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_arm(args, wl, rank):
    if rank != 0:
        return
    ref = CpuReference(wl)
    value, dt = ref.run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl["dtype"] if wl["dtype"] != "bf16" else "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "volumes_per_step": ref.workers,
                   "volume_shape": list(cpu_unit_shape(wl)),
                   "note": "the reference's own CPU implementation of the path on the host cores, whole "
                           "volumes in the workload's dtype; rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                         "sample": ref.describe(args.steps, dt)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    ref.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm: workloads
class StackWorkload:
    """cfg1 / cfg2 / cfg4 / cfg5: a pool of softmax stacks resident in HBM -> UncertaintyPipeline."""

    def __init__(self, vb, wl, dev, rank, overlap, pool=None):
        self.vb, self.wl, self.dev = vb, wl, dev
        self.pool = pool or wl["pool"]
        self.V = int(np.prod(wl["spatial"]))
        gen = torch.Generator(device=dev).manual_seed(1234 + 1000 * wl["cfg"] + rank)
        self.stack = make_stack(gen, self.pool, wl, dev, DTYPES[wl["dtype"]])
        self.pool_bytes = self.stack.numel() * self.stack.element_size()
        self.thr = None
        if wl["thr"]:   # 0.98-quantile of a pilot map per uncertainty type (stand-in for threshold_analysis.json)
            pilot = vb.uncertainty_fused(self.stack[:1])
            sub = slice(None, None, max(1, self.V // (1 << 20)))
            self.thr = tuple(float(torch.quantile(m.reshape(-1)[sub].float(), 0.98).item())
                             for m in (pilot.pred_entropy, pilot.expected_entropy, pilot.mutual_information))
            del pilot
        self.cfg = vb.AggregationConfig(patch_size=wl["patch"], thresholds=self.thr, overlap=overlap)
        self.pipe = vb.UncertaintyPipeline(self.cfg)
        self.events = []
        self.pipe.k1_timer = self.events   # (start, end, n_volumes) per K1 launch, on the launch stream
        self.kernel = "k1 fused N x C reduction (" + ("k1_tma_kernel" if wl["dtype"] != "f64" or wl["N"] in (5, 8, 10, 16)
                                                      else "k1_smem_kernel") + ")"
        self.bytes_per_voxel = algorithmic_bytes_per_voxel(wl)
        self.step_bytes = self.bytes_per_voxel * self.V * self.pool   # V (N C s + 13) per volume
        self.units_per_step = self.pool * self.V
        self.row_len = 3 * 7

    def step(self):
        return self.pipe.run(self.stack, mean_argmax=True)

    def table(self, res):
        return res.scores.reshape(self.pool, -1)

    def dominant(self):
        """(avg launch ms, algorithmic bytes per launch, launches, sum of launch ms)"""
        ms = [s.elapsed_time(e) for s, e, _ in self.events]
        vols = [n for _, _, n in self.events]
        return (sum(ms) / len(ms), self.bytes_per_voxel * self.V * (sum(vols) / len(vols)), len(ms), sum(ms))


class StitchWorkload:
    """cfg3: K3 stitches every patch of one volume into the fp64 raw sums (each output voxel written once),
    K1 (fp64) computes the maps on the raw sums, the save path divides them by clip(count, 1)."""

    def __init__(self, vb, wl, dev, rank):
        self.vb, self.wl, self.dev = vb, wl, dev
        shape, p, ov = tuple(wl["spatial"]), wl["stitch"]["patch"], wl["stitch"]["overlap"]
        self.pool, self.V = 1, int(np.prod(shape))
        crops = vb.patch_grid(shape, p, ov)
        gen = torch.Generator(device=dev).manual_seed(1234 + 1000 * wl["cfg"] + rank)
        pdt = DTYPES[wl["patch_dtype"]]
        self.patches = torch.empty((wl["N"], len(crops), wl["C"], p, p, p), dtype=pdt, device=dev)
        for n in range(wl["N"]):
            logits = torch.randn(self.patches.shape[1:], generator=gen, device=dev) * 3.0
            self.patches[n] = torch.softmax(logits, dim=1).to(pdt)
            del logits
        self.crop_lo = vb.stitching.crops_to_lo(crops, dev)
        self.weight = (vb.gaussian_importance_factors((p, p, p), device=dev)
                       if wl["stitch"].get("weight") == "gaussian" else None)
        self.clip_min = 1.0 if self.weight is None else 0.0
        self.sums = torch.empty((wl["N"], wl["C"]) + shape, dtype=torch.float64, device=dev)
        self.count = torch.empty(shape, dtype=torch.float64, device=dev)
        self.maps = torch.empty((1, 3) + shape, dtype=torch.float32, device=dev)
        self.scores = torch.zeros((1, 3, 7), dtype=torch.float64, device=dev)
        self.argmax = torch.empty((1,) + shape, dtype=torch.uint8, device=dev)
        self.pool_bytes = self.patches.numel() * self.patches.element_size()
        self.events = []
        self.kernel = "k3 stitch accumulator (stitch_box_kernel" + (", separable weights)" if self.weight is not None else ")")
        k3 = self.pool_bytes + self.sums.numel() * 8 + self.count.numel() * 8
        k1 = self.V * (wl["N"] * wl["C"] * 8 + 13)
        norm = self.V * (3 * (4 + 8) + 8)
        self.k3_bytes, self.step_bytes = k3, k3 + k1 + norm
        self.units_per_step = self.V
        self.row_len = 3 * 7
        self.cfg = vb.AggregationConfig(patch_size=None)

    def step(self):
        vb = self.vb
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vb.stitch_accumulate(self.patches, self.crop_lo, self.sums, self.count, accumulate=False, weight=self.weight)
        e1.record()
        self.events.append((e0, e1, 1))
        vb.uncertainty_fused(self.sums.unsqueeze(0), maps=True, mean_argmax=True, scores=True, out_maps=self.maps,
                             volume_major=True, out_scores=self.scores[:, :, :3], out_argmax=self.argmax)
        self.norm = vb.normalize_maps(self.maps[0], self.count, self.clip_min)
        return self

    def wait(self):
        return self

    def table(self, res):
        return self.scores.reshape(1, -1)

    def dominant(self):
        ms = [s.elapsed_time(e) for s, e, _ in self.events]
        return (sum(ms) / len(ms), self.k3_bytes, len(ms), sum(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)   # the BURST region (~65 ms on cfg5); the sustained region follows
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=0, help="override volumes per rank per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="K2b on a side stream under the next K1")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.pool:
        wl["pool"] = args.pool
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return

    import torch.distributed as dist

    import values_b200 as vb

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    work = StitchWorkload(vb, wl, dev, rank) if "stitch" in wl else StackWorkload(vb, wl, dev, rank, args.overlap)
    peak, peak_sustained, peak_src = measured_peaks()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    align = torch.zeros(1, device=dev)

    def timed(n_steps):
        """EXACTLY n_steps steps between two events, barrier + synchronize on both sides, ONE score gather
        at the end of the region (inside it), max over ranks."""
        sync_all()
        work.events.clear()
        launches0 = vb._lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            if world > 1:   # ranks leave the host barrier milliseconds apart: an on-stream collective right in
                dist.all_reduce(align)   # front of e0 starts the timed region at the same instant on every GPU
            e0.record()
            tables = []
            for _ in range(n_steps):          # only the score rows of a step are kept (its maps / arg-max are not)
                tables.append(work.table(work.step()))
            tables = torch.stack(tables)
            if world > 1:   # the only collective of the path: every rank gets every image's scores
                gathered = torch.empty((world,) + tuple(tables.shape), dtype=tables.dtype, device=dev)
                dist.all_gather_into_tensor(gathered, tables)
            e1.record()
            sync_all()
        ms = e0.elapsed_time(e1)
        k_ms, k_bytes, k_n, k_sum = work.dominant()
        per_rank = None
        if world > 1:   # the line reports the MAX over ranks; every rank's own figures go along for diagnosis
            mine = torch.tensor([ms, k_ms], dtype=torch.float64, device=dev)
            allr = torch.empty((world, 2), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allr, mine)
            allr = allr.cpu()
            ms = float(allr[:, 0].max())
            per_rank = {"elapsed_ms": [round(float(v), 3) for v in allr[:, 0]],
                        "kernel_avg_ms": [round(float(v), 4) for v in allr[:, 1]]}
            k_ms = float(allr[:, 1].max())      # the roofline fraction is the slowest rank's too
        achieved = k_bytes / (k_ms * 1e-3) / 1e9
        return {"elapsed_ms": ms, "steps": n_steps, "launches": vb._lib.launch_count() - launches0,
                "value": float(work.units_per_step) * world * n_steps / (ms * 1e-3), "ms_per_step": ms / n_steps,
                "achieved": achieved, "kernel_ms": k_ms, "kernel_bytes": k_bytes, "kernel_launches": k_n,
                "kernel_share": k_sum / ms, "pipeline_gbs": work.step_bytes * n_steps / (ms * 1e-3) / 1e9,
                "clocks": clocks.summary(), "per_rank": per_rank}

    for _ in range(args.warmup):
        work.step()
    burst = timed(args.steps)
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(math.ceil(SUSTAINED_S * 1e3 / burst["ms_per_step"])))
        s = timed(n_sus)
        sus_peak = peak_sustained or peak
        sustained = {"value": s["value"], "unit": UNIT, "steps": n_sus, "seconds": s["elapsed_ms"] * 1e-3,
                     "ms_per_step": s["ms_per_step"], "achieved": s["achieved"], "peak": sus_peak,
                     "frac": s["achieved"] / sus_peak, "pipeline_frac": s["pipeline_gbs"] / sus_peak,
                     "avg_launch_ms": s["kernel_ms"], "clocks": s["clocks"],
                     "peak_source": "MEASURED_PEAKS.json sustained figure" if peak_sustained else
                                    "MEASURED_PEAKS.json hbm_gbs (burst copy figure; no sustained figure in the file)"}

    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath) and args.workload == "cfg5":
        with open(tpath) as f:
            tj = json.load(f)
        # DRAM bytes per launch = per-voxel figure of the ncu --set full capture x voxels per launch
        traffic = tj["dram_bytes_per_voxel"] * burst["kernel_bytes"] / work.bytes_per_voxel
        traffic_src = tj.get("source", "profiles/k1_traffic.json") + " (ncu dram__bytes_read.sum + dram__bytes_write.sum per voxel x voxels per launch)"
    roofline = {"bound": "hbm", "achieved": burst["achieved"], "peak": peak, "unit": "GB/s",
                "frac": burst["achieved"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": burst["kernel_bytes"], "kernel": work.kernel,
                "algorithmic_bytes_per_voxel": getattr(work, "bytes_per_voxel", None),
                "avg_launch_ms": burst["kernel_ms"], "launches_timed": burst["kernel_launches"],
                "kernel_share_of_step": burst["kernel_share"],
                # the whole step against the same roofline: algorithmic bytes of the fused pipeline
                # (V (N C s + 13) per volume; cfg3: B3 + K1 + normalisation) over the step time
                "pipeline_frac": burst["pipeline_gbs"] / peak, "peak_source": peak_src}

    sweep = None
    if wl.get("batch_sweep") and not args.no_sweep and isinstance(work, StackWorkload):
        sweep = []
        keep = work
        for b in wl["batch_sweep"]:   # cfg4 batch sweep B in {1, 2, 4, 6, 12}
            work = StackWorkload(vb, wl, dev, rank, args.overlap, pool=b)
            for _ in range(3):
                work.step()
            s = timed(max(5, min(args.steps, 40 // b + 4)))
            sweep.append({"B": b, "value": s["value"], "ms_per_step": s["ms_per_step"], "k1_frac": s["achieved"] / peak,
                          "pipeline_frac": s["pipeline_gbs"] / peak})
            del work
            torch.cuda.empty_cache()
        work = keep

    # end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(vb, wl, work, dev, world, args)
        e2e["numa"] = numa
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(wl)
        cpu_steps = 3
        cpu_value, cpu_s = ref.run(cpu_steps, 1)
        cpu = {"value": cpu_value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
               "sample": ref.describe(cpu_steps, cpu_s)}
        ref.close()
        if isinstance(work, StackWorkload):
            parity = parity_sample(vb, work)
    if rank == 0:
        line = {
            "metric": METRIC, "value": burst["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": burst["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["name"], "volumes_per_rank_per_step": work.pool,
                       "pool_bytes_per_rank": work.pool_bytes,
                       "l2": "inputs (pool >> 126 MB L2) stream from HBM every step",
                       "map_chunk_bytes": work.cfg.chunk_bytes,
                       "streams": "K2b on a side stream under the next batch's K1" if work.cfg.overlap else "single stream",
                       "aggregations": "image_level" + (" + threshold(0.98-quantile pilot)" if wl["thr"] else "")
                                       + (f" + patch_level({wl['patch']})" if wl["patch"] else ""),
                       "sharding": f"volumes sharded over {world} rank(s); one all_gather of the score tables of "
                                   f"the timed region ({args.steps} x {work.pool} rows per rank), inside it"},
            "roofline": roofline, "sustained": sustained, "per_rank": burst["per_rank"], "cpu_baseline": cpu,
            "parity": parity,
            "clocks": burst["clocks"], "e2e": e2e, "gpu_launches": burst["launches"],
        }
        if sweep is not None:
            line["batch_sweep"] = sweep
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_sample(vb, work):
    """Decision parity of ONE whole volume of the pool against the reference's own functions on the host
    (oracle/parity.py): how many voxels of the threshold masks and of the arg-max differ, and whether every
    one of them lies inside the fp32 tolerance of the maps."""
    from oracle.parity import parity_counts

    impl, kind = _cpu_impl()
    x = work.stack[0]
    xh = x.float().cpu() if x.dtype == torch.bfloat16 else x.cpu()
    ref = impl.calculate_uncertainty(xh)
    thr = work.thr or tuple(float(np.median(np.asarray(ref[k]))) for k in MAPS)
    res = vb.uncertainty_fused(x.unsqueeze(0), mean_argmax=True)
    got = {k: v.cpu().numpy() for k, v in res.as_dict(0).items()}
    means = torch.mean(xh.double(), dim=0).numpy()
    rep = parity_counts(got, {k: np.asarray(ref[k]) for k in MAPS}, thr,
                        rtol=1e-3 if x.dtype == torch.bfloat16 else 1e-5, atol=1e-5 if x.dtype == torch.bfloat16 else 1e-6,
                        got_argmax=res.mean_argmax[0].cpu().numpy(),
                        ref_argmax=torch.argmax(torch.mean(xh, dim=0), dim=0).numpy(), class_means=means)
    return {"sample": f"volume 0 of the pool, {kind}'s calculate_uncertainty on the host", "voxels": rep["voxels"],
            "explained_by_map_tolerance": rep["explained"], "argmax_mismatches": rep.get("argmax_mismatches"),
            "mask_flips": {k: rep[k]["mask_flips"] for k in MAPS}, "thresholds": [rep[k]["threshold"] for k in MAPS],
            "max_abs_diff": {k: rep[k]["max_abs_diff"] for k in MAPS},
            "beyond_tolerance": {k: rep[k]["beyond_tolerance"] for k in MAPS}}


def run_e2e(vb, wl, work, dev, world, args):
    """Same metric through the public API from pinned HOST buffers: per step, H2D of that step's
    inputs (double-buffered against compute on a copy stream) and D2H of the score table."""
    import torch.distributed as dist

    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    if isinstance(work, StitchWorkload):
        return run_e2e_stitch(vb, wl, work, dev, world, args, main, copy_stream)
    stack = work.stack
    n = min(wl["e2e_pool"], stack.shape[0])
    V = work.V
    host = torch.empty((n,) + tuple(stack.shape[1:]), dtype=stack.dtype, pin_memory=True)
    host.copy_(stack[:n])
    # several small volumes travel as one batch (cfg1 / cfg2: a 64^3 stack is 10-20 MB)
    group = max(1, min(n, (256 << 20) // max(1, host[0].numel() * host.element_size())))
    groups = [(i, min(i + group, n)) for i in range(0, n, group)]
    host_scores = torch.empty((n, 3 * 7), dtype=torch.float64, pin_memory=True)
    pipe = vb.UncertaintyPipeline(work.cfg)
    bufs = [torch.empty((group,) + tuple(stack.shape[1:]), dtype=stack.dtype, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def step():
        for gi, (a, b) in enumerate(groups):
            s = gi & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                bufs[s][:b - a].copy_(host[a:b], non_blocking=True)
                ready[s].record(copy_stream)
            main.wait_event(ready[s])
            res = pipe.run(bufs[s][:b - a], mean_argmax=True)
            host_scores[a:b].copy_(res.scores.reshape(b - a, -1), non_blocking=True)
            freed[s].record(main)
        main.synchronize()  # the caller reads the scores: the step ends when they are on the host

    for s in range(2):
        freed[s].record(main)
    for _ in range(2):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d = int(host.numel() * host.element_size())
    return {"value": n * V * world * steps / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(host_scores.numel() * 8),
            "volumes_per_step": n, "steps": steps, "h2d_gbs_per_gpu": h2d * steps / (ms * 1e-3) / 1e9,
            "api": "UncertaintyPipeline.run on pinned host stacks (H2D double-buffered) -> host score table"}


def run_e2e_stitch(vb, wl, work, dev, world, args, main, copy_stream):
    """cfg3 end to end: the patches of one volume come from pinned host memory (one TTA sample per copy,
    double-buffered against the stitch of the previous sample), the three normalised-map sums go back."""
    import torch.distributed as dist

    N = wl["N"]
    host = torch.empty(tuple(work.patches.shape), dtype=work.patches.dtype, pin_memory=True)
    host.copy_(work.patches)
    host_scores = torch.empty((3 * 7,), dtype=torch.float64, pin_memory=True)
    bufs = [torch.empty_like(work.patches[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def step():
        for n in range(N):
            s = n & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                bufs[s].copy_(host[n], non_blocking=True)
                ready[s].record(copy_stream)
            main.wait_event(ready[s])
            vb.stitch_accumulate(bufs[s].unsqueeze(0), work.crop_lo, work.sums[n:n + 1],
                                 work.count if n == 0 else None, accumulate=False, weight=work.weight)
            freed[s].record(main)
        vb.uncertainty_fused(work.sums.unsqueeze(0), maps=True, mean_argmax=True, scores=True, out_maps=work.maps,
                             volume_major=True, out_scores=work.scores[:, :, :3], out_argmax=work.argmax)
        work.norm = vb.normalize_maps(work.maps[0], work.count, work.clip_min)
        host_scores.copy_(work.scores.reshape(-1), non_blocking=True)
        main.synchronize()

    for s in range(2):
        freed[s].record(main)
    step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d = int(host.numel() * host.element_size())
    return {"value": work.V * world * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(host_scores.numel() * 8), "volumes_per_step": 1, "steps": steps,
            "h2d_gbs_per_gpu": h2d * steps / (ms * 1e-3) / 1e9,
            "api": "stitch_accumulate on pinned host patches (one TTA sample per copy, double-buffered) + "
                   "uncertainty_fused + normalize_maps -> host score row"}


if __name__ == "__main__":
    main()
