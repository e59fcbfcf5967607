#!/usr/bin/env python
"""Experiment (GPU box): K1 on a large SM partition and K2b on a small one at the same time (CUDA green
contexts through cuda-python).  K1 is HBM-bound with issue headroom in bursts; K2b is bound by the shared-memory
pipe, the L2 fabric and latency -- do they add up to less than their sum when each has its own SMs?

    python tools/green_ctx_probe.py [sms_for_k2b]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda.bindings import driver as cu

import bench
import values_b200 as vb


def ck(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def main():
    n_small = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    dev = ck(cu.cuDeviceGet(0))
    sm = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    print("SMs:", sm.sm.smCount)
    groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, sm, 0, n_small))
    small, large = groups[0], rem
    print("split:", small.sm.smCount, "+", large.sm.smCount)
    streams = []
    for res in (large, small):
        desc = ck(cu.cuDevResourceGenerateDesc([res], 1))
        gctx = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        st = ck(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        streams.append(torch.cuda.ExternalStream(int(st)))
    s_big, s_small = streams
    wl = dict(bench.WORKLOADS["cfg5"])
    gen = torch.Generator(device="cuda").manual_seed(1)
    stack = bench.make_stack(gen, 32, wl, torch.device("cuda"), torch.float32)
    spatial = tuple(wl["spatial"])
    maps = [torch.empty((32, 3) + spatial, dtype=torch.float32, device="cuda") for _ in range(2)]
    sc = torch.zeros((32, 3, 7), dtype=torch.float64, device="cuda")
    flat = sc.view(96, 7)
    ws1 = torch.empty(vb._lib.lib.values_uncertainty_workspace_bytes(32, stack[0, 0, 0].numel(), 0) + 8, dtype=torch.uint8, device="cuda")
    ws2 = [torch.empty(vb.aggregation.patch_max_workspace_bytes(96, spatial, 10) + 8, dtype=torch.uint8, device="cuda") for _ in range(2)]

    def k1(i):
        vb.uncertainty_fused(stack, maps=True, mean_argmax=False, scores=True, thresholds=(0.5, 0.4, 0.05), out_maps=maps[i & 1],
                             volume_major=True, out_scores=sc[:, :, :3], workspace=ws1)

    def k2(i):
        vb.patch_max(maps[i & 1].view((96,) + spatial), 10, out_score=flat[:, 3], out_bbox=flat[:, 4:7], workspace=ws2[i & 1])

    def timed(fn, n=10):
        fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for i in range(n):
            fn(i)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    print("default stream: K1 %.3f ms, K2b %.3f ms, K1 then K2b %.3f ms" % (timed(k1), timed(k2), timed(lambda i: (k1(i), k2(i)))))
    with torch.cuda.stream(s_big):
        print("K1 on %d SMs: %.3f ms" % (large.sm.smCount, timed(k1)))
    with torch.cuda.stream(s_small):
        print("K2b on %d SMs: %.3f ms" % (small.sm.smCount, timed(k2)))
    # pipelined: K1 of step i on the large partition, K2b of step i-1 on the small one
    n = 12
    ev_k1 = [torch.cuda.Event() for _ in range(n + 1)]
    ev_k2 = [torch.cuda.Event() for _ in range(n + 1)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_big)
    for i in range(n):
        with torch.cuda.stream(s_big):
            if i >= 2:
                s_big.wait_event(ev_k2[i - 2])       # the K2b that read this map buffer two steps ago is done
            k1(i)
            ev_k1[i].record(s_big)
        with torch.cuda.stream(s_small):
            s_small.wait_event(ev_k1[i])
            k2(i)
            ev_k2[i].record(s_small)
    s_big.wait_event(ev_k2[n - 1])
    e1.record(s_big)
    torch.cuda.synchronize()
    print("pipelined on the two partitions: %.3f ms per step" % (e0.elapsed_time(e1) / n))


if __name__ == "__main__":
    main()
