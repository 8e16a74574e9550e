#!/usr/bin/env python
"""Developer micro-benchmark: per-op device time and algorithmic GB/s (CUDA events).
Not the contract benchmark (that is bench.py); used while tuning kernels."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops, synth  # noqa: E402


ITERS, WARMUP = [20], [5]


def timeit(fn, iters=None, warmup=None, flush=None):
    iters = ITERS[0] if iters is None else iters
    warmup = WARMUP[0] if warmup is None else warmup
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shape", default="kitti")
    ap.add_argument("--ops", default="project,knn,assemble,unproject,select,lovasz")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--sensor-order", action="store_true", help="points sorted by (beam, azimuth)")
    args = ap.parse_args()
    shp = synth.SHAPES[args.shape]
    ITERS[0], WARMUP[0] = args.iters, args.warmup
    B = args.batch
    one, offs1, _, _ = synth.make_batch(shp, min(B, 8), seed0=1000, sensor_order=args.sensor_order)
    reps = (B + 7) // 8
    pts = np.concatenate([one] * reps, 0)
    sizes = np.tile(np.diff(offs1), reps)[:B]
    pts = pts[:sizes.sum()]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    N, HW = pts.shape[0], shp.proj_h * shp.proj_w
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    bufs = ops.ProjectionBuffers(B, N, 4, shp.proj_h, shp.proj_w, "cuda")
    if "project" in args.ops:
        for mode in ("0", "1", "fused_kernel", "cluster_kernel"):
            f64 = mode == "1"
            two = mode == "fused_kernel"
            clu = mode == "cluster_kernel"
            kern = None
            if os.environ.get("C3D_MB_PROFILE"):   # per-kernel events perturb the total
                with ops.profile("") as prof:
                    timeit(lambda: ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs, exact_f64=f64, fused_kernel=two, cluster_kernel=clu), flush=flush)
                    kern = {n: round(1e3 * v[0] / (ITERS[0] + WARMUP[0]), 1) for n, v in prof.all().items()}
            med, mn = timeit(lambda: ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs, exact_f64=f64, fused_kernel=two, cluster_kernel=clu), flush=flush)
            by = 28 * N + 28 * B * HW
            res["project_f64only=" + mode] = dict(ms=med, ms_min=mn, GBs=by / med / 1e6, scans_s=B / med * 1e3,
                                                  kernels_us_per_call=kern)
    pr = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs)
    if "knn" in args.ops:
        for idt, name in ((torch.int64, "i64"), (torch.int32, "i32")):
            am = torch.randint(0, shp.n_classes, pr.proj_idx.shape, device="cuda").to(idt)
            px, py = pr.uproj_x_idx.to(idt), pr.uproj_y_idx.to(idt)
            out = torch.empty(N, dtype=idt, device="cuda")
            med, mn = timeit(lambda: ops.knn_batch(pr.proj_range, am, pr.uproj_depth, px, py, O, 5, 5, 1.0, 1.0,
                                                   shp.n_classes, out=out), flush=flush)
            by = 12 * B * HW + 28 * N
            res["knn_" + name] = dict(ms=med, ms_min=mn, GBs_ref_dtypes=by / med / 1e6, scans_s=B / med * 1e3)
    if "assemble" in args.ops:
        weak = torch.randint(0, shp.n_classes, (N,), device="cuda", dtype=torch.int32)
        sem = torch.randint(1, shp.n_classes, (N,), device="cuda", dtype=torch.int32)
        mean = torch.tensor([12.12, 10.88, 0.23, -1.04, 0.21], device="cuda")
        std = torch.tensor([12.32, 11.47, 6.91, 0.86, 0.16], device="cuda")
        med, mn = timeit(lambda: ops.project_assemble_batch(P, O, fov, shp.proj_h, shp.proj_w, sem, weak, mean, std,
                                                             buffers=bufs), flush=flush)
        by = 28 * N + 44 * B * HW  # points 16 + upx/upy/udepth 12 per point; feature 20 + labels 16 + range 4 + idx 4 per pixel
        res["project_assemble"] = dict(ms=med, ms_min=mn, GBs=by / med / 1e6, scans_s=B / med * 1e3)
    if "unproject" in args.ops:
        am = torch.randint(0, shp.n_classes, pr.proj_idx.shape, device="cuda")
        lab = torch.randint(0, shp.n_classes, (N,), device="cuda")
        conf = torch.zeros((shp.n_classes, shp.n_classes), dtype=torch.int64, device="cuda")
        med, mn = timeit(lambda: ops.unproject_confusion_batch(am, pr.uproj_x_idx, pr.uproj_y_idx, O, shp.n_classes,
                                                               labels=lab, conf_matrix=conf), flush=flush)
        by = (4 + 4 + 8 + 8 + 8) * N  # px, py i32; label i64; gathered class i64; output i64
        res["unproject_confusion"] = dict(ms=med, ms_min=mn, GBs=by / med / 1e6, scans_s=B / med * 1e3)
    if "select" in args.ops:
        C = shp.n_classes
        g = torch.Generator(device="cuda").manual_seed(3)
        probs = torch.softmax(2 * torch.randn(B, C, shp.proj_h, shp.proj_w, device="cuda", generator=g), 1)
        ev = torch.rand(B, shp.proj_h, shp.proj_w, device="cuda", generator=g) < 0.7
        tl = torch.randint(1, C, ev.shape, device="cuda", generator=g) * (torch.rand(ev.shape, device="cuda", generator=g) < 1e-3) * ev
        wss = tl.gt(0)
        ws = torch.empty(ops.lib.c3d_entropy_select_workspace_bytes(B, C, HW), dtype=torch.uint8, device="cuda")
        with ops.profile("") as prof:
            med, mn = timeit(lambda: ops.entropy_select_batch(probs, wss, ev, tl, 0.5, seed=5, workspace=ws), flush=flush)
            kern = {n: round(1e3 * v[0] / v[1], 1) for n, v in prof.all().items()}
        by = (4 * C + 1 + 8 + 1 + 8 + 1) * B * HW  # probs, eval mask, train label, weak mask; label i64 + mask out
        res["entropy_select"] = dict(ms=med, ms_min=mn, GBs=by / med / 1e6, scans_s=B / med * 1e3, kernels_us=kern)
    if "lovasz" in args.ops:
        C = shp.n_classes
        g = torch.Generator(device="cuda").manual_seed(4)
        probs = torch.softmax(torch.randn(B, C, shp.proj_h, shp.proj_w, device="cuda", generator=g), 1).requires_grad_(True)
        lab = torch.randint(1, C, (B, shp.proj_h, shp.proj_w), device="cuda", generator=g) * \
            (torch.rand(B, shp.proj_h, shp.proj_w, device="cuda", generator=g) < shp.label_ratio)
        ws = torch.empty(ops.lib.c3d_lovasz_workspace_bytes(C, ops.LOVASZ_MAX_VALID), dtype=torch.uint8, device="cuda")

        def lov():
            probs.grad = None
            loss, _ = ops.lovasz_softmax(probs, lab, ignore=0, workspace=ws)
            loss.backward()
        with ops.profile("") as prof:
            med, mn = timeit(lov, flush=flush)
            kern = {n: round(1e3 * v[0] / v[1], 1) for n, v in prof.all().items()}
        by = (8 + 4 * C) * B * HW   # labels read + dense gradient written
        res["lovasz_fwd_bwd"] = dict(ms=med, ms_min=mn, GBs=by / med / 1e6, scans_s=B / med * 1e3,
                                     valid=ops.lovasz_info(ws)[0], kernels_us=kern)
    print(json.dumps(dict(batch=B, shape=args.shape, results=res), indent=1))


if __name__ == "__main__":
    main()
