#!/usr/bin/env python
"""Benchmark of the COARSE3D per-scan hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic scans per GPU:
projection -> prototype loss fwd+bwd -> EMA prototype update (+ all-reduce of the
packed prototype sums when N > 1) -> KNN vote.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/s (project+proto-loss fwd/bwd+KNN)"

# BASELINE.json configs[0..4] -> --config 1..5.  Every config is the full step of the metric
# (projection -> EMA prototype update -> prototype loss fwd+bwd -> KNN 5x5) on its shape/batch.
# Default = configs[4], the batch-64-per-GPU sweep the metric ("at 1/2/4/8 B200") is quoted on;
# it fits one GPU, so it is also the N=1 workload.
CONFIGS = {
    1: dict(shape="kitti", batch=1, dim=256, what="single KITTI-shaped scan, D=256 (the reference's CPU-runnable case)"),
    2: dict(shape="kitti", batch=8, dim=128, what="batch 8 KITTI-shaped scans, D=128"),
    3: dict(shape="nuscenes", batch=32, dim=128, what="batch 32 nuScenes-shaped scans at 0.01% labels, D=128"),
    4: dict(shape="poss", batch=16, dim=128, what="batch 16 SemanticPOSS-shaped scans, KNN 5x5, D=128"),
    5: dict(shape="kitti", batch=64, dim=128, what="scan-sharded sweep, batch 64 KITTI-shaped scans per GPU, D=128"),
}
DEFAULT_CONFIG = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_eager"])
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--shape", default=None, help="override the config's scan shape")
    ap.add_argument("--batch-per-gpu", type=int, default=None, help="override the config's batch")
    ap.add_argument("--dim", type=int, default=None, help="override the config's feature dim")
    ap.add_argument("--point-order", default="random", choices=["random", "sensor"],
                    help="order of the points inside a scan: random (BASELINE's synthetic spec, the hard "
                         "case for the z-buffer atomics and the KNN gathers) or sensor = (beam, azimuth), "
                         "what a spinning LiDAR's .bin file holds")
    ap.add_argument("--label-frac", type=float, default=None,
                    help="pseudo-label regime the reference trains in once entropy_selection is on "
                         "(tasks/weak_segmentation/trainer.py:654-686): the contrastive loss sees the weak labels "
                         "PLUS the class of this fraction of the occupied pixels (e.g. 0.3), the EMA update keeps "
                         "the weak labels (model.forward, :625-630); a measurement variant, not a BASELINE config")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one stream, no overlap of the chains")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-scans", type=int, default=8)
    ap.add_argument("--min-seconds", type=float, default=1.5,
                    help="the K-step block is repeated until the timed region is this long; "
                         "the median block is reported")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.shape = args.shape or cfg["shape"]
    args.batch_per_gpu = args.batch_per_gpu or cfg["batch"]
    args.dim = args.dim or cfg["dim"]
    return args


def load_synth():
    """The synthetic scan generator by file path: importing the `coarse3d_b200` package would
    map libcoarse3d_b200.so, which the reference arm must not do."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import baselines
    return baselines.load_synth()


def workload_config(args, world, synth):
    """Identical in both arms (the driver compares it)."""
    shp = synth.SHAPES[args.shape]
    HW = shp.proj_h * shp.proj_w
    inputs_mb = args.batch_per_gpu * (shp.n_points * 16 + HW * (4 * args.dim + 4 * shp.n_classes + 17)) >> 20
    return {
        "workload": "BASELINE.json configs[%d]: %s; %d %s-shaped scans per GPU (%d points, %dx%d, "
                    "%d classes, %.2g%% weak labels), D=%d, M=20, A=512; project + EMA prototype update "
                    "+ proto-loss fwd/bwd + KNN 5x5 k=5" % (
                        args.config - 1, CONFIGS[args.config]["what"], args.batch_per_gpu, shp.name,
                        shp.n_points, shp.proj_h, shp.proj_w, shp.n_classes, 100 * shp.label_ratio, args.dim),
        "config_id": args.config, "batch_per_gpu": args.batch_per_gpu,
        "global_batch": args.batch_per_gpu * world, "points_per_scan": shp.n_points,
        "proj": [shp.proj_h, shp.proj_w], "feature_dim": args.dim,
        "point_order": getattr(args, "point_order", "random"),
        "loss_label_frac": getattr(args, "label_frac", None),
        "parallelism": "scan-sharded x%d, one all-reduce of [K*D|K] prototype sums" % world,
        "l2": "no flush: 3 rotating input sets of ~%d MB re-read inputs each and a %d MB gradient "
              "written per step, both larger than the 126 MB L2" % (
                  inputs_mb, args.batch_per_gpu * HW * args.dim * 4 >> 20),
    }


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the
    timed region runs (B200_PROFILING.md 'clocks DURING the timed region')."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index, period=0.02):
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        self.period, self.marks = period, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.max = None
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), clk, util, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.t.join(timeout=1)

    def summary(self, t0, t1):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": [], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or \
                 [s for s in self.samples if s[2] > 0] or self.samples
        reasons = set()
        for _, _, _, rs in inside:
            for bit, name in {**self.BAD, **self.NOTE}.items():
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max,
                "reasons": sorted(reasons), "samples": len(inside),
                "samples_in_timed_region": len([s for s in self.samples if t0 <= s[0] <= t1])}


# ------------------------------------------------------------ CPU baseline --
def cpu_port_step(synth, n_scans, shape_name, dim, seed0=1000):
    """The oracle (CPU restatement of the reference) over `n_scans` scans of the workload:
    projection, EMA prototype update, loss fwd+bwd on the updated bank, KNN.  Returns seconds."""
    from oracle import knn as oknn, projection as oproj, proto_ema as oema, proto_loss as oloss
    shp = synth.SHAPES[shape_name]
    H, W, C, M = shp.proj_h, shp.proj_w, shp.n_classes, 20
    fov = oproj.Fov(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
    g = torch.Generator().manual_seed(seed0)
    scans = [synth.make_scan(shp, seed0 + i) for i in range(n_scans)]
    feats = torch.randn(n_scans, dim, H, W, generator=g)
    probs = torch.softmax(torch.randn(n_scans, C, H, W, generator=g), 1)
    argmax = torch.randint(0, C, (n_scans, H, W), generator=g).numpy()
    queue = torch.nn.functional.normalize(torch.randn(C, M, dim, generator=g), dim=-1)
    ln = [torch.ones(dim), torch.zeros(dim), torch.ones(C), torch.zeros(C)]
    t0 = time.perf_counter()
    projs, labels = [], []
    for pts, _, weak in scans:
        o = oproj.project(pts, fov)
        lab = np.zeros((H, W), np.int64)
        v = o["proj_idx"] >= 0
        lab[v] = weak[o["proj_idx"][v]]
        projs.append(o), labels.append(lab)
    labels = torch.from_numpy(np.stack(labels))
    new_q, _, _, _ = oema.prototype_learning(feats, labels, queue, *ln, C, 0, 0.999, gumbel=None,
                                             labelled_only=True)
    f = feats.clone().requires_grad_(True)
    loss, _, _ = oloss.contrast_mem_loss(f, probs, labels, labels > 0, new_q[None], temperature=0.07,
                                         num_anchor=512)
    loss.backward()
    for o, am in zip(projs, argmax):
        oknn.knn_vote(o["proj_range"], o["uproj_depth"], am, o["uproj_x_idx"], o["uproj_y_idx"],
                      5, 5, 1.0, 1.0, C)
    return time.perf_counter() - t0


class CpuArm:
    """The reference's CPU implementation of the path: the unmodified reference modules when a
    reference tree is present (build container), else the oracle port (GPU box)."""

    def __init__(self, args, synth):
        import baselines
        self.args, self.synth = args, synth
        self.kind, self.real = "port", None
        root = baselines.find_reference_root()
        if root is not None and os.environ.get("C3D_BENCH_PORT_ONLY", "0") != "1":
            try:
                self.real = baselines.RealReference(root)
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("reference tree at %s not usable (%r): timing the oracle port\n" % (root, e))
        torch.set_num_threads(os.cpu_count() or 1)

    def step(self, n_scans, seed0=1000):
        if self.real is not None:
            return self.real.step(self.synth, self.synth.SHAPES[self.args.shape], n_scans, self.args.dim, seed0)
        return cpu_port_step(self.synth, n_scans, self.args.shape, self.args.dim, seed0)

    def describe(self):
        return ("unmodified reference modules (RangeProjection, prototype block of SalsaNextProto, "
                "ContrastMEMLoss, KNN)" if self.real is not None else
                "oracle/ port (numpy projection 1 thread, torch-CPU EMA / loss fwd+bwd / KNN)")


def cpu_baseline(args, synth):
    arm = CpuArm(args, synth)
    arm.step(1)  # warm-up (imports, allocator)
    n = max(1, min(args.cpu_scans, args.batch_per_gpu))
    dts = [arm.step(n, seed0=1000 + 10 * r) for r in range(3)]
    dt = float(np.median(dts))
    return {"value": n / dt, "unit": "scans/s", "cores": torch.get_num_threads(), "kind": arm.kind,
            "sample": "median of 3 passes over %d scans of the workload through the %s; %.2f s per pass, "
                      "%.1f s of CPU work" % (n, arm.describe(), dt, sum(dts))}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on the host
    cores, W warm-up + exactly K timed steps, each a bounded sample of the workload's batch
    (sized so that the run ends within a few minutes).  Rank 0 only."""
    if rank != 0:
        return
    synth = load_synth()
    arm = CpuArm(args, synth)
    K, Wu = args.steps, max(args.warmup, 1)
    t1 = arm.step(1)                                     # calibration = first warm-up step
    budget = float(os.environ.get("C3D_REF_BUDGET_S", "120"))
    n = int(max(1, min(args.batch_per_gpu, budget / ((K + Wu) * max(t1, 1e-3)))))
    for _ in range(Wu - 1):
        arm.step(n)
    t_all = 0.0
    for i in range(K):
        t_all += arm.step(n, seed0=1000 + 10 * i)
    val = K * n / t_all
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "scans/s", "n_gpus": args.gpus,
        "steps": K, "warmup": Wu, "ms_per_step": 1e3 * t_all / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world, synth),
        "cpu_baseline": {"value": val, "unit": "scans/s", "cores": torch.get_num_threads(), "kind": arm.kind,
                         "sample": "each step = %d of the batch's %d scans through the %s" % (
                             n, args.batch_per_gpu, arm.describe())},
        "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_eager(args, rank, world):
    """`--impl torch_eager`: the reference's algorithm with stock torch ops on the same B200
    (tools/baselines.py), the "existing Blackwell path" of SURVEY.md 2b / 8d.  A bounded sample
    of the workload's batch per step (the dense prototype similarity needs 1.6 KB per pixel)."""
    if rank != 0:
        return
    synth = load_synth()
    import baselines
    shp = synth.SHAPES[args.shape]
    dev = torch.device("cuda", 0)
    H, W, C, M, D = shp.proj_h, shp.proj_w, shp.n_classes, 20, args.dim
    n = max(1, min(args.batch_per_gpu, args.cpu_scans))
    g = torch.Generator(device=dev).manual_seed(1000)
    state = dict(feats=torch.randn((n, D, H, W), device=dev, generator=g),
                 probs=torch.softmax(torch.randn((n, C, H, W), device=dev, generator=g), 1),
                 argmax=torch.randint(0, C, (n, H, W), device=dev, generator=g),
                 protos=torch.nn.functional.normalize(torch.randn((C, M, D), device=dev, generator=g), dim=-1),
                 ln_d=torch.nn.LayerNorm(D).to(dev), ln_c=torch.nn.LayerNorm(C).to(dev))
    scans = [synth.make_scan(shp, 1000 + i) for i in range(n)]
    K, Wu = args.steps, max(args.warmup, 1)
    for _ in range(Wu):
        baselines.torch_eager_step(synth, shp, scans, dev, D, state)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(K):
        baselines.torch_eager_step(synth, shp, scans, dev, D, state)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    val = K * n / dt
    print(json.dumps({
        "impl": "torch_eager", "metric": METRIC, "value": val, "unit": "scans/s", "n_gpus": 1, "steps": K,
        "warmup": Wu, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, synth),
        "sample": "each step = %d of the batch's %d scans: numpy projection on the host (as in the "
                  "reference loaders) + H2D, then the reference's torch statements on cuda:0 (dense "
                  "LayerNorm + prototype similarity, per-(scan, class) multinomial loop, autograd "
                  "backward, two unfolds + topk per scan for KNN); wall clock incl. the host part" % (
                      n, args.batch_per_gpu),
        "torch": torch.__version__, "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------- main --
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.impl == "torch_eager":
        run_torch_eager(args, rank, world)
        return

    import torch.distributed as dist
    from coarse3d_b200 import ops, synth
    from coarse3d_b200.pipeline import HotPathStep
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: coarse3d_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shp = synth.SHAPES[args.shape]
    B, K, Wu = args.batch_per_gpu, args.steps, max(args.warmup, 3)

    step = HotPathStep(shp, B, dim=args.dim, seed0=1000 + 10000 * rank, device=dev,
                       concurrent=not args.serial, sensor_order=args.point_order == "sensor",
                       loss_label_frac=args.label_frac)
    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (eager), then graph capture
    for i in range(Wu):
        step.run(i, seed=i)
    barrier()
    graphed = (not args.no_graph) and step.capture()
    for i in range(Wu):
        step.step(i)
    barrier()

    # ---- timed region: blocks of exactly K steps (CUDA events, barrier + synchronize on both
    # sides of every block, max over ranks per block); blocks are repeated until the region is
    # >= --min-seconds long so that the clock sampler sees it, and the MEDIAN block is reported.
    l0 = ops.launch_count()
    block_ms = []
    t_region0 = time.perf_counter()
    n_blocks = 1
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(K):
            step.step(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        block_ms.append(float(ms.item()))
        if len(block_ms) == 1:   # every rank derives the same block count from the reduced time
            n_blocks = int(min(5000, max(1, math.ceil(args.min_seconds * 1e3 / max(block_ms[0], 1e-3)))))
        if len(block_ms) >= n_blocks:
            break
    t_region1 = time.perf_counter()
    ms_total = float(np.median(block_ms))
    eager_launches_per_step = None
    if graphed:
        la = ops.launch_count(); step.run(0); torch.cuda.synchronize(dev)
        eager_launches_per_step = ops.launch_count() - la
        launches = eager_launches_per_step * K  # kernels inside the replayed graphs, per block
    else:
        launches = (ops.launch_count() - l0) // len(block_ms)
    clocks = sampler.summary(t_region0, t_region1)

    # ---- N > 1: every rank must hold the bit-identical prototype bank
    banks_identical = None
    if world > 1:
        mine = step.protos.detach().clone()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1 if torch.equal(mine, ref) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        banks_identical = bool(same.item())
        assert banks_identical, "prototype banks diverged across ranks"
        if step.peer is not None:
            assert step.peer.errors() == 0, "peer exchange: a rank's payload timed out (bits %x)" % step.peer.errors()

    # ---- per-kernel device times (eager, events inside the library around each launch)
    n_prof = 20
    step.concurrent = False  # one stream, so each kernel's events bracket only itself
    with ops.profile("") as prof:
        for i in range(n_prof):
            step.run(i, seed=i)
        torch.cuda.synchronize(dev)
        per_kernel = {k: {"us": 1e3 * v[0] / max(v[1], 1), "launches_per_step": v[1] / n_prof}
                      for k, v in prof.all().items()}
    # the dominant kernel in its concurrent setting (the fill daemon runs under the other chains)
    dom_concurrent_us = None
    if not args.serial and step.schedule == "fill_daemon":
        step.concurrent = True
        with ops.profile("fill_daemon_kernel") as prof:
            for i in range(n_prof):
                step.run(i, seed=i)
            torch.cuda.synchronize(dev)
            v = prof.read("fill_daemon_kernel")
            dom_concurrent_us = 1e3 * v[0] / max(v[1], 1)
    sampler.stop()

    alg = step.algorithmic_bytes()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # Dominant kernel = the one that writes the dense gradient (84 % of the step's bytes): the
    # fill daemon, the stand-alone fill, or the KNN vote carrying the fill.
    if "fill_daemon_kernel" in per_kernel:
        dom, dom_bytes = "fill_daemon_kernel", alg["loss_grad_fill"]
    elif "knn_vote_fill_kernel" in per_kernel:
        # the vote writes the share of the zero fill the schedule leaves with it
        dom, dom_bytes = "knn_vote_fill_kernel", step.fill_bytes_by_carrier()["knn_vote"] + alg["knn"]
    else:
        dom, dom_bytes = "fill_zero_kernel", alg["loss_grad_fill"]
    # time of ALL the kernel's launches in a step (the vote runs as two launches at batch 20..47)
    dom_launches = per_kernel.get(dom, {}).get("launches_per_step") or 1
    dom_us = per_kernel.get(dom, {}).get("us")
    if dom_us:
        dom_us *= dom_launches
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9 if dom_us else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            for tj in json.load(open(tpath)).get("entries", []):
                if (tj.get("kernel") == dom and tj.get("batch") == B and tj.get("dim") == args.dim
                        and tj.get("shape") == args.shape
                        and tj.get("cofill_bytes", dom_bytes - alg["knn"]) == dom_bytes - alg["knn"]):
                    traffic = tj.get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
    step_bytes = alg["project"] + alg["knn"] + alg["loss"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": dom_bytes,   # all launches of the kernel in one step
                "kernel_us_alone": dom_us, "kernel_launches_per_step": dom_launches,
                "kernel_us_under_the_step": dom_concurrent_us,
                "step_algorithmic_bytes": step_bytes,
                "step_frac_of_peak": step_bytes / (ms_total / K * 1e-3) / 1e9 / peak}

    # ---- end to end through the public, reference-shaped API with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, step, dev, world, min(K, 50))

    value = world * B * K / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": K, "warmup": Wu,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, synth),
        "run": {"cuda_graph": bool(graphed), "concurrent_chains": not args.serial, "schedule": step.schedule,
                "prototype_exchange": (None if world == 1 else
                                       "fused all-reduce + EMA kernel over NVLink peer memory (c3d_proto_ema_apply_peers)"
                                       if step.peer is not None else "ncclAllReduce + c3d_proto_ema_apply"),
                "fill_bytes_by_carrier": step.fill_bytes_by_carrier(), "vote_after_loss_rows": step.knn_after_rows,
                "blocks": len(block_ms), "block_ms_min_median_max": [min(block_ms), ms_total, max(block_ms)],
                "timed_region_s": t_region1 - t_region0, "banks_identical_across_ranks": banks_identical,
                "kernels_per_step": eager_launches_per_step},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "kernels_us": {k: round(v["us"], 2) for k, v in sorted(per_kernel.items())},
        "e2e": e2e,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, load_synth())
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: NCCL communicators referenced by captured CUDA graphs can make
        # destroy_process_group() block; all results are out, so drop the graphs,
        # synchronise with the peers and leave without the collective tear-down.
        step.graphs = None
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_e2e(args, step, dev, world, K):
    """Same step through the reference-shaped classes, inputs starting in pinned
    host memory every step (raw points, CSR offsets, per-point weak labels) and
    results read back to the host (loss scalar, per-point KNN labels).  The CNN
    activations (features / probabilities / argmax) are produced on the device in
    the reference too (trainer.py:625-638), so they stay resident.  The projected weak
    label image is built on the device by the fused projection (8f-1), so the labels
    travel per point."""
    import torch.distributed as dist
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    from coarse3d_b200.pc_processor.loss import ContrastMEMLoss
    from coarse3d_b200.pc_processor.models import PrototypeBank
    from coarse3d_b200.pc_processor.postproc import KNN
    shp, B = step.shape, step.batch
    H, W, C = shp.proj_h, shp.proj_w, shp.n_classes
    rp = RangeProjection(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W, device=dev)
    crit = ContrastMEMLoss(ignore_label=0, temperature=0.07, num_anchor=512)
    bank = PrototypeBank(C, step.M, step.dim, proto_mom=0.999).to(dev)
    with torch.no_grad():
        bank.prototypes.copy_(step.protos)     # one bank value on every rank (a replicated parameter)
    knn = KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), C)
    # per-point labels travel as uint8 class ids (C <= 255), a quarter of the loaders' int32
    label_np = np.int32 if os.environ.get("C3D_E2E_LABELS", "u8") == "i32" else np.uint8
    label_t = torch.int32 if label_np is np.int32 else torch.uint8
    host = []
    for s in step.sets:
        host.append(dict(points=torch.from_numpy(s.host_points).pin_memory(),
                         offsets=torch.from_numpy(s.host_offsets).pin_memory(),
                         weak=torch.from_numpy(s.host_weak.astype(label_np)).pin_memory()))
    # double-buffered device inputs / pinned outputs: the H2D copies of step i+1 (copy-in
    # stream) and the D2H of step i-1 (copy-out stream) overlap the compute of step i
    NB = min(int(os.environ.get("C3D_E2E_BUFFERS", "3")), len(step.sets))
    d_in = [dict(points=torch.empty_like(step.sets[0].points), offsets=torch.empty_like(step.sets[0].offsets),
                 weak=torch.empty((step.n_points,), dtype=label_t, device=dev))
            for _ in range(NB)]
    # KNN labels go back as the reference's int64 by default; C3D_E2E_KNN=u8 takes the opt-in uint8
    knn_u8 = os.environ.get("C3D_E2E_KNN", "i64") == "u8"
    knn_t = torch.uint8 if knn_u8 else torch.int64
    h_out = [dict(loss=torch.zeros((), dtype=torch.float32).pin_memory(),
                  knn=torch.zeros((step.n_points,), dtype=knn_t).pin_memory()) for _ in range(NB)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_done = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]
    main = torch.cuda.current_stream(dev)
    for e in ev_done + ev_out:
        e.record(main)
    h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in host[0])
    d2h = 4 + step.n_points * (1 if knn_u8 else 8)

    s0 = step.sets[0]
    d_loss = [torch.zeros((), device=dev) for _ in range(NB)]
    d_knn = [torch.zeros((step.n_points,), dtype=knn_t, device=dev) for _ in range(NB)]

    use_classes = os.environ.get("C3D_E2E_API", "classes") == "classes"   # "pipeline": HotPathStep.run_inputs

    def compute(j):
        """One step on the device inputs of buffer j."""
        di = d_in[j]
        if not use_classes:
            # the package's step pipeline: the same C-ABI calls, chains on streams
            loss, lab, _ = step.run_inputs(di["points"], di["offsets"], di["weak"], step.proj_bufs[j], j)
            d_loss[j].copy_(loss)
            d_knn[j].copy_(lab)
            return
        # the reference-shaped classes called one after the other (drop-in form)
        # projection fused with label-image assembly: the weak labels travel per point
        pr = rp.doProjectionAssembleBatch(di["points"], di["offsets"], weak_label=di["weak"],
                                          buffers=step.proj_bufs[j])
        labels = pr.train_label
        bank.update(s0.feats.detach(), labels)           # model.forward's prototype block comes first
        feats = s0.feats.requires_grad_(True)
        feats.grad = None
        loss = crit(feats=feats, output=s0.probs, labels=labels, keep_mask=None,
                    proto_queue=bank.prototypes.detach().unsqueeze(0))
        loss.backward()
        lab = knn.forward_batch(pr.proj_range, pr.uproj_depth, s0.argmax, pr.uproj_x_idx,
                                pr.uproj_y_idx, di["offsets"], out_uint8=knn_u8)
        d_loss[j].copy_(loss.detach())
        d_knn[j].copy_(lab)

    # The compute part (the same public-API calls) is captured in one CUDA graph per
    # buffer, as a user would do for a fixed-shape training step; falls back to eager.
    graphs = None
    for j in range(NB):                      # valid inputs before any warm-up / capture
        for k in ("points", "offsets", "weak"):
            d_in[j][k].copy_(host[0][k])
    torch.cuda.synchronize(dev)
    if not args.no_graph:   # with N > 1 the captured step contains the NCCL all-reduce, like HotPathStep's
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for j in range(NB):
                    compute(j)
            main.wait_stream(side)
            torch.cuda.synchronize(dev)
            graphs = []
            for j in range(NB):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    compute(j)
                graphs.append(g)
        except Exception as e:  # noqa: BLE001
            graphs = None
            torch.cuda.synchronize(dev)
            sys.stderr.write("e2e graph capture failed, running eager: %r\n" % (e,))

    skip = set(os.environ.get("C3D_E2E_SKIP", "").split(","))   # diagnosis only: h2d,compute,d2h

    def one(i):
        h, j = host[i % len(host)], i % NB
        di, ho = d_in[j], h_out[j]
        s_in.wait_event(ev_done[j])          # buffer j free again (compute of step i-NB done)
        with torch.cuda.stream(s_in):
            if "h2d" not in skip:
                for k in ("points", "offsets", "weak"):
                    di[k].copy_(h[k], non_blocking=True)
            ev_in[j].record(s_in)
        main.wait_event(ev_in[j])
        main.wait_event(ev_out[j])           # outputs of step i-NB have left the device
        if "compute" in skip:
            pass
        elif graphs is not None:
            graphs[j].replay()
        else:
            compute(j)
        ev_done[j].record(main)
        s_out.wait_event(ev_done[j])
        with torch.cuda.stream(s_out):
            if "d2h" not in skip:
                ho["loss"].copy_(d_loss[j], non_blocking=True)
                ho["knn"].copy_(d_knn[j], non_blocking=True)
            ev_out[j].record(s_out)

    for i in range(3):
        one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    # blocks of K steps (max over ranks per block), repeated for >= 1 s; the MEDIAN block is reported
    # (a single 50 ms block right after start-up was seen 30 % off on a fresh box)
    blocks, n_blocks = [], 1
    while True:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for i in range(K):
            one(i)
        main.wait_stream(s_out)
        e1.record(main)
        torch.cuda.synchronize(dev)
        bms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bms, op=dist.ReduceOp.MAX)
        blocks.append(float(bms.item()))
        if len(blocks) == 1:
            n_blocks = int(min(200, max(1, math.ceil(min(args.min_seconds, 1.0) * 1e3 / max(blocks[0], 1e-3)))))
        if len(blocks) >= n_blocks:
            break
    ms = torch.tensor([float(np.median(blocks))], dtype=torch.float64)
    for s in step.sets:
        s.feats.requires_grad_(False)
    return {"value": world * B * K / (float(ms.item()) * 1e-3), "unit": "scans/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": K, "blocks": len(blocks),
            "ms_per_step": float(ms.item()) / K,
            "api": ("RangeProjection.doProjectionAssembleBatch + PrototypeBank.update + "
                    "ContrastMEMLoss()(..).backward() + KNN.forward_batch (classes called in sequence)") if use_classes
                   else "coarse3d_b200.pipeline.HotPathStep.run_inputs: c3d_project_assemble_batch -> "
                        "c3d_knn_batch(+fill) || c3d_proto_loss_forward/backward || c3d_proto_ema_* on streams",
            "host_inputs": "points f32 (N,4), offsets, per-point weak labels %s (pinned); " % label_t +
                           "CNN activations resident on device as in the reference",
            "pipelining": "%d-deep buffered: H2D / compute / D2H of consecutive steps on three streams" % NB,
            "knn_labels_to_host": str(knn_t), "h2d_gbs": h2d * K / (float(ms.item()) * 1e-3) / 1e9,
            "bound": "host->device link: %.1f MB per step at the PCIe Gen5 x16 rate (~52 GB/s measured, "
                     "tools/pcie_probe.py) is %.2f ms" % (h2d / 1e6, h2d / 52e9 * 1e3),
            "compute_graph": graphs is not None}


if __name__ == "__main__":
    main()
