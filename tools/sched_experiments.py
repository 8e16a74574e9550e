#!/usr/bin/env python
"""Developer tool: graph-replay time of the step under different schedules / fill-daemon
settings, one HotPathStep per batch size (inputs are built once).

    python tools/sched_experiments.py [batch ...]   ->  one line per variant
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth  # noqa: E402
from coarse3d_b200.pipeline import HotPathStep  # noqa: E402

# fill_spread: shares of the zero fill carried by (projection, label split, EMA rows, loss rows)
SPREAD = [(0.15, 0.05, 0.1, 0.1), (0.1, 0, 0, 0), (0.2, 0, 0, 0), (0.3, 0, 0, 0), (0, 0.05, 0, 0), (0, 0.1, 0, 0),
          (0, 0, 0.1, 0), (0, 0, 0.2, 0), (0, 0, 0, 0.1), (0, 0, 0, 0.2), (0.2, 0.05, 0.1, 0.1),
          (0.2, 0.1, 0.15, 0.15), (0.25, 0.1, 0.2, 0.2), (0.1, 0.05, 0.05, 0.05), (0.3, 0.1, 0.2, 0.3)]
VARIANTS = [
    ("fill_in_knn", None),
    ("fill_after_projection", None),
    ("fill_daemon", (0, 1, 8192, 4)),
    ("fill_daemon", (0, 1, 8192, 2)),
    ("fill_daemon", (0, 1, 8192, 8)),
    ("fill_daemon", (0, 1, 4096, 8)),
    ("fill_daemon", (0, 1, 4096, 16)),
    ("fill_daemon", (0, 1, 2048, 16)),
    ("fill_daemon", (0, 2, 4096, 4)),
    ("fill_daemon", (0, 2, 4096, 8)),
    ("fill_daemon", (1, 1, 0, 0)),
    ("fill_daemon", (1, 2, 0, 0)),
    ("fill_daemon", (1, 4, 0, 0)),
]


def time_steps(step, n):
    for i in range(6):
        step.step(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step.step(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / n)
    return best


def main():
    global VARIANTS
    if os.environ.get("C3D_SCHED_QUICK"):
        VARIANTS = VARIANTS[:2]
    if os.environ.get("C3D_SCHED_SPREAD"):
        spread = SPREAD
        if os.environ["C3D_SCHED_SPREAD"] != "1":     # "a,b,c,d;a,b,c,d;..."
            spread = [tuple(float(v) for v in t.split(",")) for t in os.environ["C3D_SCHED_SPREAD"].split(";")]
        VARIANTS = VARIANTS[:1] + [("fill_spread", sh) for sh in spread] + VARIANTS[:1]
    batches = [int(a) for a in sys.argv[1:]] or [8, 64]
    out = []
    for B in batches:
        step = HotPathStep(synth.KITTI, B, n_sets=3)
        for sched, daemon in VARIANTS:
            if sched == "fill_spread":     # shares; 5th value: 1 = vote after the loss rows; 6th: scans voted early
                step.set_schedule(sched, fill_shares=daemon[:4], knn_after_rows=len(daemon) > 4 and daemon[4] > 0,
                                  knn_split=int(daemon[5]) if len(daemon) > 5 else 0)
            else:
                step.set_schedule(sched, daemon, knn_after_rows=False, knn_split=0)
            for i in range(3):
                step.run(i, seed=i)
            torch.cuda.synchronize()
            ok = step.capture()
            us = time_steps(step, 100 if B <= 8 else 30)
            rec = dict(batch=B, schedule=sched, daemon=daemon, graph=bool(ok), us_per_step=round(us, 1))
            print(json.dumps(rec), flush=True)
            out.append(rec)
        # where the daemon's own time goes, alone on the GPU
        from coarse3d_b200 import ops
        for daemon in [] if (os.environ.get("C3D_SCHED_QUICK") or os.environ.get("C3D_SCHED_SPREAD")) else [(0, 1, 8192, 4), (0, 1, 4096, 16), (0, 2, 4096, 8), (1, 1, 0, 0), (1, 2, 0, 0)]:
            for _ in range(2):
                ops.zero_fill_background(step.grad, *daemon)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.zero_fill_background(step.grad, *daemon)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000 / 5
            gbs = step.grad.numel() * 4 / us / 1e3
            print(json.dumps(dict(batch=B, daemon_alone=daemon, us=round(us, 1), gbs=round(gbs, 1))), flush=True)
        del step
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
