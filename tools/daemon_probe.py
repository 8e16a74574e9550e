#!/usr/bin/env python
"""Developer tool: why / whether the fill daemon slows down next to the other kernels.
Graph-replay time of chain subsets with the daemon in its wait / priority variants."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth  # noqa: E402
from coarse3d_b200.pipeline import HotPathStep  # noqa: E402


def timed(step, n=60):
    for i in range(3):
        step.run(i, seed=i)
    torch.cuda.synchronize()
    ok = step.capture()
    for i in range(6):
        step.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step.step(i)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1000 / n, 1), bool(ok)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    step = HotPathStep(synth.KITTI, B, n_sets=3)
    ALL = ["proj", "knn", "fill", "loss", "ema"]
    for parts in (["fill"], ["proj", "knn"], ["loss", "ema"], ["fill", "knn"], ["fill", "proj"], ["fill", "proj", "knn"],
                  ["fill", "loss", "ema"], ALL):
        for daemon in ((0, 1, 8192, 4), (10, 1, 8192, 4), (20, 1, 8192, 4), (20, 1, 16384, 4), (20, 2, 8192, 4),
                       (1, 4, 0, 0)):
            for prio in (-1, 0):
                if "fill" not in parts and (daemon != (0, 1, 8192, 4) or prio != -1):
                    continue
                step.set_schedule("fill_daemon", daemon, parts=parts, fill_priority=prio)
                us, ok = timed(step)
                print(json.dumps(dict(batch=B, parts="+".join(parts), daemon=daemon, fill_prio=prio, us=us, graph=ok)),
                      flush=True)


if __name__ == "__main__":
    main()
