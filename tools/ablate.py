#!/usr/bin/env python
"""Developer tool: graph-replay time of the step with subsets of the chains, to see
which chain bounds the step (events cannot be placed inside a captured graph)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth
from coarse3d_b200.pipeline import HotPathStep

ALL = ["proj", "knn", "fill", "loss", "ema"]
SETS = [ALL, ["proj"], ["proj", "knn"], ["fill"], ["proj", "knn", "fill"], ["loss"], ["ema"],
        ["loss", "fill"], ["loss", "ema"], ["proj", "knn", "loss", "ema"], ["proj", "fill", "loss", "ema"]]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:      # e.g. "proj+knn+fill,loss+ema"
    SETS = [t.split("+") for t in sys.argv[2].split(",")]
for parts in SETS:
    step = HotPathStep(synth.KITTI, B, parts=parts, n_sets=2)
    for i in range(4):
        step.run(i)
    torch.cuda.synchronize()
    assert step.capture(), getattr(step, "capture_error", "")
    for i in range(10):
        step.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        step.step(i)
    e1.record()
    torch.cuda.synchronize()
    print("%-34s %7.1f us/step" % ("+".join(parts), e0.elapsed_time(e1) * 1000 / 200), flush=True)
    del step
    torch.cuda.empty_cache()
