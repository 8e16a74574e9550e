#!/usr/bin/env python
"""Developer tool: event timeline of one concurrent step (which kernels overlap)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops, synth
from coarse3d_b200.pipeline import HotPathStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
step = HotPathStep(synth.KITTI, B)
for i in range(5):
    step.run(i, seed=i)
torch.cuda.synchronize()
with ops.profile("") as prof:
    step.run(0); step.run(1)
    torch.cuda.synchronize()
    tl = prof.timeline()
half = tl[len(tl) // 2:]          # second step
t0 = min(a for _, a, _ in half)
for name, a, b in half:
    print("%-28s %8.1f -> %8.1f  (%6.1f us)" % (name, a - t0, b - t0, b - a))
