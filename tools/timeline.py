#!/usr/bin/env python
"""Developer tool: event timeline of one concurrent step (which kernels overlap).  The step is
enqueued behind a device-side sleep, so that -- as in a CUDA-graph replay -- every chain's first
kernel is released at the same instant instead of being staggered by the host's launch latency."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops, synth
from coarse3d_b200.pipeline import HotPathStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
parts = sys.argv[2].split("+") if len(sys.argv) > 2 else None
step = HotPathStep(synth.KITTI, B, parts=parts)
for i in range(5):
    step.run(i, seed=i)
torch.cuda.synchronize()
with ops.profile("") as prof:
    torch.cuda._sleep(4_000_000)
    step.run(0)
    torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    step.run(1)
    torch.cuda.synchronize()
    tl = prof.timeline()
half = tl[len(tl) // 2:]          # second step
t0 = min(a for _, a, _ in half)
for name, a, b in sorted(half, key=lambda r: r[1]):
    print("%-28s %8.1f -> %8.1f  (%6.1f us)" % (name, a - t0, b - t0, b - a))
print("step span %.1f us" % (max(b for _, _, b in half) - t0))
