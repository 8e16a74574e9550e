#!/usr/bin/env python
"""One serial step (every kernel once, one stream) for `ncu`: 2 warm-up steps, then the profiled one.
usage: ncu ... python tools/ncu_step.py <batch> [config shape]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth
from coarse3d_b200.pipeline import HotPathStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shape = synth.SHAPES[sys.argv[2]] if len(sys.argv) > 2 else synth.KITTI
step = HotPathStep(shape, B, n_sets=1, concurrent=False)
for i in range(2):
    step.run(i, seed=i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.run(2, seed=2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
