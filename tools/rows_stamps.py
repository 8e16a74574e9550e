# Needs a library built with -DC3D_DEBUG_STAMPS (add it to COMMON in coarse3d_b200/build.py).
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops, synth, _lib
from coarse3d_b200.pipeline import HotPathStep
step = HotPathStep(synth.KITTI, 8, n_sets=1)
s = step.sets[0]
for i in range(3):
    step._loss_fwd(s, i)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
_lib.lib.c3d_debug_rows_stamps.restype = ctypes.c_int
_lib.lib.c3d_debug_rows_stamps(buf)
v = list(buf)
names = ["start", "P0 loads issued", "bank staged+sync", "P1 logits done", "P2 softmax done", "P3 grad product done", "group loop done", "after last-CTA barrier"]
for i in range(1, 8):
    print("%-26s +%7d cycles" % (names[i], v[i] - v[i - 1]))
print("total", v[7] - v[0])
