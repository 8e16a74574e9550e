#!/usr/bin/env python
"""Developer tool: daemon with a lead hold, and the over-launched claim form, in graph replay."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth
from coarse3d_b200.pipeline import HotPathStep
from daemon_probe import timed

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
step = HotPathStep(synth.KITTI, B, n_sets=3)
ALL = ["proj", "knn", "fill", "loss", "ema"]
step.set_schedule("fill_in_knn", parts=ALL)
print(json.dumps(dict(batch=B, schedule="fill_in_knn", us=timed(step)[0])), flush=True)
for parts in (["fill", "proj", "knn"], ["fill", "loss", "ema"], ALL):
    for lead in (0, 2000, 4000, 8000):
        for daemon in ((0, 1, 8192, 4), (10, 1, 8192, 8), (0, 1, 16384, 4), (0, 2, 8192, 4), (2, 1, 32, 8192, 4), (2, 1, 128, 8192, 4),
                       (2, 2, 64, 8192, 4)):
            if daemon[0] == 2 and lead:
                continue
            step.daemon_lead_ns = lead
            step.set_schedule("fill_daemon", daemon, parts=parts)
            us, ok = timed(step)
            print(json.dumps(dict(batch=B, parts="+".join(parts), lead_ns=lead, daemon=daemon, us=us, graph=ok)), flush=True)
