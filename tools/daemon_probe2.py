#!/usr/bin/env python
"""Developer tool: where the daemon's CTAs land and how long they live, in graph replay next
to other grids (claim form, debug records)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth
from coarse3d_b200.pipeline import HotPathStep
from daemon_probe import timed

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
step = HotPathStep(synth.KITTI, B, n_sets=3)
ALL = ["proj", "knn", "fill", "loss", "ema"]
for parts in (["fill"], ["fill", "proj"], ["fill", "proj", "knn"], ["fill", "loss", "ema"], ALL):
    for daemon in ((2, 1, 4, 8192, 4), (2, 1, 8, 8192, 4), (2, 2, 8, 8192, 4), (2, 1, 4, 16384, 2), (2, 2, 8, 4096, 8),
                   (2, 4, 16, 4096, 4)):
        step.daemon_dbg = torch.zeros((148 * daemon[2], 4), dtype=torch.int64, device="cuda")
        step.set_schedule("fill_daemon", daemon, parts=parts)
        us, ok = timed(step)
        d = step.daemon_dbg.cpu()
        act = d[d[:, 3] > 0]
        sms = len(set(act[:, 0].tolist()))
        life = (act[:, 2] - act[:, 1]).float() / 1e3
        span = (act[:, 2].max() - act[:, 1].min()).item() / 1e3 if len(act) else 0
        print(json.dumps(dict(batch=B, parts="+".join(parts), daemon=daemon, us=us, workers=len(act), sms=sms,
                              life_us_min=round(life.min().item(), 1), life_us_max=round(life.max().item(), 1),
                              span_us=round(span, 1), pages_min=int(act[:, 3].min()), pages_max=int(act[:, 3].max()))),
              flush=True)
