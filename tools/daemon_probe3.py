#!/usr/bin/env python
"""Developer tool: on which SMs the daemon's CTAs land in ONE step released all at once."""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import synth
from coarse3d_b200.pipeline import HotPathStep

B = 8
step = HotPathStep(synth.KITTI, B, n_sets=3)
for parts in (["fill"], ["fill", "loss", "ema"], ["fill", "proj"], ["proj", "knn", "fill", "loss", "ema"]):
    for daemon in ((2, 1, 1, 8192, 4), (2, 1, 4, 8192, 4), (2, 4, 4, 8192, 4)):
        step.set_schedule("fill_daemon", daemon, parts=parts)
        step.daemon_dbg = torch.zeros((148 * daemon[2], 4), dtype=torch.int64, device="cuda")
        for rep in range(3):
            step.daemon_dbg.fill_(-7)
            torch.cuda.synchronize()
            torch.cuda._sleep(4_000_000)
            step.run(rep, seed=rep)
            torch.cuda.synchronize()
            d = step.daemon_dbg.cpu()
            landed = d[d[:, 0] >= 0]
            per_sm = collections.Counter(landed[:, 0].tolist())
            workers = d[d[:, 3] > 0]
            life = (workers[:, 2] - workers[:, 1]).float() / 1e3
            t_first = workers[:, 1].min().item()
            starts = ((workers[:, 1] - t_first).float() / 1e3)
            print("parts=%s daemon=%s rep=%d: landed=%d on %d SMs (max %d per SM), workers=%d on %d SMs, "
                  "life us min/med/max=%.0f/%.0f/%.0f, start spread us max=%.0f, pages min/max=%d/%d" % (
                      "+".join(parts), daemon, rep, len(landed), len(per_sm), max(per_sm.values()),
                      len(workers), len(set(workers[:, 0].tolist())), life.min(), life.median(), life.max(),
                      starts.max(), workers[:, 3].min(), workers[:, 3].max()), flush=True)
