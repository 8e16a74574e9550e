#!/usr/bin/env python
"""Developer tool: pinned host <-> device copy bandwidth of the box (what bounds bench.py's e2e)."""
import torch

dev = torch.device("cuda:0")


def bw(nbytes, direction, iters=20, both=False):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(iters):
        with torch.cuda.stream(s1):
            if direction == "h2d":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


for mb in (1, 4, 16, 64):
    n = mb << 20
    print("%3d MiB  h2d %.1f GB/s  d2h %.1f GB/s  h2d while d2h %.1f GB/s" % (
        mb, bw(n, "h2d"), bw(n, "d2h"), bw(n, "h2d", both=True)))
