#!/usr/bin/env python
"""Developer probe: zero fill by device-to-device memcpy from a small resident zero buffer
(copy engines instead of SMs?), alone and next to the KNN vote."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops, synth
from coarse3d_b200.pipeline import HotPathStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
step = HotPathStep(synth.KITTI, B, n_sets=1)
grad = step.grad.view(-1)
cudart = ctypes.CDLL("libcudart.so.12")
nbytes = grad.numel() * 4


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n


for zmb in (8, 32, 128):
    zeros = torch.zeros(zmb << 18, device="cuda")          # zmb MB of float32 zeros
    zb = zeros.numel() * 4

    def dma_fill(stream=None):
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream.cuda_stream
        off = 0
        while off < nbytes:
            n = min(zb, nbytes - off)
            cudart.cudaMemcpyAsync(ctypes.c_void_p(grad.data_ptr() + off), ctypes.c_void_p(zeros.data_ptr()),
                                   ctypes.c_size_t(n), 3, ctypes.c_void_p(st))
            off += n
    us = timeit(dma_fill)
    print(json.dumps(dict(batch=B, what="memcpy D2D fill alone", zero_mb=zmb, us=round(us, 1),
                          gbs=round(nbytes / us / 1e3))), flush=True)

s = step.sets[0]
pr = step._last_proj(step.proj_bufs[0])
C = step.shape.n_classes
knn_only = lambda: step._knn(s, pr, C)
print(json.dumps(dict(batch=B, what="knn alone", us=round(timeit(knn_only), 1))), flush=True)
print(json.dumps(dict(batch=B, what="knn carrying the fill (shipped)", us=round(timeit(lambda: step._knn(s, pr, C, cofill=step.grad)), 1))), flush=True)
side = torch.cuda.Stream()
zeros = torch.zeros(32 << 18, device="cuda")
zb = zeros.numel() * 4


def both():
    side.wait_stream(torch.cuda.current_stream())
    off = 0
    while off < nbytes:
        n = min(zb, nbytes - off)
        cudart.cudaMemcpyAsync(ctypes.c_void_p(grad.data_ptr() + off), ctypes.c_void_p(zeros.data_ptr()),
                               ctypes.c_size_t(n), 3, ctypes.c_void_p(side.cuda_stream))
        off += n
    knn_only()
    torch.cuda.current_stream().wait_stream(side)


print(json.dumps(dict(batch=B, what="knn || memcpy fill (two streams)", us=round(timeit(both), 1))), flush=True)
ms = lambda: cudart.cudaMemsetAsync(ctypes.c_void_p(grad.data_ptr()), 0, ctypes.c_size_t(nbytes),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
print(json.dumps(dict(batch=B, what="cudaMemsetAsync alone", us=round(timeit(ms), 1))), flush=True)
