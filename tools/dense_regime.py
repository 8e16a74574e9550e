#!/usr/bin/env python
"""Developer tool: loss fwd+bwd and EMA timings when labels are dense (the reference's
entropy_selection regime: pseudo-labels on up to 50 % of the pixels)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coarse3d_b200 import ops

B, D, H, W, C, M, A = 8, 128, 64, 2048, 20, 20, 512
g = torch.Generator(device="cuda").manual_seed(0)
feats = torch.randn(B, D, H, W, device="cuda", generator=g)
probs = torch.softmax(torch.randn(B, C, H, W, device="cuda", generator=g), 1)
labels = torch.randint(1, C, (B, H, W), device="cuda", generator=g)
queue = torch.nn.functional.normalize(torch.randn(C, M, D, device="cuda", generator=g), dim=-1)
ln = [torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")]
cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
for frac in (1e-3, 0.05, 0.3):
    keep = torch.rand(B, H, W, device="cuda", generator=g) < frac
    ws = ops.proto_loss_workspace(B, C, H * W, D, M, A, "cuda")
    loss = torch.zeros((), device="cuda"); go = torch.ones((), device="cuda"); grad = torch.empty_like(feats)
    for tc in (False, True):
        def run():
            ops.proto_loss_forward_raw(feats, probs, labels, keep, queue, cfg, None, 1, ws, loss, tensor_cores=tc)
            ops.proto_loss_backward_raw(feats.shape, cfg, C, M, ws, go, grad)
        for _ in range(3): run()
        torch.cuda.synchronize()
        with ops.profile("") as prof:
            for _ in range(5): run()
            torch.cuda.synchronize()
            k = {n: round(1e3 * v[0] / v[1], 1) for n, v in prof.all().items()}
        T, nlab, flags = ops.proto_loss_info(ws)
        rows = int(ws[20:24].view(torch.int32).item())      # info[5]: distinct sampled rows
        print("loss  frac=%g tensor_cores=%s labelled=%d segments=%d rows=%d flags=%d loss=%.7f  us: %s" % (
            frac, tc, nlab, T, rows, flags, float(loss), k))
    lab_ema = (labels * keep).contiguous()
    mr = min(B * H * W, int(nlab * 1.1) + 1024)
    def run2():
        acc = ops.proto_ema_accumulate(feats, lab_ema, queue, *ln, assign_mode=ops.ASSIGN_GUMBEL_DEVICE, seed=1, max_rows=mr)
        ops.proto_ema_apply(queue, acc.packed, 0.999)
        return acc
    for _ in range(2): acc = run2()
    torch.cuda.synchronize()
    with ops.profile("") as prof:
        for _ in range(3): acc = run2()
        torch.cuda.synchronize()
        k = {n: round(1e3 * v[0] / v[1], 1) for n, v in prof.all().items()}
    print("ema   frac=%g rows=%d flags=%d  us: %s" % (frac, nlab, ops.proto_ema_info(acc.workspace)[2], k))
