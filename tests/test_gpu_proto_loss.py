"""GPU parity: c3d_proto_loss_forward/backward (through the reference-shaped
ContrastMEMLoss module) against the CPU oracle and the reference golden vectors.

Tolerances (BASELINE.json north_star): loss <= 1e-5 relative; gradients <= 1e-4 relative,
ELEMENT-WISE: |diff| <= 1e-4 * |ref| + 5e-6 * max|ref| (the absolute term covers elements that
are small only through cancellation of much larger terms)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import proto_loss as oloss

pytestmark = pytest.mark.gpu

LOSS = load_golden("proto_loss")
LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


def _rel(a, b):
    """Largest element-wise violation of |a-b| <= GRAD_RTOL*|b| + 5e-6*max|b|, scaled so that
    a value <= GRAD_RTOL passes (keeps the call sites' `_rel(..) <= GRAD_RTOL` form)."""
    bound = GRAD_RTOL * b.abs() + 5e-6 * b.abs().max()
    return float(((a - b).abs() / bound).max()) * GRAD_RTOL


def _module(A, temperature=0.07, **kw):
    from coarse3d_b200.pc_processor.loss import ContrastMEMLoss
    return ContrastMEMLoss(ignore_label=0, temperature=temperature, num_anchor=A, **kw)


@pytest.mark.parametrize("case", sorted(LOSS))
def test_module_matches_reference_golden(cuda_device, case):
    g = LOSS[case]
    feats = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    crit = _module(int(g["num_anchor"]), float(g["temperature"]), is_debug=True)
    loss = crit(feats=feats, output=torch.from_numpy(g["output"]).cuda(),
                labels=torch.from_numpy(g["labels"]).cuda(),
                keep_mask=torch.from_numpy(g["keep_mask"]).cuda(),
                proto_queue=torch.from_numpy(g["queue"]).cuda(),
                keep=torch.from_numpy(g["keep"]).cuda())
    assert loss.dim() == 0 and loss.is_cuda
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    gref = torch.from_numpy(g["grad"])
    got = feats.grad.cpu()
    assert _rel(got, gref) <= GRAD_RTOL
    assert torch.equal(got != 0, gref != 0)  # same sparsity pattern, dense elsewhere zero


def _random_problem(B, D, H, W, C, M, frac, seed, normalized_bank=True):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, D, H, W, generator=g)
    output = torch.softmax(torch.randn(B, C, H, W, generator=g) * 2, 1)
    labels = torch.randint(0, C, (B, H, W), generator=g)
    keep_mask = torch.rand(B, H, W, generator=g) < frac
    queue = torch.randn(1, C, M, D, generator=g)
    if normalized_bank:
        queue = torch.nn.functional.normalize(queue, dim=-1)
    return feats, output, labels, keep_mask, queue


@pytest.mark.parametrize("B,D,H,W,C,M,A,frac", [
    (2, 128, 16, 512, 20, 20, 512, 0.002),   # KITTI-like weak labels, one bank tile
    (1, 256, 8, 256, 20, 20, 128, 0.01),     # D=256: bank streamed in two smem tiles
    (3, 64, 8, 128, 17, 20, 64, 0.3),        # dense labels (pseudo-label regime)
    (2, 32, 4, 100, 14, 7, 33, 0.05),        # odd sizes, W not a multiple of 4
    (1, 512, 4, 64, 6, 4, 16, 0.1),          # wide features: 4 chunks per thread
    (1, 24, 4, 64, 20, 20, 16, 0.2),         # D not a multiple of 32: padded (unswizzled) bank
])
def test_matches_oracle_with_injected_anchors(cuda_device, B, D, H, W, C, M, A, frac):
    feats, output, labels, keep_mask, queue = _random_problem(B, D, H, W, C, M, frac, 7)
    f_cpu = feats.clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(3)
    want, keep, segs = oloss.contrast_mem_loss(f_cpu, output, labels, keep_mask, queue, keep=None,
                                               temperature=0.07, num_anchor=A, generator=gen)
    (want * 0.1).backward()
    f_gpu = feats.cuda().requires_grad_(True)
    crit = _module(A, is_debug=True)
    loss = crit(feats=f_gpu, output=output.cuda(), labels=labels.cuda(), keep_mask=keep_mask.cuda(),
                proto_queue=queue.cuda(), keep=keep.cuda())
    (loss * 0.1).backward()  # trainer.py:688-690 scales by loss_w_contrast
    assert abs(loss.item() - want.item()) <= LOSS_RTOL * abs(want.item())
    assert _rel(f_gpu.grad.cpu(), f_cpu.grad) <= GRAD_RTOL
    assert torch.equal(f_gpu.grad.cpu() != 0, f_cpu.grad != 0)


def test_device_sampler_is_consistent_and_reproducible(cuda_device):
    """Device-drawn anchors: multiplicities sum to A per segment, land only on kept
    pixels of the segment's class, are reproducible for a seed, and the loss /
    gradient equal the oracle's when it is fed the same anchors."""
    from coarse3d_b200 import ops
    B, D, H, W, C, M, A = 2, 32, 8, 128, 9, 5, 256
    feats, output, labels, keep_mask, queue = _random_problem(B, D, H, W, C, M, 0.1, 11)
    f_gpu = feats.cuda().requires_grad_(True)
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
    args = (output.cuda(), labels.cuda(), keep_mask.cuda(), queue[0].cuda(), cfg)
    loss, ws = ops.proto_loss(f_gpu, *args, seed=1234)
    loss.backward()
    pix, cls, cnt = (t.cpu().long() for t in ops.proto_loss_rows(ws, B, D, H * W, C, M, A))
    lab = oloss.masked_labels(labels, keep_mask, 0).view(-1)
    assert torch.equal(lab[pix], cls) and (cls != 0).all()
    assert pix.numel() == int((lab != 0).sum())
    segs = oloss.segments(lab.view(B, -1), 0)
    keep = []
    for b, c in segs:
        sel = (pix // (H * W) == b) & (cls == c)
        assert int(cnt[sel].sum()) == A
        assert torch.equal(pix[sel], torch.sort(pix[sel])[0])
        keep.append(torch.repeat_interleave(pix[sel] - b * H * W, cnt[sel]))
    keep = torch.stack(keep)
    f_cpu = feats.clone().requires_grad_(True)
    want, _, _ = oloss.contrast_mem_loss(f_cpu, output, labels, keep_mask, queue, keep=keep,
                                         temperature=0.07, num_anchor=A)
    want.backward()
    assert abs(loss.item() - want.item()) <= LOSS_RTOL * abs(want.item())
    assert _rel(f_gpu.grad.cpu(), f_cpu.grad) <= GRAD_RTOL
    loss2, ws2 = ops.proto_loss(feats.cuda(), *args, seed=1234)
    assert loss2.item() == loss.item()  # bitwise reproducible
    assert torch.equal(ops.proto_loss_rows(ws2, B, D, H * W, C, M, A)[2].cpu().long(), cnt)
    loss3, ws3 = ops.proto_loss(feats.cuda(), *args, seed=99)
    assert not torch.equal(ops.proto_loss_rows(ws3, B, D, H * W, C, M, A)[2].cpu().long(), cnt)


def test_device_sampler_follows_entropy_weights(cuda_device):
    """One segment, many draws: empirical frequencies match w / sum(w) (:46-49,:112-116)."""
    from coarse3d_b200 import ops
    B, D, H, W, C, M, A = 1, 8, 4, 64, 3, 2, 200000
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, D, H, W, generator=g)
    output = torch.softmax(torch.randn(B, C, H, W, generator=g) * 3, 1)
    labels = torch.ones(B, H, W, dtype=torch.long)
    keep_mask = torch.rand(B, H, W, generator=g) < 0.2
    queue = torch.randn(C, M, D, generator=g)
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
    _, ws = ops.proto_loss(feats.cuda(), output.cuda(), labels.cuda(), keep_mask.cuda(), queue.cuda(),
                           cfg, seed=42)
    pix, _, cnt = (t.cpu().long() for t in ops.proto_loss_rows(ws, B, D, H * W, C, M, A))
    w = oloss.entropy_weights(output).view(-1)[pix]
    p = (w / w.sum()).double()
    freq = cnt.double() / A
    sigma = torch.sqrt(p * (1 - p) / A)
    assert ((freq - p).abs() <= 5 * sigma + 1e-6).all()


def test_no_labelled_pixel_is_loud(cuda_device):
    feats, output, labels, keep_mask, queue = _random_problem(1, 16, 4, 32, 5, 3, 0.5, 3)
    keep_mask[:] = False
    crit = _module(8)
    loss = crit(feats=feats.cuda(), output=output.cuda(), labels=labels.cuda(),
                keep_mask=keep_mask.cuda(), proto_queue=queue.cuda())
    assert torch.isnan(loss).item()
    with pytest.raises(AssertionError):
        _module(8, is_debug=True)(feats=feats.cuda(), output=output.cuda(), labels=labels.cuda(),
                                  keep_mask=keep_mask.cuda(), proto_queue=queue.cuda())
    with pytest.raises(AssertionError):
        crit(feats=feats.cuda(), output=output.cuda(), labels=labels.cuda(), keep_mask=None,
             proto_queue=None)


def test_bad_injected_anchor_is_flagged(cuda_device):
    from coarse3d_b200 import ops
    feats, output, labels, keep_mask, queue = _random_problem(1, 16, 4, 32, 5, 3, 0.5, 4)
    lab = oloss.masked_labels(labels, keep_mask, 0).view(1, -1)
    segs = oloss.segments(lab, 0)
    keep = torch.zeros((len(segs), 8), dtype=torch.long)  # pixel 0 is not in every segment
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, 8)
    _, ws = ops.proto_loss(feats.cuda(), output.cuda(), labels.cuda(), keep_mask.cuda(),
                           queue[0].cuda(), cfg, keep=keep.cuda())
    assert ops.proto_loss_info(ws)[2] & ops.FLAG_BAD_KEEP


def test_full_size_config2_properties(cuda_device):
    """BASELINE config 2 (8 KITTI-shaped scans, D=128, C=20, M=20, A=512, 0.1 % labels):
    bitwise run-to-run determinism, gradient support == sampled pixels, and parity
    with the oracle on one scan of the batch."""
    from coarse3d_b200 import ops
    B, D, H, W, C, M, A = 8, 128, 64, 2048, 20, 20, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn(B, D, H, W, device="cuda", generator=g)
    output = torch.softmax(torch.randn(B, C, H, W, device="cuda", generator=g), 1)
    labels = torch.randint(1, C, (B, H, W), device="cuda", generator=g)
    keep_mask = torch.rand(B, H, W, device="cuda", generator=g) < 1e-3
    queue = torch.nn.functional.normalize(torch.randn(C, M, D, device="cuda", generator=g), dim=-1)
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
    outs = []
    for _ in range(2):
        f = feats.clone().requires_grad_(True)
        loss, ws = ops.proto_loss(f, output, labels, keep_mask, queue, cfg, seed=7)
        loss.backward()
        outs.append((loss.detach().clone(), f.grad))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    pix, cls, cnt = ops.proto_loss_rows(ws, B, D, H * W, C, M, A)
    support = torch.zeros(B * H * W, dtype=torch.bool, device="cuda")
    support[pix[cnt > 0].long()] = True
    gsup = (outs[0][1] != 0).any(dim=1).view(-1)
    assert torch.equal(gsup, support)
    # oracle on scan 0 alone with the anchors the device drew for scan 0
    sel0 = (pix.long() < H * W)
    segs0 = torch.unique(cls[sel0]).tolist()
    keep0 = torch.stack([torch.repeat_interleave(pix[sel0 & (cls == c)].long(), cnt[sel0 & (cls == c)].long())
                         for c in segs0]).cpu()
    f_cpu = feats[:1].cpu().requires_grad_(True)
    want, _, _ = oloss.contrast_mem_loss(f_cpu, output[:1].cpu(), labels[:1].cpu(), keep_mask[:1].cpu(),
                                         queue[None].cpu(), keep=keep0, temperature=0.07, num_anchor=A)
    want.backward()
    f0 = feats[:1].clone().requires_grad_(True)
    got, _ = ops.proto_loss(f0, output[:1].contiguous(), labels[:1].contiguous(),
                            keep_mask[:1].contiguous(), queue, cfg, keep=keep0.cuda())
    got.backward()
    assert abs(got.item() - want.item()) <= LOSS_RTOL * abs(want.item())
    assert _rel(f0.grad.cpu(), f_cpu.grad) <= GRAD_RTOL


@pytest.mark.parametrize("B,D,H,W,C,M,A,frac", [
    (2, 128, 16, 512, 20, 20, 512, 0.002),    # KITTI-like weak labels
    (2, 128, 16, 256, 20, 20, 64, 0.3),       # dense labels: many distinct rows per segment
    (1, 64, 8, 128, 7, 5, 32, 0.1),           # tile rows not a multiple of 8 (30 bank rows)
    (1, 256, 8, 128, 20, 20, 64, 0.05),       # D = 256: two bank tiles
    (2, 32, 8, 64, 5, 3, 16, 0.2),
])
def test_tensor_core_rows_match_oracle(cuda_device, B, D, H, W, C, M, A, frac):
    """need_grad bit 1: the two row x bank products on the tensor cores (mma.sync m16n8k8,
    3xTF32).  Same bars as the FFMA form: loss 1e-5 relative, gradients 1e-4 element-wise."""
    from coarse3d_b200 import ops
    feats, output, labels, keep_mask, queue = _random_problem(B, D, H, W, C, M, frac, 31)
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
    f_cpu = feats.clone().requires_grad_(True)
    want, keep, _ = oloss.contrast_mem_loss(f_cpu, output, labels, keep_mask, queue, temperature=0.07,
                                            num_anchor=A, generator=torch.Generator().manual_seed(3))
    want.backward()
    ws = ops.proto_loss_workspace(B, C, H * W, D, M, A, "cuda")
    dev = [t.cuda() for t in (feats, output, labels, keep_mask, queue[0])]
    outs = []
    for tc in (False, True):
        loss = torch.zeros((), device="cuda")
        ops.proto_loss_forward_raw(*dev, cfg, keep.cuda(), 0, ws, loss, tensor_cores=tc)
        grad = torch.empty_like(dev[0])
        ops.proto_loss_backward_raw(dev[0].shape, cfg, C, M, ws, torch.ones((), device="cuda"), grad)
        assert abs(loss.item() - want.item()) <= LOSS_RTOL * abs(want.item()), tc
        assert _rel(grad.cpu(), f_cpu.grad) <= GRAD_RTOL, tc
        assert torch.equal(grad.cpu() != 0, f_cpu.grad != 0)
        outs.append((loss.item(), grad))
    # the two forms agree far inside the bar
    assert abs(outs[0][0] - outs[1][0]) <= 2e-6 * abs(outs[0][0])
