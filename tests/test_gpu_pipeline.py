"""GPU: the step pipeline (coarse3d_b200.pipeline.HotPathStep) -- schedules are only
schedules: every variant must produce the same loss, gradient, prototypes and KNN labels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _step(monkeypatch, schedule, concurrent=True, **kw):
    from coarse3d_b200 import synth
    from coarse3d_b200.pipeline import HotPathStep
    monkeypatch.setenv("C3D_SCHEDULE", schedule)
    return HotPathStep(synth.NUSCENES, 3, dim=32, sub_protos=4, num_anchor=16, n_sets=2, seed0=77,
                       concurrent=concurrent, **kw)


def _outputs(step):
    torch.cuda.synchronize()
    return (step.loss.clone(), step.grad.clone(), step.protos.clone(), step.knn_out.clone())


def test_schedules_agree(cuda_device, monkeypatch):
    ref = None
    for schedule, concurrent in [("fill_after_projection", False), ("fill_after_projection", True),
                                 ("fill_in_knn", False), ("fill_in_knn", True), ("fill_first", True),
                                 ("fill_daemon", True), ("fill_daemon:1,2,0,0", True),
                                 ("fill_daemon:0,1,4096,1", True), ("fill_spread", False),
                                 ("fill_spread", True), ("fill_spread=0.3,0.2,0.25,0.25", True),
                                 ("fill_spread=0,0,0,1", True), ("fill_spread=0.001,0.5,0,0", False)]:
        schedule, _, shares = schedule.partition("=")
        if shares:
            monkeypatch.setenv("C3D_FILL_SHARES", shares)
        else:
            monkeypatch.delenv("C3D_FILL_SHARES", raising=False)
        schedule, _, daemon = schedule.partition(":")
        if daemon:
            monkeypatch.setenv("C3D_DAEMON", daemon)
        else:
            monkeypatch.delenv("C3D_DAEMON", raising=False)
        step = _step(monkeypatch, schedule, concurrent)
        step.grad.fill_(7.0)                       # the step must overwrite every element
        step.run(0, seed=5)
        out = _outputs(step)
        if ref is None:
            ref = out
            assert torch.isfinite(out[0]) and int((out[1] != 0).sum()) > 0
        for a, b in zip(out, ref):
            assert torch.equal(a, b), (schedule, concurrent)


def test_vote_released_after_the_loss_rows(cuda_device, monkeypatch):
    """C3D_KNN_AFTER_ROWS=1 (the large-batch default): the vote and its share of the fill wait for
    the rows kernels of both chains -- an ordering only, same results."""
    monkeypatch.setenv("C3D_KNN_AFTER_ROWS", "0")
    step = _step(monkeypatch, "fill_in_knn")
    step.run(0, seed=5)
    want = _outputs(step)
    monkeypatch.setenv("C3D_KNN_AFTER_ROWS", "1")
    for schedule, shares, early in [("fill_in_knn", None, 0), ("fill_spread", "0.2,0,0.15,0.15", 0),
                                    ("fill_spread", "0,0,0,0.2", 1), ("fill_in_knn", None, 2)]:
        if shares:
            monkeypatch.setenv("C3D_FILL_SHARES", shares)
        monkeypatch.setenv("C3D_KNN_SPLIT", str(early))      # scans voted right after the projection
        monkeypatch.setenv("C3D_KNN_BINNED", "1" if early == 0 else "0")   # binned vote in two of the cases
        held = _step(monkeypatch, schedule)
        assert held.knn_after_rows and held.knn_split == early
        held.grad.fill_(2.0)
        held.knn_out.fill_(-1)
        held.run(0, seed=5)
        for a, b in zip(_outputs(held), want):
            assert torch.equal(a, b), schedule
        assert held.capture()
        held.protos.copy_(step.protos); held.bank_n.copy_(step.bank_n)


def test_two_phase_loss_forward_and_held_knn(cuda_device, monkeypatch):
    """c3d_proto_loss_forward_phase: selection and rows as two calls (with the KNN + fill kernel
    released in between, C3D_KNN_AFTER_SELECT=1) give the one-call results."""
    step = _step(monkeypatch, "fill_in_knn")
    step.run(0, seed=5)
    want = _outputs(step)
    monkeypatch.setenv("C3D_KNN_AFTER_SELECT", "1")
    held = _step(monkeypatch, "fill_in_knn")
    assert held.knn_after_select
    held.grad.fill_(2.0)
    held.run(0, seed=5)
    for a, b in zip(_outputs(held), want):
        assert torch.equal(a, b)


def test_two_launch_knn_matches(cuda_device, monkeypatch):
    """C3D_KNN_SPLIT: the vote (+ fill) as two launches over scans [0, b1) and [b1, B)."""
    step = _step(monkeypatch, "fill_in_knn")
    step.run(0, seed=5)
    want = _outputs(step)
    monkeypatch.setenv("C3D_KNN_SPLIT", "1")
    split = _step(monkeypatch, "fill_in_knn")
    assert split.knn_split == 1
    split.grad.fill_(2.0)
    split.knn_out.fill_(-1)
    split.run(0, seed=5)
    for a, b in zip(_outputs(split), want):
        assert torch.equal(a, b)


def test_graph_replay_matches_eager(cuda_device, monkeypatch):
    from coarse3d_b200 import ops
    step = _step(monkeypatch, "fill_in_knn")
    bank0 = step.protos.clone()
    step.run(1, seed=0)
    want = _outputs(step)
    assert step.capture(), getattr(step, "capture_error", "")
    step.grad.fill_(3.0)
    step.protos.copy_(bank0)                      # the bank evolves in place, step after step
    ops.bank_normalise(bank0, out=step.bank_n)
    step.seed_counters.zero_()                    # ... and so do the device-side step counters
    step.step(1)
    for a, b in zip(_outputs(step), want):
        assert torch.equal(a, b)


def test_run_inputs_equals_resident_step(cuda_device, monkeypatch):
    """Host-origin form (raw points + per-point weak labels through the fused projection) must
    give the resident-label step's results: the label image is the same image."""
    from coarse3d_b200 import ops
    step = _step(monkeypatch, "fill_in_knn")
    step.run(0, seed=9)
    want = _outputs(step)
    s = step.sets[0]
    weak = torch.from_numpy(s.host_weak.astype(np.int32)).cuda()
    bufs = ops.ProjectionBuffers(step.batch, step.n_points, 4, step.shape.proj_h, step.shape.proj_w, "cuda")
    step.grad.fill_(1.0)
    loss, lab, asm = step.run_inputs(s.points, s.offsets, weak, bufs, set_index=0, seed=9)
    got = _outputs(step)
    assert torch.equal(asm.train_label, s.labels)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_loss_reads_the_updated_bank(cuda_device, monkeypatch):
    """The EMA update precedes the loss (salsanext_proto.py:520-527 inside model.forward,
    trainer.py:675-686 after it): the step's loss must equal the loss of the UPDATED bank."""
    import dataclasses
    from coarse3d_b200 import ops, synth
    from coarse3d_b200.pipeline import HotPathStep
    monkeypatch.setenv("C3D_SCHEDULE", "fill_in_knn")
    # enough labels and a fast momentum, so that the bank visibly moves within one step
    shape = dataclasses.replace(synth.NUSCENES, label_ratio=2e-2)
    step = HotPathStep(shape, 3, dim=32, sub_protos=4, num_anchor=16, n_sets=2, seed0=77, momentum=0.5)
    bank0 = step.protos.clone()
    step.run(0, seed=5)
    loss, _, bank1, _ = _outputs(step)
    assert float((bank0 - bank1).abs().max()) > 1e-3
    s = step.sets[0]
    ws = ops.proto_loss_workspace(step.batch, step.shape.n_classes, step.shape.proj_h * step.shape.proj_w,
                                  step.dim, step.M, step.cfg.num_anchor, "cuda")
    out = torch.zeros((), device="cuda")
    ops.proto_loss_forward_raw(s.feats, s.probs, s.labels, s.keep_mask, bank1, step.cfg, None, 5, ws, out)
    assert torch.equal(out, loss)
    ops.proto_loss_forward_raw(s.feats, s.probs, s.labels, s.keep_mask, bank0, step.cfg, None, 5, ws, out)
    assert not torch.equal(out, loss)


def test_fused_step_equals_separate_operators(cuda_device, monkeypatch):
    """c3d_proto_step (one label split shared by the EMA update and the loss) must reproduce the
    two operators called one after the other, bit for bit."""
    outs = []
    monkeypatch.setenv("C3D_DEVICE_SEEDS", "0")     # the separate operators take `seed` alone
    for fused in ("0", "1"):
        monkeypatch.setenv("C3D_FUSED_STEP", fused)
        for concurrent in (False, True):
            step = _step(monkeypatch, "fill_in_knn", concurrent)
            assert step.fused_step == (fused == "1")
            step.grad.fill_(5.0)
            step.run(0, seed=3)
            step.run(1, seed=4)          # second step: the bank of the first one is the input
            outs.append(_outputs(step))
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert torch.equal(a, b)


def test_proto_step_phases_against_the_two_entry_points(cuda_device):
    from coarse3d_b200 import ops
    B, D, H, W, C, M, A = 2, 32, 16, 128, 7, 4, 64
    g = torch.Generator().manual_seed(21)
    feats = torch.randn(B, D, H, W, generator=g).cuda()
    probs = torch.softmax(torch.randn(B, C, H, W, generator=g), 1).cuda()
    labels = (torch.randint(1, C, (B, H, W), generator=g) * (torch.rand(B, H, W, generator=g) < 0.05)).cuda()
    protos = torch.nn.functional.normalize(torch.randn(C, M, D, generator=g), dim=-1).cuda()
    ln = [torch.ones(D).cuda(), torch.zeros(D).cuda(), torch.ones(C).cuda(), torch.zeros(C).cuda()]
    cfg = ops.ProtoLossConfig(0, 0.07, 0.07, A)
    # separate operators: EMA (argmax assignment), then the loss on the updated bank
    acc = ops.proto_ema_accumulate(feats, labels, protos, *ln, assign_mode=ops.ASSIGN_ARGMAX)
    bank1 = ops.proto_ema_apply(protos, acc.packed, 0.9)
    f1 = feats.clone().requires_grad_(True)
    want, _ = ops.proto_loss(f1, probs, labels, None, bank1, cfg, seed=17)
    want.backward()
    # fused
    max_rows = B * H * W
    ws = ops.proto_step_workspace(B, C, H * W, D, M, A, max_rows, "cuda")
    packed = torch.empty(C * M * D + C * M, device="cuda")
    loss = torch.zeros((), device="cuda")
    bank = protos.clone()
    common = dict(assign_mode=ops.ASSIGN_ARGMAX, seed=17)
    ops.proto_step_raw(ops.STEP_SPLIT | ops.STEP_SAMPLE | ops.STEP_ACCUMULATE, feats, probs, labels, None, bank,
                       *ln, cfg, ws, packed, loss, max_rows, **common)
    assert torch.equal(packed, acc.packed)
    ops.proto_ema_apply(bank, packed, 0.9, out=bank)
    assert torch.equal(bank, bank1)
    ops.proto_step_raw(ops.STEP_LOSS_ROWS, feats, probs, labels, None, bank, *ln, cfg, ws, packed, loss,
                       max_rows, **common)
    assert torch.equal(loss, want.detach())
    grad = torch.empty_like(feats)
    ops.proto_loss_backward_raw(feats.shape, cfg, C, M, ws, torch.ones((), device="cuda"), grad)
    assert torch.equal(grad, f1.grad)
    assert ops.proto_loss_info(ws)[1] == int((labels > 0).sum())


def test_device_step_counters_vary_the_draws_in_graph_replays(cuda_device, monkeypatch):
    """A captured graph bakes `seed` in; the device-side counters (advanced by the sampler and by
    the EMA apply) make every replay draw new anchors and new Gumbel noise."""
    import dataclasses
    from coarse3d_b200 import ops, synth
    from coarse3d_b200.pipeline import HotPathStep
    monkeypatch.setenv("C3D_SCHEDULE", "fill_in_knn")
    shape = dataclasses.replace(synth.NUSCENES, label_ratio=2e-2)   # many pixels per segment to draw from
    step = HotPathStep(shape, 3, dim=32, sub_protos=4, num_anchor=16, n_sets=2, seed0=77, momentum=0.5)
    assert step.capture(), getattr(step, "capture_error", "")
    step.seed_counters.zero_()
    bank0 = step.protos.clone()
    losses = []
    for _ in range(3):
        step.protos.copy_(bank0)
        ops.bank_normalise(bank0, out=step.bank_n)
        step.step(0)
        torch.cuda.synchronize()
        losses.append(float(step.loss))
    assert step.seed_counters.tolist() == [3, 3]
    assert len(set(losses)) == 3
    # same counters, same draws
    step.seed_counters.zero_()
    step.protos.copy_(bank0)
    ops.bank_normalise(bank0, out=step.bank_n)
    step.step(0)
    torch.cuda.synchronize()
    assert float(step.loss) == losses[0]
