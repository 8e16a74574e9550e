"""GPU parity: c3d_proto_ema_accumulate / c3d_proto_ema_apply (through PrototypeBank)
against the CPU oracle and the reference golden vectors.

Tolerance: prototypes are unit vectors; |diff| <= 1e-5 absolute (fp32 sums in a
different order than torch); proto_target and counts are exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import proto_ema as oema

pytestmark = pytest.mark.gpu

EMA = load_golden("proto_ema")
ATOL = 1e-5


def _bank(C, M, D, mom, protos0, ln, deterministic):
    from coarse3d_b200.pc_processor.models import PrototypeBank
    bank = PrototypeBank(C, M, D, ignore_label=0, proto_mom=mom, deterministic=deterministic).cuda()
    with torch.no_grad():
        bank.prototypes.copy_(protos0)
        bank.feat_norm.weight.copy_(ln[0]); bank.feat_norm.bias.copy_(ln[1])
        bank.mask_norm.weight.copy_(ln[2]); bank.mask_norm.bias.copy_(ln[3])
    return bank


def _gumbel_rows(g):
    rows = [torch.from_numpy(g["gumbel"][c, :int(n)]) for c, n in enumerate(g["n_per_class"]) if n]
    return torch.cat(rows, 0).contiguous()


@pytest.mark.parametrize("case", sorted(EMA))
def test_bank_update_matches_reference_golden(cuda_device, case):
    g = EMA[case]
    C, M, D = g["prototypes0"].shape
    ln = [torch.from_numpy(g[k]) for k in ("ln_d_w", "ln_d_b", "ln_c_w", "ln_c_b")]
    use_gumbel = bool(g["use_gumbel"])
    bank = _bank(C, M, D, float(g["momentum"]), torch.from_numpy(g["prototypes0"]), ln, not use_gumbel)
    gum = _gumbel_rows(g).cuda() if use_gumbel else None
    target = bank.update(torch.from_numpy(g["embedding"]).cuda(), torch.from_numpy(g["label"]).cuda(),
                         gumbel=gum, want_target=True)
    got = bank.prototypes.detach().cpu()
    want = torch.from_numpy(g["prototypes1"])
    assert isinstance(bank.prototypes, torch.nn.Parameter) and not bank.prototypes.requires_grad
    assert (got - want).abs().max() <= ATOL
    assert np.array_equal(target.cpu().numpy(), g["proto_target"].reshape(-1))
    # the bank actually moved (fixture exercises the EMA)
    assert (got - torch.nn.functional.normalize(torch.from_numpy(g["prototypes0"]), dim=-1)).abs().max() > 1e-3


def _problem(B, D, H, W, C, M, frac, seed):
    g = torch.Generator().manual_seed(seed)
    centers = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=-1)
    label = torch.randint(1, C, (B, H, W), generator=g)
    emb = torch.nn.functional.normalize(
        torch.randn(B, D, H, W, generator=g) * 0.7 + 1.5 * centers[label].permute(0, 3, 1, 2), dim=1)
    protos0 = torch.randn(C, M, D, generator=g) * 0.02 + centers[:, None, :] * 0.05
    label = label * (torch.rand(B, H, W, generator=g) < frac)
    ln = [1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g),
          1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)]
    return emb.contiguous(), label, protos0, ln


def _oracle(emb, label, protos0, ln, C, mom, gumbel_rows=None):
    gum = None
    if gumbel_rows is not None:
        gum, off = {}, 0
        flat = label.reshape(-1)
        for c in range(1, C):
            n = int((flat == c).sum())
            if n:
                gum[c] = gumbel_rows[off:off + n]
                off += n
    return oema.prototype_learning(emb, label, protos0, *ln, C, 0, mom, gumbel=gum, labelled_only=True)


@pytest.mark.parametrize("B,D,H,W,C,M,frac,use_gumbel", [
    (3, 128, 16, 256, 20, 20, 0.01, False),   # KITTI-like weak labels
    (3, 128, 16, 256, 20, 20, 0.01, True),
    (2, 256, 8, 128, 17, 20, 0.02, True),     # D=256: bank streamed in smem tiles
    (2, 64, 8, 100, 14, 7, 0.2, False),       # denser labels, odd sizes
    (2, 32, 16, 128, 6, 4, 0.5, True),        # ~400 rows per class: block Sinkhorn, Q in shared memory
    (1, 16, 32, 256, 3, 20, 0.9, False),      # ~3700 rows per class: Q in the global scratch
])
def test_matches_oracle(cuda_device, B, D, H, W, C, M, frac, use_gumbel):
    from coarse3d_b200 import ops
    emb, label, protos0, ln = _problem(B, D, H, W, C, M, frac, 5)
    n_rows = int((label != 0).sum())
    gum = None
    if use_gumbel:
        gum = -torch.empty(n_rows, M).exponential_(generator=torch.Generator().manual_seed(9)).log()
    want, sums, counts, target = _oracle(emb, label, protos0, ln, C, 0.9, gum)
    lnc = [t.cuda() for t in ln]
    acc = ops.proto_ema_accumulate(emb.cuda(), label.cuda(), protos0.cuda(), *lnc,
                                   gumbel=None if gum is None else gum.cuda(),
                                   assign_mode=None if use_gumbel else ops.ASSIGN_ARGMAX, want_target=True)
    got = ops.proto_ema_apply(protos0.cuda(), acc.packed, 0.9)
    segs, rows, flags = ops.proto_ema_info(acc.workspace)
    assert rows == n_rows and flags == 0
    K = C * M
    assert torch.equal(acc.packed[K * D:].cpu().view(C, M), counts)       # counts exact
    assert (acc.packed[:K * D].cpu().view(C, M, D) - sums).abs().max() <= 1e-4 * max(1.0, float(sums.abs().max()))
    assert np.array_equal(acc.proto_target.cpu().numpy(), target.numpy())
    assert counts.sum() > 0
    assert (got.cpu() - want).abs().max() <= ATOL


def test_two_rank_sum_equals_oracle_of_summed_payloads(cuda_device):
    """ranks all-reduce sums+counts, then apply one EMA.  Emulated on
    one GPU by accumulating two scan shards separately and adding the payloads."""
    from coarse3d_b200 import distributed, ops
    B, D, H, W, C, M = 4, 64, 8, 128, 12, 6
    emb, label, protos0, ln = _problem(B, D, H, W, C, M, 0.05, 8)
    lnc = [t.cuda() for t in ln]
    total, osum, ocnt = None, 0, 0
    for r in range(2):
        idx = list(distributed.shard_scans(B, r, 2))
        e, l = emb[idx].contiguous(), label[idx].contiguous()
        acc = ops.proto_ema_accumulate(e.cuda(), l.cuda(), protos0.cuda(), *lnc,
                                       assign_mode=ops.ASSIGN_ARGMAX)
        total = acc.packed.clone() if total is None else total + acc.packed
        _, s, c, _ = _oracle(e, l, protos0, ln, C, 0.9)
        osum, ocnt = osum + s, ocnt + c
    got = ops.proto_ema_apply(protos0.cuda(), total, 0.9)
    want = oema.ema_from_sums(oema.l2_normalize(protos0), osum, ocnt, 0.9)
    assert (got.cpu() - want).abs().max() <= ATOL


def test_device_gumbel_is_reproducible_and_overflow_is_flagged(cuda_device):
    from coarse3d_b200 import ops
    B, D, H, W, C, M = 2, 32, 8, 64, 8, 5
    emb, label, protos0, ln = _problem(B, D, H, W, C, M, 0.2, 6)
    args = (emb.cuda(), label.cuda(), protos0.cuda(), *[t.cuda() for t in ln])
    a = ops.proto_ema_accumulate(*args, seed=3).packed.clone()
    b = ops.proto_ema_accumulate(*args, seed=3).packed.clone()
    c = ops.proto_ema_accumulate(*args, seed=4).packed.clone()
    assert torch.equal(a, b) and not torch.equal(a, c)
    K = C * M
    det = ops.proto_ema_accumulate(*args, assign_mode=ops.ASSIGN_ARGMAX).packed
    assert a[K * D:].sum() == det[K * D:].sum()  # same masked rows, different assignment
    small = ops.proto_ema_accumulate(*args, assign_mode=ops.ASSIGN_ARGMAX, max_rows=4)
    assert ops.proto_ema_info(small.workspace)[2] & ops.EMA_FLAG_OVERFLOW
    assert float(small.packed.abs().sum()) == 0.0
    out = ops.proto_ema_apply(protos0.cuda(), small.packed, 0.9)
    assert (out.cpu() - oema.l2_normalize(protos0)).abs().max() <= 1e-6


def test_full_size_config2(cuda_device):
    """BASELINE config 2 shape (B=8, D=128, 64x2048, C=20, M=20, ~0.1 % labels) vs the oracle."""
    from coarse3d_b200 import ops
    B, D, H, W, C, M = 8, 128, 64, 2048, 20, 20
    emb, label, protos0, ln = _problem(B, D, H, W, C, M, 1e-3, 12)
    want, sums, counts, _ = _oracle(emb, label, protos0, ln, C, 0.999)
    acc = ops.proto_ema_accumulate(emb.cuda(), label.cuda(), protos0.cuda(), *[t.cuda() for t in ln],
                                   assign_mode=ops.ASSIGN_ARGMAX)
    got = ops.proto_ema_apply(protos0.cuda(), acc.packed, 0.999)
    K = C * M
    assert torch.equal(acc.packed[K * D:].cpu().view(C, M), counts) and counts.sum() > 100
    assert (got.cpu() - want).abs().max() <= ATOL
    assert torch.allclose(got.norm(dim=-1), torch.ones(C, M, device="cuda"), atol=1e-5)


def test_bank_surfaces_overflow_without_a_stall(cuda_device):
    """More labelled pixels than max_rows: the update is skipped (flag 16); the bank examines the
    flags of an update asynchronously and raises at the next call / on check_flags()."""
    from coarse3d_b200.pc_processor.models import PrototypeBank
    emb, label, protos0, ln = _problem(2, 32, 8, 64, 6, 4, 0.5, 9)
    bank = PrototypeBank(6, 4, 32, max_rows=4).cuda()
    bank.update(emb.cuda(), label.cuda())
    with pytest.raises(RuntimeError):
        bank.check_flags()
    ok = PrototypeBank(6, 4, 32).cuda()
    ok.update(emb.cuda(), label.cuda())
    ok.check_flags()


@pytest.mark.parametrize("C,M,D", [(20, 20, 128), (7, 4, 32), (14, 20, 256)])
def test_peer_exchange_kernel_on_one_rank(cuda_device, C, M, D):
    """c3d_proto_ema_apply_peers with world = 1 (the rank exchanges with itself through its own
    CUDA-IPC-exportable buffer): the fused all-reduce + EMA kernel must reproduce
    c3d_proto_ema_apply bit for bit, over several steps (both slot parities, the device step
    counter) and inside a captured CUDA graph.  The two-rank form is tests/test_gpu_multi.py."""
    from coarse3d_b200 import distributed, ops
    g = torch.Generator().manual_seed(C * 100 + D)
    bank = torch.nn.functional.normalize(torch.randn(C, M, D, generator=g), dim=-1).cuda()
    px = distributed.PeerExchange(C, M, D, "cuda")
    assert px.available, px.why_not
    want, cur = bank.clone(), bank.clone()
    bank_n = torch.empty_like(cur)
    K = C * M
    payloads = []
    for it in range(4):
        packed = torch.randn(K * D + K, generator=g)
        packed[K * D:] = torch.randint(0, 3, (K,), generator=g).float()        # counts, some zero
        packed[K * D + M:K * D + 2 * M] = 0                                     # a class without rows
        payloads.append(packed.cuda())
    for it, packed in enumerate(payloads):
        mine = packed.clone()
        px.apply(cur, mine, 0.9, out=cur, normalised_out=bank_n)
        want = ops.proto_ema_apply(want, packed, 0.9)
        assert torch.equal(mine, packed), "summed payload of a single rank must be the payload"
        assert torch.equal(cur, want), "step %d" % it
        assert torch.equal(bank_n, ops.bank_normalise(cur))
    # graph replay: the step counter lives on the device
    buf = payloads[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        px.apply(cur, buf, 0.9, out=cur)
    torch.cuda.current_stream().wait_stream(side)
    want = ops.proto_ema_apply(want, payloads[0], 0.9)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        px.apply(cur, buf, 0.9, out=cur)
    for it in range(3):
        buf.copy_(payloads[it + 1])
        graph.replay()
        want = ops.proto_ema_apply(want, payloads[it + 1], 0.9)
        assert torch.equal(cur, want), "graph replay %d" % it
    assert px.errors() == 0
    px.close()
