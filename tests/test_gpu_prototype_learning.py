"""GPU parity: the drop-in `prototype_learning` (same signature and return as
SalsaNextProto.prototype_learning, salsanext_proto.py:337-402) through
c3d_proto_ema_accumulate_dense, against the reference golden vectors: the dense inputs are
built exactly as the reference's forward builds them (:497-510)."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu

EMA = load_golden("proto_ema")
ATOL = 1e-5


def _gumbel_rows(g):
    rows = [torch.from_numpy(g["gumbel"][c, :int(n)]) for c, n in enumerate(g["n_per_class"]) if n]
    return torch.cat(rows, 0).contiguous()


def _dense_inputs(g, dev):
    """salsanext_proto.py:497-510 with stock torch ops on the device."""
    emb = torch.from_numpy(g["embedding"]).to(dev)
    B, D, H, W = emb.shape
    C = g["prototypes0"].shape[0]
    ln = [torch.from_numpy(g[k]).to(dev) for k in ("ln_d_w", "ln_d_b", "ln_c_w", "ln_c_b")]
    out_feat = emb.permute(0, 2, 3, 1).reshape(-1, D)
    out_feat = F.normalize(F.layer_norm(out_feat, (D,), ln[0], ln[1]), p=2, dim=-1)
    protos = F.normalize(torch.from_numpy(g["prototypes0"]).to(dev), p=2, dim=-1)
    sim = torch.einsum("nd,kmd->nmk", out_feat, protos)
    nearest = F.layer_norm(torch.amax(sim, dim=1), (C,), ln[2], ln[3])
    nearest = nearest.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    return out_feat, protos, sim, nearest


@pytest.mark.parametrize("case", sorted(EMA))
def test_prototype_learning_dropin_matches_reference_golden(cuda_device, case):
    from coarse3d_b200.pc_processor.models import prototype_learning
    g = EMA[case]
    C, M, D = g["prototypes0"].shape
    out_feat, protos, sim, nearest = _dense_inputs(g, cuda_device)
    use_gumbel = bool(g["use_gumbel"])
    # any object with the attributes the reference's method reads works as `self`
    model = types.SimpleNamespace(prototypes=torch.nn.Parameter(protos.clone(), requires_grad=False),
                                  nclasses=C, ignore_label=0, sub_proto_size=M,
                                  proto_mom=float(g["momentum"]), deterministic=not use_gumbel)
    label = torch.from_numpy(g["label"]).to(cuda_device).view(-1)
    gum = _gumbel_rows(g).to(cuda_device) if use_gumbel else None
    logits, target = prototype_learning(model, out_feat, nearest, label, None, sim, gumbel=gum)
    assert logits.shape == (out_feat.shape[0], M * C) and logits.data_ptr() == sim.data_ptr()   # :343-345
    assert target.shape == label.shape and target.dtype == torch.float32                        # :346
    assert np.array_equal(target.cpu().numpy(), g["proto_target"].reshape(-1))
    assert isinstance(model.prototypes, torch.nn.Parameter) and not model.prototypes.requires_grad
    assert (model.prototypes.detach().cpu() - torch.from_numpy(g["prototypes1"])).abs().max() <= ATOL


def test_dense_and_labelled_only_paths_agree(cuda_device):
    """PrototypeBank.update (labelled rows gathered from the embedding) and the dense-argument
    method must produce the same packed sums, counts and targets."""
    from coarse3d_b200 import ops
    from coarse3d_b200.pc_processor.models import PrototypeBank
    g = EMA["weak_gumbel"]
    C, M, D = g["prototypes0"].shape
    out_feat, protos, sim, nearest = _dense_inputs(g, cuda_device)
    label = torch.from_numpy(g["label"]).to(cuda_device)
    gum = _gumbel_rows(g).to(cuda_device)
    acc_d = ops.proto_ema_accumulate_dense(out_feat, nearest, label.view(-1), sim, gumbel=gum)
    ln = [torch.from_numpy(g[k]).to(cuda_device) for k in ("ln_d_w", "ln_d_b", "ln_c_w", "ln_c_b")]
    acc_s = ops.proto_ema_accumulate(torch.from_numpy(g["embedding"]).to(cuda_device), label, protos, *ln,
                                     gumbel=gum, want_target=True)
    K = C * M
    assert torch.equal(acc_d.packed[K * D:], acc_s.packed[K * D:])            # counts
    assert torch.equal(acc_d.proto_target, acc_s.proto_target)
    assert (acc_d.packed[:K * D] - acc_s.packed[:K * D]).abs().max() <= 1e-5
    bank = PrototypeBank(C, M, D)
    assert callable(bank.prototype_learning)


def test_sync_average_equals_sum_on_one_rank(cuda_device):
    from coarse3d_b200 import distributed, ops
    g = EMA["tiny_det"]
    ln = [torch.from_numpy(g[k]).to(cuda_device) for k in ("ln_d_w", "ln_d_b", "ln_c_w", "ln_c_b")]
    emb = torch.from_numpy(g["embedding"]).to(cuda_device)
    label = torch.from_numpy(g["label"]).to(cuda_device)
    p0 = torch.from_numpy(g["prototypes0"]).to(cuda_device)
    a, _ = distributed.prototype_update(emb, label, p0, *ln, 0.9, assign_mode=ops.ASSIGN_ARGMAX, sync="sum")
    b, _ = distributed.prototype_update(emb, label, p0, *ln, 0.9, assign_mode=ops.ASSIGN_ARGMAX, sync="average")
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        distributed.prototype_update(emb, label, p0, *ln, 0.9, assign_mode=ops.ASSIGN_ARGMAX, sync="nope")
