"""The prototype memory bank of the reference's *Proto models, running on the B200.

The reference keeps the bank inside its three model classes
(`self.prototypes`, `self.feat_norm`, `self.mask_norm`,
pc_processor/models/salsanext_proto.py:322-328) and updates it in
`forward` (:494-530) -> `prototype_learning` (:337-402).  `PrototypeBank` holds
the same three members under the same names and exposes that block as
`update(embedding, label)`; see INTEGRATION.md for the four-line patch that
makes SalsaNextProto / RangeNetProto / SqueezeSegV3Proto call it.

State compatibility: `prototypes` stays a (C, M, D) float32
`nn.Parameter(requires_grad=False)`, so checkpoints load unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from coarse3d_b200 import distributed, ops


def momentum_update(old_value, new_value, momentum, debug=False):
    """salsanext_proto.py:19-31 (host-side helper kept for API compatibility)."""
    update = momentum * old_value + (1 - momentum) * new_value
    if debug:
        print("# old prot: {:.3f} x |{:.3f}|, new val: {:.3f} x |{:.3f}|, result= |{:.3f}|".format(
            momentum, torch.norm(old_value, p=2), (1 - momentum), torch.norm(new_value, p=2),
            torch.norm(update, p=2)))
    return update


def l2_normalize(x):
    return F.normalize(x, p=2, dim=-1)  # salsanext_proto.py:34-35


class PrototypeBank(nn.Module):
    def __init__(self, nclasses=20, sub_proto_size=20, proj_dim=256, ignore_label=0,
                 proto_mom=0.999, deterministic=False, max_rows=None):
        super().__init__()
        self.nclasses = nclasses
        self.sub_proto_size = sub_proto_size
        self.ignore_label = ignore_label
        self.proto_mom = proto_mom
        self.deterministic = deterministic
        self.max_rows = max_rows
        self.prototypes = nn.Parameter(torch.randn(nclasses, sub_proto_size, proj_dim),
                                       requires_grad=False)
        nn.init.trunc_normal_(self.prototypes, std=0.02)  # salsanext_proto.py:322-325
        self.feat_norm = nn.LayerNorm(proj_dim)  # :327
        self.mask_norm = nn.LayerNorm(nclasses)  # :328
        self.last = None

    @torch.no_grad()
    def update(self, embedding, label, gumbel=None, seed=None, want_target=False, group=None):
        """salsanext_proto.py:497-527 with proto_loss=True: one EMA step of the bank
        from this step's embedding (B,D,H,W) and labels (B,H,W); prototype sums are
        all-reduced over `group` before the EMA.  Returns proto_target (n,) or None."""
        mode = ops.ASSIGN_ARGMAX if (self.deterministic and gumbel is None) else None
        new, acc = distributed.prototype_update(
            embedding.contiguous(), label.contiguous().long(), self.prototypes.data,
            self.feat_norm.weight.data, self.feat_norm.bias.data, self.mask_norm.weight.data,
            self.mask_norm.bias.data, self.proto_mom, ignore_label=self.ignore_label,
            gumbel=gumbel, assign_mode=mode, seed=seed, max_rows=self.max_rows,
            want_target=want_target, group=group)
        self.prototypes = nn.Parameter(new, requires_grad=False)  # :394
        self.last = acc
        return acc.proto_target
