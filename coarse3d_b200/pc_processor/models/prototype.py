"""The prototype memory bank of the reference's *Proto models, running on the B200.

The reference keeps the bank inside its three model classes
(`self.prototypes`, `self.feat_norm`, `self.mask_norm`,
pc_processor/models/salsanext_proto.py:322-328) and updates it in
`forward` (:494-530) -> `prototype_learning` (:337-402).  `PrototypeBank` holds
the same three members under the same names and exposes that block as
`update(embedding, label)`; see INTEGRATION.md for the four-line patch that
makes SalsaNextProto / RangeNetProto / SqueezeSegV3Proto call it, and `prototype_learning`
for the method those classes already call (`install()` binds it onto them).

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


@torch.no_grad()
def prototype_learning(self, out_feat, nearest_proto_distance, label, eval_mask, feat_proto_sim,
                       gumbel=None, seed=None):
    """Drop-in for `SalsaNextProto.prototype_learning` (salsanext_proto.py:337-402; the same
    method of RangeNetProto, rangenet_proto.py:460-567, and SqueezeSegV3Proto,
    squeezesegv3_Proto.py:253-351): same arguments, same return `(proto_logits, proto_target)`,
    same side effect (`self.prototypes` rebound to the updated, L2-normalised bank).
    `install()` binds it onto those classes; `self` needs `prototypes`, `nclasses`,
    `sub_proto_size`, `ignore_label`, `proto_mom`.  Only the labelled rows of the dense inputs
    are read.  With a process group, ranks combine as `self.proto_sync` says ("sum", default:
    sums and counts all-reduced before one identical EMA; "average": the reference's
    post-EMA average, :397-400)."""
    acc = ops.proto_ema_accumulate_dense(
        out_feat.contiguous().float(), nearest_proto_distance.contiguous().float(),
        label.contiguous().long(), feat_proto_sim.contiguous().float(),
        ignore_label=self.ignore_label, gumbel=gumbel, seed=seed,
        assign_mode=ops.ASSIGN_ARGMAX if (getattr(self, "deterministic", False) and gumbel is None) else None,
        max_rows=getattr(self, "max_rows", None), want_target=True)
    new = distributed.finish_update(self.prototypes.data, acc, self.proto_mom, self.ignore_label,
                                    group=getattr(self, "proto_group", None),
                                    sync=getattr(self, "proto_sync", "sum"))
    self.prototypes = nn.Parameter(new, requires_grad=False)                     # :394
    proto_logits = feat_proto_sim.reshape(feat_proto_sim.shape[0], -1)           # :343-345
    return proto_logits, acc.proto_target.view(label.shape)                      # :402


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
        self._flag_host, self._flag_event, self._flag_pending = None, None, False

    prototype_learning = prototype_learning   # the reference's method, same signature (:337-339)

    @torch.no_grad()
    def update(self, embedding, label, gumbel=None, seed=None, want_target=False, group=None):
        """salsanext_proto.py:497-527 with proto_loss=True: one EMA step of the bank
        from this step's embedding (B,D,H,W) and labels (B,H,W); prototype sums are
        all-reduced over `group` before the EMA.  Returns proto_target (n,) or None."""
        self.check_flags(wait=False)
        mode = ops.ASSIGN_ARGMAX if (self.deterministic and gumbel is None) else None
        new, acc = distributed.prototype_update(
            embedding.contiguous(), label.contiguous().long(), self.prototypes.data,
            self.feat_norm.weight.data, self.feat_norm.bias.data, self.mask_norm.weight.data,
            self.mask_norm.bias.data, self.proto_mom, ignore_label=self.ignore_label,
            gumbel=gumbel, assign_mode=mode, seed=seed, max_rows=self.max_rows,
            want_target=want_target, group=group, out=self.prototypes.data)
        # :394 rebinds `self.prototypes` to a new Parameter; the bank is updated in place here
        # (same values; optimisers / DDP / CUDA graphs keep a valid reference)
        self.last = acc
        if not torch.cuda.is_current_stream_capturing():
            # status flags of this update travel to the host asynchronously and are examined by
            # the NEXT call (or by check_flags()): no stall, and an overflow cannot go unnoticed
            if self._flag_host is None:
                self._flag_host = torch.zeros(4, dtype=torch.int32).pin_memory()
                self._flag_event = torch.cuda.Event()
            self._flag_host.copy_(acc.workspace[:16].view(torch.int32), non_blocking=True)
            self._flag_event.record()
            self._flag_pending = True
        return acc.proto_target

    def check_flags(self, wait=True):
        """Raise if the last examined update overflowed `max_rows` (it was then skipped
        entirely) or saw labels outside [0, C).  wait=False only looks at a finished copy."""
        if not self._flag_pending or self._flag_event is None or torch.cuda.is_current_stream_capturing():
            return
        if wait:
            self._flag_event.synchronize()
        elif not self._flag_event.query():
            return
        self._flag_pending = False
        flags = int(self._flag_host[2])
        if flags & ops.EMA_FLAG_OVERFLOW:
            raise RuntimeError("PrototypeBank.update: %d labelled pixels exceed max_rows; the update was "
                               "skipped -- construct the bank with a larger max_rows" % int(self._flag_host[1]))
        if flags & ops.EMA_FLAG_BAD_LABEL:
            raise ValueError("PrototypeBank.update: label outside [0, nclasses)")
