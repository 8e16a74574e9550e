"""ContrastMEMLoss with the reference's interface, running on the B200.

Mirrors pc_processor/loss/contrast_pixel_loss.py:8-195: same constructor
arguments, same keyword `forward(feats=, output=, labels=, keep_mask=,
proto_queue=)` returning a 0-dim loss that participates in autograd (gradient
to `feats` only), same assertions.  The arithmetic runs in
`c3d_proto_loss_forward` / `c3d_proto_loss_backward`.

Differences a caller can observe:
  * anchors are sampled on the device (Philox) instead of by torch.multinomial
    on the global generator -- same distribution, different stream.  Pass
    `keep=` (T, num_anchor) int64 to inject the draws (parity tests do);
  * importing this module does not reseed the global RNG (the reference calls
    torch.random.manual_seed(0) at import, contrast_pixel_loss.py:5);
  * with no labelled pixel the reference crashes on a None tensor; here the
    loss is NaN, and with `is_debug=True` an AssertionError is raised.
"""
import torch
import torch.nn as nn

from coarse3d_b200 import ops


class ContrastMEMLoss(nn.Module):
    def __init__(self, ignore_label=0, temperature=0.1, base_temperature=0.07, num_anchor=50,
                 is_debug=False):
        super(ContrastMEMLoss, self).__init__()
        self.temperature = temperature
        self.base_temperature = base_temperature
        self.num_anchor = num_anchor
        self.ignore_label = ignore_label
        self.is_debug = is_debug
        self.sub_proto = True
        self.last_workspace = None

    def forward(self, feats=None, output=None, labels=None, keep_mask=None, proto_queue=None,
                keep=None, seed=None):
        assert proto_queue is not None
        proto_queue = proto_queue.squeeze(0)
        if self.is_debug:
            print("queue size, max views : ", proto_queue.shape)
        assert output is not None  # the reference asserts weights is not None (:111)
        assert labels.shape[-1] == feats.shape[-1], "{} {}".format(labels.shape, feats.shape)
        cfg = ops.ProtoLossConfig(self.ignore_label, self.temperature, self.base_temperature,
                                  self.num_anchor)
        if keep_mask is not None and keep_mask.dtype != torch.bool:
            keep_mask = keep_mask.bool()
        loss, ws = ops.proto_loss(
            feats.contiguous(), output.contiguous(), labels.contiguous().long(),
            None if keep_mask is None else keep_mask.contiguous(),
            proto_queue.contiguous().float(), cfg, keep=keep, seed=seed)
        self.last_workspace = ws
        if self.is_debug:
            T, n_lab, flags = ops.proto_loss_info(ws)
            assert not (flags & ops.FLAG_NO_ANCHOR), "no anchor feature is selected for loss"
            assert not (flags & (ops.FLAG_BAD_KEEP | ops.FLAG_KEEP_ROWS)), "bad injected anchors"
            assert not (flags & ops.FLAG_BAD_LABEL), "label outside [0, C)"
        return loss
