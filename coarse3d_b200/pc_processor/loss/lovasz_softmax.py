"""Lovasz_softmax with the reference's interface, running on the B200.

Mirrors pc_processor/loss/lovasz_softmax.py:160-179 (`Lovasz_softmax`) and :67-98
(`lovasz_softmax`): same constructor arguments, `forward(probas, labels)` returning a 0-dim
loss with autograd to `probas`, the same NaN assertion.  The arithmetic runs in
`c3d_lovasz_forward` / `c3d_lovasz_backward`.

Differences a caller can observe:
  * ties between equal errors rank by ascending pixel index (torch.sort is unstable);
  * any number of valid pixels up to 2^24 per call (weak labels ~1e3 per batch: one CTA per class
    sorts in shared memory; dense / pseudo labels: a device-wide radix sort).  The capacity is
    chosen per call from the number of valid pixels (one host read); a caller that fixes it
    (`max_valid=`, e.g. inside a CUDA graph) and exceeds it is never silently wrong: the loss
    comes back NaN, so the NaN assertion below (the reference's own, :178) fires; `strict=True`
    raises a ValueError with the pixel count instead;
  * with no valid pixel the reference returns an empty tensor (and the trainer skips such
    batches, trainer.py:586-589); here the loss is 0 with zero gradient;
  * labels >= C are ignored pixels here (the reference counts them as background of every class).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from coarse3d_b200 import ops


def lovasz_softmax(probas, labels, classes="present", per_image=False, ignore=None, softmax=False,
                   strict=False, max_valid=None):
    if softmax:
        probas = F.softmax(probas, 1)                              # lovasz_softmax.py:80-81
    if probas.dim() == 3:                                          # (B,C,N) form, :143-147
        probas = probas.unsqueeze(-1)
        labels = labels.unsqueeze(-1)
    labels = labels.long()
    if per_image:                                                  # :82-89: mean over images
        losses = [_one(p.unsqueeze(0), l.unsqueeze(0), classes, ignore, strict, max_valid)
                  for p, l in zip(probas, labels)]
        acc = losses[0]
        for v in losses[1:]:
            acc = acc + v
        return acc if len(losses) == 1 else acc / len(losses)
    return _one(probas, labels, classes, ignore, strict, max_valid)


def _one(probas, labels, classes, ignore, strict, max_valid=None):
    loss, ws = ops.lovasz_softmax(probas.float(), labels, ignore=ignore, classes=classes, max_valid=max_valid)
    if strict:
        n_valid, _, flags = ops.lovasz_info(ws)
        if flags & 1:
            raise ValueError("Lovasz_softmax: %d labelled pixels exceed max_valid=%s" % (n_valid, max_valid))
    return loss


class Lovasz_softmax(nn.Module):
    def __init__(self, classes="present", per_image=False, ignore=None, softmax=False, strict=False,
                 max_valid=None):
        super(Lovasz_softmax, self).__init__()
        self.max_valid = max_valid
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore
        self.softmax = softmax
        self.strict = strict

    def forward(self, probas, labels):
        loss = lovasz_softmax(probas, labels, self.classes, self.per_image, self.ignore, self.softmax,
                              self.strict, self.max_valid)
        assert not torch.any(torch.isnan(loss)), "lov loss is none"   # lovasz_softmax.py:178
        return loss
