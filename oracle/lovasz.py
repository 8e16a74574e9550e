"""Oracle: Lovasz-softmax loss (TEST INFRASTRUCTURE).

Restates pc_processor/loss/lovasz_softmax.py:51-157 (`lovasz_grad`, `lovasz_softmax_flat`,
`flatten_probas`, `mean`) in torch-CPU for per_image=False.  Rules fixed where the reference
is undefined: `torch.sort(descending=True)` is unstable, here equal errors keep ascending
pixel order (stable sort); with no valid pixel the reference returns an empty tensor, here 0.
Gradients come from torch autograd on these statements.
"""
import torch


def lovasz_grad(gt_sorted):                                        # :51-64
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1.0 - intersection / union
    if p > 1:
        jaccard[1:p] = jaccard[1:p] - jaccard[0:-1]
    return jaccard


def lovasz_softmax(probas, labels, ignore=None, classes="present"):
    """probas (B,C,H,W) float32 (may require grad), labels (B,H,W) int64 -> 0-dim loss."""
    B, C, H, W = probas.shape
    pred = probas.permute(0, 2, 3, 1).contiguous().view(-1, C)     # :148-150
    lab = labels.view(-1)
    if ignore is not None:                                         # :153-156
        valid = lab != ignore
        pred = pred[torch.nonzero(valid, as_tuple=False).squeeze(1)]
        lab = lab[valid]
    if pred.numel() == 0:
        return probas.sum() * 0.0
    losses = []
    class_to_sum = list(range(C)) if classes in ("all", "present") else list(classes)   # :117
    for c in class_to_sum:                                         # :119-133
        fg = (lab == c).float()
        if classes == "present" and fg.sum() == 0:
            continue
        errors = (fg - pred[:, c]).abs()
        errors_sorted, perm = torch.sort(errors, dim=0, descending=True, stable=True)
        fg_sorted = fg[perm]
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg_sorted)))
    acc = losses[0]                                                # mean(), :33-48
    for v in losses[1:]:
        acc = acc + v
    return acc if len(losses) == 1 else acc / len(losses)
