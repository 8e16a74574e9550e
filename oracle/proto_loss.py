"""Oracle: class-prototype contrastive loss, forward + backward (TEST INFRASTRUCTURE).

Restates `ContrastMEMLoss` from the reference,
pc_processor/loss/contrast_pixel_loss.py:8-195, in torch-CPU float32.  The
gradient is torch autograd of the restated forward (what the reference does).

Randomness contract (the reference is non-deterministic):

* `torch.multinomial` anchor sampling (contrast_pixel_loss.py:114-116): the
  oracle takes the sampled indices `keep` (T, A) int64 as an input so both
  sides consume identical samples; with `keep=None` it samples like the
  reference does (used only for the timed CPU baseline and distribution tests).
* `torch.randperm` of each class's sub-prototypes (:142-143) only changes the
  float summation order; the oracle takes an optional permutation and defaults
  to the identity.
* The hard-coded `.cuda()` calls (:96-97,134-135,163) are dropped (device
  agnostic); nothing else differs.
"""
import torch
import torch.nn.functional as F


def entropy_weights(output):
    """contrast_pixel_loss.py:46-49.  output (B,C,H,W) probs -> (B,H,W)."""
    entropy = -torch.sum(output * torch.log(output + 1e-10), dim=1)
    entropy = entropy * entropy
    return torch.exp(-1 * entropy)


def masked_labels(labels, keep_mask, ignore_label):
    """contrast_pixel_loss.py:36-38."""
    labels = labels.clone()
    if keep_mask is not None:
        labels[keep_mask.bool() == False] = ignore_label  # noqa: E712
    return labels


def segments(labels_flat, ignore_label):
    """The (scan, class) pairs in the reference's X_ptr order
    (contrast_pixel_loss.py:82-89,100-109): scans ascending, classes ascending
    (torch.unique sorts), ignore label dropped."""
    segs = []
    for b in range(labels_flat.shape[0]):
        for c in torch.unique(labels_flat[b]).tolist():
            if c != ignore_label:
                segs.append((b, int(c)))
    return segs


def sample_anchors(labels_flat, weights_flat, segs, num_anchor, generator=None):
    """contrast_pixel_loss.py:111-116 -> keep (T, A) int64 pixel indices."""
    keep = torch.empty((len(segs), num_anchor), dtype=torch.int64)
    for t, (b, c) in enumerate(segs):
        w = weights_flat[b].clone()
        w[labels_flat[b] != c] = 0
        keep[t] = torch.multinomial(w.reshape(-1), num_anchor, replacement=True,
                                    generator=generator)
    return keep


def expand_queue(queue, perms=None):
    """contrast_pixel_loss.py:131-149: classes 1..C-1, M rows each."""
    C, M, D = queue.shape
    xs, ys = [], []
    for c in range(1, C):
        q = queue[c] if perms is None else queue[c, perms[c - 1]]
        xs.append(q)
        ys.append(torch.full((M,), float(c)))
    return torch.cat(xs, 0).float(), torch.cat(ys, 0)


def contrastive(X_anchor, y_anchor, queue, temperature, base_temperature, perms=None):
    """contrast_pixel_loss.py:151-195.  X_anchor (T,A,D), y_anchor (T,)."""
    num_anchor = X_anchor.shape[1]
    y_anchor = y_anchor.contiguous().view(-1, 1)
    anchor_feature = torch.cat(torch.unbind(X_anchor, dim=1), dim=0)  # row a*T+t
    X_contrast, y_contrast = expand_queue(queue, perms)
    y_contrast = y_contrast.view(-1, 1)
    mask = torch.eq(y_anchor, y_contrast.T).float()
    anchor_feature = F.normalize(anchor_feature, p=2, dim=-1)
    contrast_feature = F.normalize(X_contrast, p=2, dim=-1)
    adc = torch.einsum("nd,kd->nk", anchor_feature, contrast_feature)
    adc = torch.div(adc, temperature)
    logits_max, _ = torch.max(adc, dim=1, keepdim=True)
    logits = adc - logits_max.detach()
    mask = mask.repeat(num_anchor, 1)
    neg_mask = 1 - mask
    neg_logits = (torch.exp(logits) * neg_mask).sum(1, keepdim=True)
    exp_logits = torch.exp(logits)
    log_prob = logits - torch.log(exp_logits + neg_logits + 1e-6)
    mean_log_prob_pos = (mask * log_prob).sum(1) / mask.sum(1)
    loss = -(temperature / base_temperature) * mean_log_prob_pos
    return loss.mean()


def contrast_mem_loss(feats, output, labels, keep_mask, proto_queue, keep=None,
                      ignore_label=0, temperature=0.1, base_temperature=0.07,
                      num_anchor=50, perms=None, generator=None):
    """ContrastMEMLoss.forward (contrast_pixel_loss.py:27-75).

    feats (B,D,H,W) f32, output (B,C,H,W) probs, labels (B,H,W) int64,
    keep_mask (B,H,W) bool, proto_queue (1,C,M,D).  Returns (loss, keep, segs).
    """
    labels = masked_labels(labels, keep_mask, ignore_label)
    assert proto_queue is not None
    queue = proto_queue.squeeze(0)
    assert labels.shape[-1] == feats.shape[-1]
    B, D, H, W = feats.shape
    feats_ = feats.permute(0, 2, 3, 1).contiguous().view(B, -1, D)
    labels_flat = labels.contiguous().view(B, -1)
    segs = segments(labels_flat, ignore_label)
    if len(segs) == 0:
        raise ValueError("no anchor feature is selected for loss")
    if keep is None:
        w = entropy_weights(output).contiguous().view(B, -1)
        keep = sample_anchors(labels_flat, w, segs, num_anchor, generator)
    assert keep.shape == (len(segs), num_anchor)
    X_ = torch.stack([feats_[b, keep[t]] for t, (b, _) in enumerate(segs)], 0)
    y_ = torch.tensor([float(c) for _, c in segs])
    loss = contrastive(X_, y_, queue, temperature, base_temperature, perms)
    return loss, keep, segs
