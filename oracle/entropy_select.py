"""Oracle: entropy-based pseudo-label selection (TEST INFRASTRUCTURE).

Restates `Trainer.entropy_based_selection`, tasks/weak_segmentation/trainer.py:447-518, in
torch-CPU.  `torch.multinomial(w, k, replacement=False)` is `topk(w / q, k)` with
q ~ Exp(1) drawn per element (aten/native/Distributions.cpp; checked against torch 2.11 in
tests/golden/make_golden.py), so the oracle takes the draws as input: `noise[b, cls]` is the
(H*W,) vector the reference would draw in iteration (b, cls).  Rule fixed where the
reference is undefined: among keys equal to the k-th largest all are selected (torch.topk
picks arbitrarily); with continuous noise ties do not occur.

Transcendentals.  torch's CPU `log` / `exp` are vectorised routines that are not correctly
rounded and differ between machines and builds, so a key that lies within a few ulp of the
(scan, class) threshold can fall on either side of it depending on where the reference runs.
`rule="rounded"` (the default, and the rule the device implements bit for bit) therefore fixes
every float32 operation of :459-466 as the correctly rounded one, in the reference's order:
x = p + 1e-10, l = log(x), t = p * l, sum over classes c = 0..C-1 sequentially, w = exp(sum),
key = w / q.  `rule="torch"` evaluates the reference's literal torch statements; the golden
vectors pin it exactly, and pin `rounded` everywhere except at pixels whose key is within
1e-5 (relative) of the threshold, which is where the reference itself is machine dependent.
"""
import numpy as np
import torch


def entropy_weights_rounded(output):
    """exp(-entropy) of (B,C,H,W) probabilities with every float32 operation correctly rounded
    (float64 log / exp rounded once to float32; sequential float32 sum over the classes)."""
    p = output.numpy().astype(np.float32)
    x = (p + np.float32(1e-10)).astype(np.float32)
    l = np.log(x.astype(np.float64)).astype(np.float32)
    t = (p * l).astype(np.float32)
    acc = np.zeros((p.shape[0],) + p.shape[2:], dtype=np.float32)
    for c in range(p.shape[1]):
        acc = (acc + t[:, c]).astype(np.float32)
    return torch.from_numpy(np.exp(acc.astype(np.float64)).astype(np.float32))


def entropy_based_selection(output, wss_mask, eval_mask, train_label, select_ratio, ignore_cls,
                            noise, rule="rounded"):
    """output (B,C,H,W) probs; wss_mask / eval_mask (B,H,W) bool; train_label (B,H,W) int64;
    noise (B,C,H*W) Exp(1) draws.  Returns (pseudo_label int64, new_wss_mask bool, keys, thr):
    keys (B,H*W) and thr {(b,cls): k-th largest key} let a checker identify near-threshold pixels."""
    bs, C, h, w = output.shape
    entropy = -torch.sum(output * torch.log(output + 1e-10), dim=1)          # :459-461
    _, pseudo_label = torch.max(output, dim=1)                                # :463
    entropy_weights = torch.exp(-1 * entropy)                                 # :466
    if rule == "rounded":
        entropy_weights = entropy_weights_rounded(output)
    else:
        assert rule == "torch"
    pseudo_label = pseudo_label.clone()
    pseudo_label[eval_mask == False] = ignore_cls                             # noqa: E712  :469
    low_entropy_mask = torch.zeros(bs, C, h, w).bool()
    keys = torch.zeros(bs, h * w)
    thr = {}
    for b in range(bs):
        for cls in torch.unique(train_label[b]):                              # :474-476
            if cls == ignore_cls:
                continue
            cls_mask = (pseudo_label[b] == cls) * (eval_mask[b] > 0)
            if cls_mask.sum() == 0:
                continue
            select_num = int(cls_mask.sum() * select_ratio)                   # :485
            if select_num < 1:
                continue
            weight_c = entropy_weights[b].clone()
            weight_c[cls_mask == False] = 0                                   # noqa: E712
            k = weight_c.reshape(-1) / noise[b, int(cls)]                     # multinomial, no replacement
            kth = torch.topk(k, select_num)[0][-1]
            sel = (k >= kth).reshape(h, w)
            keys[b][cls_mask.reshape(-1)] = k[cls_mask.reshape(-1)]
            thr[(b, int(cls))] = float(kth)
            low_entropy_mask[b, int(cls)] = (sel * cls_mask).bool()           # :498-506
    low = low_entropy_mask.sum(1).bool()                                      # :509
    pseudo_label = (pseudo_label * low).long()                                # :512
    pseudo_label[wss_mask] = train_label[wss_mask]                            # :515
    return pseudo_label, pseudo_label != ignore_cls, keys, thr                # :516
