"""Oracle: EMA prototype update (TEST INFRASTRUCTURE).

Restates, in torch-CPU float32,
  * the similarity pre-step inside `SalsaNextProto.forward`,
    pc_processor/models/salsanext_proto.py:497-510,
  * `prototype_learning`, salsanext_proto.py:337-402 (identical logic in
    rangenet_proto.py:460-567 and squeezesegv3_Proto.py:253-351),
  * `momentum_update` / `l2_normalize`, salsanext_proto.py:19-35,
  * `distributed_sinkhorn`, pc_processor/models/sinkhorn.py:5-33.

Randomness contract: `F.gumbel_softmax(Q, tau=0.5, hard=True)` (sinkhorn.py:31)
draws Gumbel noise per present class in class order.  The oracle takes that
noise as input (`gumbel[c]`, shape (n_c, M)); `gumbel=None` selects the
deterministic assignment `one_hot(argmax Q)` the reference keeps commented out
at sinkhorn.py:30.

Multi-GPU rule: ranks all-reduce the per-(class, sub-prototype)
feature sums and counts and then apply one identical EMA (`ema_from_sums`);
on one rank this coincides with the reference.  The reference instead averages
the per-rank prototypes after local EMAs (salsanext_proto.py:397-400).
"""
import torch
import torch.nn.functional as F


def l2_normalize(x):
    return F.normalize(x, p=2, dim=-1)  # salsanext_proto.py:34-35


def momentum_update(old_value, new_value, momentum):
    return momentum * old_value + (1 - momentum) * new_value  # :19-31


def pre_step(embedding, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b, eps=1e-5):
    """salsanext_proto.py:497-510.  embedding (B,D,H,W); prototypes (C,M,D).
    Returns out_feat (n,D), protos_n (C,M,D), sim (n,M,C), nearest (n,C)."""
    B, D, H, W = embedding.shape
    C = prototypes.shape[0]
    out_feat = embedding.permute(0, 2, 3, 1).reshape(-1, D)
    out_feat = F.layer_norm(out_feat, (D,), ln_d_w, ln_d_b, eps)
    out_feat = l2_normalize(out_feat)
    protos_n = l2_normalize(prototypes)
    sim = torch.einsum("nd,kmd->nmk", out_feat, protos_n)
    nearest = torch.amax(sim, dim=1)
    nearest = F.layer_norm(nearest, (C,), ln_c_w, ln_c_b, eps)
    return out_feat, protos_n, sim, nearest


def sinkhorn(out, gumbel=None, iterations=3, epsilon=0.05):
    """sinkhorn.py:5-33.  out (n_c, M) -> (q (n_c, M), idx (n_c,))."""
    Q = torch.exp(out / epsilon).t()
    Bn, K = Q.shape[1], Q.shape[0]
    Q = Q / torch.sum(Q)
    for _ in range(iterations):
        Q = Q / torch.sum(Q, dim=1, keepdim=True)
        Q = Q / K
        Q = Q / torch.sum(Q, dim=0, keepdim=True)
        Q = Q / Bn
    Q = Q * Bn
    Q = Q.t()
    idx = torch.argmax(Q, dim=1)
    if gumbel is None:
        q = F.one_hot(idx, num_classes=Q.shape[1]).float()
    else:
        # F.gumbel_softmax(Q, tau=0.5, hard=True) with the noise injected
        y_soft = ((Q + gumbel) / 0.5).softmax(-1)
        index = y_soft.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(Q).scatter_(-1, index, 1.0)
        q = y_hard - y_soft + y_soft
    return q, idx


def segment_sums(out_feat, nearest, label, sim, nclasses, ignore_label, gumbel=None):
    """The per-class loop of prototype_learning (salsanext_proto.py:340-377,
    390-392) up to the feature sums.  Rows may be restricted to labelled
    pixels.  Returns sums (C,M,D), counts (C,M), proto_target (n,)."""
    n, M, C = sim.shape
    D = out_feat.shape[1]
    pred = torch.max(nearest, 1)[1]
    mask = label == pred
    sums = torch.zeros(C, M, D)
    counts = torch.zeros(C, M)
    proto_target = torch.zeros_like(label).float()
    for c in range(nclasses):
        if c == ignore_label:
            continue
        sel = label == c
        init_q = sim[sel][..., c]
        if init_q.shape[0] == 0:
            continue
        g = None if gumbel is None else gumbel[c]
        q, idx = sinkhorn(init_q, g)
        m_c = mask[sel].float()
        m_q = q * m_c[:, None]
        c_q = out_feat[sel] * m_c[:, None]
        sums[c] = m_q.transpose(0, 1) @ c_q
        counts[c] = torch.sum(m_q, dim=0)
        proto_target[sel] = idx.float() + M * c
    return sums, counts, proto_target


def ema_from_sums(protos_n, sums, counts, momentum, ignore_label=0):
    """salsanext_proto.py:379-394: normalise sums, EMA where count != 0,
    renormalise every prototype."""
    protos = protos_n.clone()
    C = protos.shape[0]
    for c in range(C):
        if c == ignore_label:
            continue
        n = counts[c]
        if torch.sum(n) > 0:
            f = F.normalize(sums[c], p=2, dim=-1)
            nz = n != 0
            protos[c, nz] = momentum_update(protos[c, nz], f[nz], momentum)
    return l2_normalize(protos)


def prototype_learning(embedding, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                       nclasses, ignore_label, momentum, gumbel=None,
                       labelled_only=False):
    """Pre-step + prototype_learning on one rank.  `labelled_only=True`
    evaluates only rows with label != ignore (the rows that influence the
    update); results are identical because every step is row-local except the
    per-class Sinkhorn, which only ever sees labelled rows."""
    label = label.reshape(-1)
    if labelled_only:
        B, D, H, W = embedding.shape
        rows = torch.nonzero(label != ignore_label).reshape(-1)
        feat_rows = embedding.permute(0, 2, 3, 1).reshape(-1, D)[rows]
        emb = feat_rows.t().reshape(1, D, 1, -1)
        out_feat, protos_n, sim, nearest = pre_step(
            emb, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b)
        sums, counts, tgt = segment_sums(out_feat, nearest, label[rows], sim,
                                         nclasses, ignore_label, gumbel)
        proto_target = torch.zeros_like(label).float()
        proto_target[rows] = tgt
    else:
        out_feat, protos_n, sim, nearest = pre_step(
            embedding, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b)
        sums, counts, proto_target = segment_sums(out_feat, nearest, label, sim,
                                                  nclasses, ignore_label, gumbel)
    new_protos = ema_from_sums(protos_n, sums, counts, momentum, ignore_label)
    return new_protos, sums, counts, proto_target
