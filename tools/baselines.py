"""Baselines bench.py reports next to the B200 numbers (none of this is product code, and
nothing here imports `oracle/` or the `coarse3d_b200` package):

* `load_synth()`       the synthetic scan generator, loaded by file path so that the reference
                       arm never maps the CUDA library;
* `RealReference`      the UNMODIFIED reference modules, when a reference tree is present
                       (`/root/reference` in the build container; never on the GPU box):
                       RangeProjection.doProjection, ContrastMEMLoss, SalsaNextProto's prototype
                       block (pre-step :497-510 + prototype_learning :337-402), KNN;
* `torch_eager_step`   the reference's algorithm written with stock torch ops on whatever device
                       the tensors live on -- the "existing Blackwell path" of SURVEY.md 2b/8d
                       (torch 2.11 eager on the same B200).  It follows the reference statement by
                       statement (file:line cited), including the dense NCHW->NHWC copy, the
                       per-(scan, class) Python loop, the dense prototype similarity and the two
                       unfolds of KNN; the projection stays numpy on the host, as in the loaders.
"""
import importlib.util
import math
import os
import sys
import time
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_synth():
    spec = importlib.util.spec_from_file_location("_c3d_synth", os.path.join(ROOT, "coarse3d_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_c3d_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def find_reference_root():
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(root, "pc_processor", "loss", "contrast_pixel_loss.py")):
            return root
    return None


# ------------------------------------------------------------------ real reference --
def _install_stubs():
    """Absent third-party imports of the reference package + `.cuda()` as
    identity when no GPU is used (contrast_pixel_loss.py:96-97,134-135,163 hard-code it)."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

    try:
        import timm.models.layers  # noqa: F401
    except Exception:  # noqa: BLE001
        mod("timm"), mod("timm.models")
        mod("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_)
    try:
        import nuscenes  # noqa: F401
    except Exception:  # noqa: BLE001
        mod("nuscenes", NuScenes=_Any).__path__ = []
        mod("nuscenes.lidarseg").__path__ = []
        mod("nuscenes.lidarseg.lidarseg_utils", colormap_to_colors=None)
        mod("nuscenes.nuscenes", NuScenes=_Any)
        mod("nuscenes.utils", splits=None).__path__ = []
        mod("nuscenes.utils.splits")
        mod("nuscenes.utils.data_classes", LidarPointCloud=_Any)
        mod("nuscenes.utils.geometry_utils", view_points=None)
    for name, attrs in (("pyquaternion", dict(Quaternion=_Any)), ("tensorboardX", dict(SummaryWriter=_Any))):
        try:
            __import__(name)
        except Exception:  # noqa: BLE001
            mod(name, **attrs)


def import_reference(root, cpu=True):
    """`import pc_processor` from the reference tree `root` (stubs for absent dependencies)."""
    _install_stubs()
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if root not in sys.path:
        sys.path.insert(0, root)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):     # pc_processor/__init__.py:8-11 prints
        import pc_processor
    return pc_processor


class RealReference:
    """One step of the hot path through the unmodified reference modules (CPU)."""

    def __init__(self, root):
        self.pcp = import_reference(root)
        from einops import rearrange
        import pc_processor.models.salsanext_proto as sp
        self.sp, self.rearrange = sp, rearrange

    def step(self, synth, shp, n_scans, dim, seed0=1000, M=20):
        pcp, sp, rearrange = self.pcp, self.sp, self.rearrange
        H, W, C = shp.proj_h, shp.proj_w, shp.n_classes
        g = torch.Generator().manual_seed(seed0)
        scans = [synth.make_scan(shp, seed0 + i) for i in range(n_scans)]
        feats = torch.randn(n_scans, dim, H, W, generator=g)
        probs = torch.softmax(torch.randn(n_scans, C, H, W, generator=g), 1)
        argmax = torch.randint(0, C, (n_scans, H, W), generator=g)
        protos = F.normalize(torch.randn(C, M, dim, generator=g), dim=-1)
        ln_d, ln_c = torch.nn.LayerNorm(dim), torch.nn.LayerNorm(C)
        t0 = time.perf_counter()
        rp = pcp.dataset.preprocess.projection.RangeProjection(
            fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
        projs, labels = [], []
        for pts, _, weak in scans:
            _, rng, idx, _ = rp.doProjection(pts)                          # projection.py:43-115
            lab = np.zeros((H, W), np.int64)
            lab[idx >= 0] = weak[idx[idx >= 0]]                            # wss_sem_kitti_loader.py:124-132
            projs.append((rng, dict(rp.cached_data))), labels.append(lab)
        labels = torch.from_numpy(np.stack(labels))
        # model.forward's prototype block: salsanext_proto.py:497-527
        fake = types.SimpleNamespace(prototypes=torch.nn.Parameter(protos.clone(), requires_grad=False),
                                     nclasses=C, ignore_label=0, sub_proto_size=M, proto_mom=0.999)
        with torch.no_grad():
            out_feat = sp.l2_normalize(ln_d(rearrange(feats, "b c h w -> (b h w) c")))
            fake.prototypes.data.copy_(sp.l2_normalize(fake.prototypes))
            sim = torch.einsum("nd,kmd->nmk", out_feat, fake.prototypes)
            nearest = rearrange(ln_c(torch.amax(sim, dim=1)), "(b h w) k -> b k h w", b=n_scans, h=H)
            sp.SalsaNextProto.prototype_learning(fake, out_feat, nearest, labels.view(-1), None, sim)
            del sim, out_feat, nearest
        crit = pcp.loss.ContrastMEMLoss(ignore_label=0, temperature=0.07, num_anchor=512)   # trainer.py:366-371
        f = feats.clone().requires_grad_(True)
        loss = crit(feats=f, output=probs, labels=labels, keep_mask=labels > 0,
                    proto_queue=fake.prototypes.detach().unsqueeze(0))      # trainer.py:675-686
        loss.backward()
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):                     # KNN.__init__ prints a banner
            knn = pcp.postproc.KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), C)
        for (rng, cd), am in zip(projs, argmax):
            knn(torch.from_numpy(rng), torch.from_numpy(cd["uproj_depth"]), am,
                torch.from_numpy(cd["uproj_x_idx"]).long(), torch.from_numpy(cd["uproj_y_idx"]).long())
        return time.perf_counter() - t0


# ------------------------------------------------------------------- torch eager --
def _np_project(points, fov_up, fov_down, H, W):
    """projection.py:43-115 (numpy, host) -- the loaders run it on the CPU."""
    fov_up, fov_down = fov_up / 180.0 * np.pi, fov_down / 180.0 * np.pi
    fov_vert = abs(fov_up) + abs(fov_down)
    depth = np.linalg.norm(points[:, :3], 2, axis=1)
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    yaw, pitch = -np.arctan2(y, x), np.arcsin(z / depth)
    px = (yaw + np.pi) / (2 * np.pi) * W
    py = (1.0 - (pitch + abs(fov_down)) / fov_vert) * H
    px = np.maximum(np.minimum(W - 1, np.floor(px)), 0).astype(np.int32)
    py = np.maximum(np.minimum(H - 1, np.floor(py)), 0).astype(np.int32)
    order = np.argsort(depth)[::-1]
    rng = np.full((H, W), -1, dtype=np.float32)
    rng[py[order], px[order]] = depth[order]
    idx = np.full((H, W), -1, dtype=np.int32)
    idx[py[order], px[order]] = np.arange(depth.shape[0])[order]
    return rng, idx, px, py, depth.astype(np.float32)


def _eager_loss(feats, output, labels, keep_mask, queue, temperature=0.07, base_temperature=0.07,
                num_anchor=512, ignore_label=0):
    """contrast_pixel_loss.py:27-195 with stock torch ops on feats.device."""
    labels = labels.clone()
    labels[keep_mask == False] = ignore_label                                  # noqa: E712  :36-38
    B, D = feats.shape[0], feats.shape[1]
    entropy = -torch.sum(output * torch.log(output + 1e-10), dim=1)             # :46-49
    weights = torch.exp(-1 * entropy * entropy).contiguous().view(B, -1)
    labels = labels.contiguous().view(B, -1)
    feats = feats.permute(0, 2, 3, 1).contiguous().view(B, -1, D)               # :58-61
    xs, ys = [], []
    for b in range(B):                                                          # :82-123
        for cls in torch.unique(labels[b]):
            if cls == ignore_label:
                continue
            w = weights[b].clone()
            w[labels[b] != cls] = 0
            keep = torch.multinomial(w.reshape(-1), num_anchor, replacement=True)
            xs.append(feats[b, keep])
            ys.append(cls)
    X = torch.stack(xs, 0)
    y = torch.stack(ys).float().view(-1, 1)
    C, M, _ = queue.shape
    Xc = queue[1:].reshape(-1, D)                                               # :131-149 (no randperm)
    yc = torch.arange(1, C, device=feats.device).float().repeat_interleave(M).view(-1, 1)
    anchor = torch.cat(torch.unbind(X, dim=1), dim=0)                           # :155
    mask = torch.eq(y, yc.T).float()
    adc = torch.div(torch.einsum("nd,kd->nk", F.normalize(anchor, p=2, dim=-1),
                                 F.normalize(Xc, p=2, dim=-1)), temperature)    # :166-172
    logits = adc - torch.max(adc, dim=1, keepdim=True)[0].detach()
    mask = mask.repeat(num_anchor, 1)
    neg = (torch.exp(logits) * (1 - mask)).sum(1, keepdim=True)
    log_prob = logits - torch.log(torch.exp(logits) + neg + 1e-6)
    return (-(temperature / base_temperature) * (mask * log_prob).sum(1) / mask.sum(1)).mean()


def _eager_sinkhorn(out, iters=3, eps=0.05):
    """sinkhorn.py:5-33."""
    Q = torch.exp(out / eps).t()
    Bn, K = Q.shape[1], Q.shape[0]
    Q /= torch.sum(Q)
    for _ in range(iters):
        Q /= torch.sum(Q, dim=1, keepdim=True)
        Q /= K
        Q /= torch.sum(Q, dim=0, keepdim=True)
        Q /= Bn
    Q *= Bn
    Q = Q.t()
    return F.gumbel_softmax(Q, tau=0.5, hard=True), torch.argmax(Q, dim=1)


@torch.no_grad()
def _eager_ema(feats, label, protos, ln_d, ln_c, momentum=0.999, ignore_label=0):
    """salsanext_proto.py:497-510 + :337-394 (dense, as the reference evaluates it)."""
    B, D, H, W = feats.shape
    C, M, _ = protos.shape
    out_feat = F.normalize(ln_d(feats.permute(0, 2, 3, 1).reshape(-1, D)), p=2, dim=-1)
    protos = F.normalize(protos, p=2, dim=-1)
    sim = torch.einsum("nd,kmd->nmk", out_feat, protos)
    nearest = ln_c(torch.amax(sim, dim=1))
    label = label.view(-1)
    mask = label == torch.max(nearest, 1)[1]
    new = protos.clone()
    for c in range(C):
        if c == ignore_label:
            continue
        sel = label == c
        init_q = sim[..., c][sel, ...]
        if init_q.shape[0] == 0:
            continue
        q, _ = _eager_sinkhorn(init_q)
        m_c = mask[sel]
        m_q = q * m_c[:, None]
        f = m_q.transpose(0, 1) @ (out_feat[sel] * m_c[:, None])
        n = torch.sum(m_q, dim=0)
        if torch.sum(n) > 0:
            f = F.normalize(f, p=2, dim=-1)
            new[c, n != 0, :] = momentum * new[c, n != 0, :] + (1 - momentum) * f[n != 0, :]
    return F.normalize(new, p=2, dim=-1)


def _eager_knn(proj_range, unproj_range, proj_argmax, px, py, knn=5, search=5, sigma=1.0, cutoff=1.0,
               nclasses=20):
    """knn.py:54-142 (one scan)."""
    dev = proj_range.device
    H, W = proj_range.shape
    P = unproj_range.shape[0]
    pad = int((search - 1) / 2)
    unfold = F.unfold(proj_range[None, None], kernel_size=(search, search), padding=(pad, pad))
    idx_list = py * W + px
    unproj = unfold[:, :, idx_list]
    unproj[unproj < 0] = float("inf")
    center = int(((search * search) - 1) / 2)
    unproj[:, center, :] = unproj_range
    k2 = torch.abs(unproj - unproj_range)
    coords = torch.arange(search)
    xg = coords.repeat(search).view(search, search)
    xy = torch.stack([xg, xg.t()], dim=-1).float()
    mean, var = (search - 1) / 2., sigma ** 2.
    gk = (1. / (2. * math.pi * var)) * torch.exp(-torch.sum((xy - mean) ** 2., dim=-1) / (2 * var))
    inv = (1 - (gk / torch.sum(gk)).view(search, search))[None, :, None].view(1, -1, 1).to(dev)
    k2 = k2 * inv
    _, knn_idx = k2.topk(knn, dim=1, largest=False, sorted=False)
    am_unfold = F.unfold(proj_argmax[None, None].float(), kernel_size=(search, search),
                         padding=(pad, pad)).long()
    am = torch.gather(input=am_unfold[:, :, idx_list], dim=1, index=knn_idx)
    if cutoff > 0:
        d = torch.gather(input=k2, dim=1, index=knn_idx)
        am[d > cutoff] = nclasses
    onehot = torch.zeros((1, nclasses + 1, P), device=dev)
    onehot = onehot.scatter_add_(1, am, torch.ones_like(am).float())
    return onehot[:, 1:-1].argmax(dim=1) + 1


def torch_eager_step(synth, shp, scans, dev, dim, state):
    """One step over `scans` (list of (points, full, weak)) with the activations of `state`
    resident on `dev`; the projection runs in numpy on the host and its outputs are copied
    to the device, as in the reference loaders (wss_sem_kitti_loader.py:117-122 -> trainer.py:599)."""
    H, W, C = shp.proj_h, shp.proj_w, shp.n_classes
    projs, labels = [], []
    for pts, _, weak in scans:
        rng, idx, px, py, depth = _np_project(pts, shp.fov_up, shp.fov_down, H, W)
        lab = np.zeros((H, W), np.int64)
        lab[idx >= 0] = weak[idx[idx >= 0]]
        labels.append(lab)
        projs.append(tuple(torch.from_numpy(a).to(dev, non_blocking=True) for a in (rng, depth, px, py)))
    labels = torch.from_numpy(np.stack(labels)).to(dev)
    state["protos"] = _eager_ema(state["feats"], labels, state["protos"], state["ln_d"], state["ln_c"])
    f = state["feats"].detach().requires_grad_(True)
    loss = _eager_loss(f, state["probs"], labels, labels > 0, state["protos"])
    loss.backward()
    out = []
    for (rng, depth, px, py), am in zip(projs, state["argmax"]):
        out.append(_eager_knn(rng, depth, am, px.long(), py.long(), nclasses=C))
    return loss.detach(), f.grad, out
