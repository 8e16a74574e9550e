#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the build container, where the reference tree is mounted read-only:

    python tests/golden/make_golden.py [/root/reference]

The reference ships no tests or golden vectors, so the
oracle is pinned against the reference's own modules run here on seeded
inputs.  The fixtures (inputs + reference outputs) are committed; the GPU box
has no /root/reference and only reads the .npz files.

Accommodations needed to run the reference on a CPU-only host:
  * empty stub modules for absent third-party imports (timm, nuscenes.*,
    pyquaternion, tensorboardX) so that `import pc_processor` succeeds;
  * `torch.Tensor.cuda` patched to identity (contrast_pixel_loss.py:96-97,
    134-135,163 hard-code `.cuda()`);
  * `torch.multinomial` wrapped to RECORD the draws (so the oracle / CUDA path
    can be fed the same anchors);
  * Gumbel noise of `F.gumbel_softmax` reproduced by re-seeding the CPU
    generator and replaying the draws in class order (checked below).
"""
import os
import sys
import types

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))  # repo root (synth)


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

    mod("timm"), mod("timm.models")
    mod("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_)
    nus = mod("nuscenes", NuScenes=_Any)
    nus.__path__ = []
    mod("nuscenes.lidarseg").__path__ = []
    mod("nuscenes.lidarseg.lidarseg_utils", colormap_to_colors=None)
    mod("nuscenes.nuscenes", NuScenes=_Any)
    mod("nuscenes.utils", splits=None).__path__ = []
    mod("nuscenes.utils.splits")
    mod("nuscenes.utils.data_classes", LidarPointCloud=_Any)
    mod("nuscenes.utils.geometry_utils", view_points=None)
    mod("pyquaternion", Quaternion=_Any)
    mod("tensorboardX", SummaryWriter=_Any)
    torch.Tensor.cuda = lambda self, *a, **k: self


def import_reference():
    install_stubs()
    sys.path.insert(0, REF)
    import pc_processor  # noqa: F401
    return pc_processor


def gold_projection(pcp):
    from coarse3d_b200 import synth
    RP = pcp.dataset.preprocess.projection.RangeProjection
    cases = {}
    specs = [
        # name, shape, n_points, H, W, seed, depth override?
        ("kitti_small", synth.KITTI, 12000, 32, 512, 11, False),
        ("kitti_w2048", synth.KITTI, 9000, 64, 2048, 12, False),
        ("nusc_small", synth.NUSCENES, 5000, 32, 1024, 13, False),
        ("poss_w1800", synth.POSS, 6000, 40, 1800, 14, False),
        ("kitti_depth_override", synth.KITTI, 8000, 32, 512, 15, True),
    ]
    for name, shp, n, H, W, seed, override in specs:
        pts, _, weak = synth.make_scan(shp, seed, n)
        rp = RP(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
        depth = None
        if override:
            # wss_sem_kitti_loader.py:134-140: unlabeled points pushed to 10000
            depth = np.linalg.norm(pts[:, :3], 2, axis=1)
            depth[weak == 0] = 10000.0
            # make the far depths distinct so the unstable argsort has no ties
            far = np.nonzero(weak == 0)[0]
            depth[far] = (10000.0 + np.arange(far.size, dtype=np.float32)).astype(np.float32)
        pc, rng_img, idx, mask = rp.doProjection(pts, depth)
        cases[name] = dict(
            points=pts, H=H, W=W, fov_up=shp.fov_up, fov_down=shp.fov_down,
            has_depth=override, depth=depth if override else np.zeros(0, np.float32),
            proj_pointcloud=pc, proj_range=rng_img, proj_idx=idx, proj_mask=mask,
            uproj_x_idx=rp.cached_data["uproj_x_idx"],
            uproj_y_idx=rp.cached_data["uproj_y_idx"],
            uproj_depth=rp.cached_data["uproj_depth"])
        # the unstable-sort tie caveat: record whether any pixel has a depth tie
        d = rp.cached_data["uproj_depth"]
        lin = rp.cached_data["uproj_y_idx"].astype(np.int64) * W + rp.cached_data["uproj_x_idx"]
        key = np.stack([lin, d.view(np.uint32).astype(np.int64)], 1)
        cases[name]["n_depth_ties"] = len(key) - len(np.unique(key, axis=0))
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "projection.npz"), **flat)
    print("projection:", {k: int(c["n_depth_ties"]) for k, c in cases.items()})


def gold_knn(pcp):
    from coarse3d_b200 import synth
    KNN = pcp.postproc.KNN
    RP = pcp.dataset.preprocess.projection.RangeProjection
    cases = {}
    specs = [
        ("k5s5", synth.KITTI, 8000, 32, 512, 21, dict(knn=5, search=5, sigma=1.0, cutoff=1.0), 20),
        ("k3s3", synth.POSS, 5000, 40, 360, 22, dict(knn=3, search=3, sigma=2.0, cutoff=0.5), 14),
        ("k7s7", synth.NUSCENES, 6000, 32, 256, 23, dict(knn=7, search=7, sigma=1.5, cutoff=2.0), 17),
    ]
    for name, shp, n, H, W, seed, params, ncls in specs:
        pts, _, _ = synth.make_scan(shp, seed, n)
        rp = RP(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
        _, rng_img, _, _ = rp.doProjection(pts)
        g = torch.Generator().manual_seed(seed)
        argmax = torch.randint(0, ncls, (H, W), generator=g)
        px = torch.from_numpy(rp.cached_data["uproj_x_idx"]).long()
        py = torch.from_numpy(rp.cached_data["uproj_y_idx"]).long()
        ur = torch.from_numpy(rp.cached_data["uproj_depth"])
        out = KNN(params, ncls)(torch.from_numpy(rng_img), ur, argmax, px, py)
        cases[name] = dict(proj_range=rng_img, unproj_range=ur.numpy(),
                           proj_argmax=argmax.numpy(), px=px.numpy(), py=py.numpy(),
                           nclasses=ncls, out=out.numpy(), **params)
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **flat)
    print("knn: ok")


def gold_loss(pcp):
    Loss = pcp.loss.ContrastMEMLoss
    cases = {}
    specs = [
        # name, B, D, H, W, C, M, A, labelled fraction, seed
        ("tiny", 2, 16, 8, 64, 6, 4, 16, 0.05, 31),
        ("weak", 2, 32, 16, 128, 20, 20, 64, 0.004, 32),
        ("dense", 1, 24, 8, 96, 8, 5, 32, 0.5, 33),
    ]
    for name, B, D, H, W, C, M, A, frac, seed in specs:
        g = torch.Generator().manual_seed(seed)
        feats = torch.randn(B, D, H, W, generator=g, requires_grad=True)
        output = torch.softmax(torch.randn(B, C, H, W, generator=g) * 2, 1)
        labels = torch.randint(0, C, (B, H, W), generator=g)
        keep_mask = torch.rand(B, H, W, generator=g) < frac
        queue = torch.nn.functional.normalize(torch.randn(1, C, M, D, generator=g), dim=-1) * \
            (0.5 + torch.rand(1, C, M, 1, generator=g))  # un-normalised bank rows
        draws = []
        real_multinomial = torch.multinomial

        def rec(*a, **k):
            r = real_multinomial(*a, **k)
            draws.append(r.clone())
            return r

        torch.multinomial = rec
        try:
            torch.manual_seed(seed)
            crit = Loss(ignore_label=0, temperature=0.07, num_anchor=A)
            loss = crit(feats=feats, output=output, labels=labels, keep_mask=keep_mask,
                        proto_queue=queue)
        finally:
            torch.multinomial = real_multinomial
        loss.backward()
        cases[name] = dict(feats=feats.detach().numpy(), output=output.numpy(),
                           labels=labels.numpy(), keep_mask=keep_mask.numpy(),
                           queue=queue.numpy(), keep=torch.stack(draws, 0).numpy(),
                           loss=loss.detach().numpy(), grad=feats.grad.numpy(),
                           temperature=0.07, base_temperature=0.07, num_anchor=A)
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "proto_loss.npz"), **flat)
    print("loss:", {k: float(c["loss"]) for k, c in cases.items()})


def gold_ema(pcp):
    from pc_processor.models.salsanext_proto import SalsaNextProto, l2_normalize
    from einops import rearrange
    cases = {}
    specs = [
        # name, B, D, H, W, C, M, labelled fraction, gumbel?, seed
        ("tiny_det", 2, 16, 8, 32, 6, 4, 0.3, False, 41),
        ("tiny_gumbel", 2, 16, 8, 32, 6, 4, 0.3, True, 42),
        ("weak_gumbel", 2, 32, 16, 64, 20, 20, 0.05, True, 43),
    ]
    for name, B, D, H, W, C, M, frac, use_gumbel, seed in specs:
        g = torch.Generator().manual_seed(seed)
        emb = torch.nn.functional.normalize(torch.randn(B, D, H, W, generator=g), dim=1)
        protos0 = torch.randn(C, M, D, generator=g) * 0.02
        # make features cluster around "their" class so that mask has hits
        label = torch.randint(1, C, (B, H, W), generator=g)
        centers = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=-1)
        emb = torch.nn.functional.normalize(
            emb + 1.5 * centers[label].permute(0, 3, 1, 2), dim=1)
        protos0 = protos0 + centers[:, None, :] * 0.05
        label = label * (torch.rand(B, H, W, generator=g) < frac)
        ln_d = torch.nn.LayerNorm(D)
        ln_c = torch.nn.LayerNorm(C)
        with torch.no_grad():
            ln_d.weight.copy_(1 + 0.1 * torch.randn(D, generator=g))
            ln_d.bias.copy_(0.1 * torch.randn(D, generator=g))
            ln_c.weight.copy_(1 + 0.1 * torch.randn(C, generator=g))
            ln_c.bias.copy_(0.1 * torch.randn(C, generator=g))
        fake = types.SimpleNamespace(
            prototypes=torch.nn.Parameter(protos0.clone(), requires_grad=False),
            nclasses=C, ignore_label=0, sub_proto_size=M, proto_mom=0.9)
        with torch.no_grad():
            # salsanext_proto.py:497-510, verbatim sequence of calls
            out_feat = rearrange(emb, "b c h w -> (b h w) c")
            out_feat = ln_d(out_feat)
            out_feat = l2_normalize(out_feat)
            fake.prototypes.data.copy_(l2_normalize(fake.prototypes))
            sim = torch.einsum("nd,kmd->nmk", out_feat, fake.prototypes)
            nearest = torch.amax(sim, dim=1)
            nearest = ln_c(nearest)
            nearest = rearrange(nearest, "(b h w) k -> b k h w", b=B, h=H)
            label_expand = label.view(-1)
            import pc_processor.models.salsanext_proto as sp
            real_sinkhorn = sp.distributed_sinkhorn
            if not use_gumbel:
                def det(out, *a, **k):
                    _, idx = real_sinkhorn(out, *a, **k)
                    return torch.nn.functional.one_hot(idx, out.shape[1]).float(), idx
                sp.distributed_sinkhorn = det
            torch.manual_seed(seed)
            try:
                logits, target = SalsaNextProto.prototype_learning(
                    fake, out_feat, nearest, label_expand, None, sim)
            finally:
                sp.distributed_sinkhorn = real_sinkhorn
            # replay the Gumbel draws: one exponential_ per present class, in order
            torch.manual_seed(seed)
            gumbel = torch.zeros(C, int((label_expand > 0).sum()), M)
            ncls = []
            for c in range(C):
                n_c = int((label_expand == c).sum()) if c != 0 else 0
                ncls.append(n_c)
                if n_c:
                    gumbel[c, :n_c] = -torch.empty(n_c, M).exponential_().log()
        cases[name] = dict(
            embedding=emb.numpy(), label=label.numpy(), prototypes0=protos0.numpy(),
            ln_d_w=ln_d.weight.detach().numpy(), ln_d_b=ln_d.bias.detach().numpy(),
            ln_c_w=ln_c.weight.detach().numpy(), ln_c_b=ln_c.bias.detach().numpy(),
            momentum=0.9, use_gumbel=use_gumbel, gumbel=gumbel.numpy(),
            n_per_class=np.asarray(ncls), prototypes1=fake.prototypes.detach().numpy(),
            proto_target=target.numpy())
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "proto_ema.npz"), **flat)
    print("ema: ok")


def gold_assemble(pcp):
    """Caller side of the projection: the statements of wss_sem_kitti_loader.py:124-132,
    159-172 and trainer.py:600-608 executed on the reference RangeProjection's output."""
    from coarse3d_b200 import synth
    RP = pcp.dataset.preprocess.projection.RangeProjection
    cases = {}
    mean = [12.12, 10.88, 0.23, -1.04, 0.21]   # config_semantic_kitti.yaml:142-153
    stds = [12.32, 11.47, 6.91, 0.86, 0.16]
    for name, shp, n, H, W, seed, normalise in [
            ("kitti_norm", synth.KITTI, 9000, 32, 512, 51, True),
            ("nusc_raw", synth.NUSCENES, 5000, 32, 256, 52, False)]:
        pts, sem, weak = synth.make_scan(shp, seed, n)
        pts[::97, 3] = -1.0   # exercise the intensity != -1 mask on real points too
        sem_label, weak_label = sem.astype(np.int32), weak.astype(np.int32)
        rp = RP(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
        proj_pointcloud, proj_range, proj_idx, proj_eval_mask = rp.doProjection(pts)
        # --- wss_sem_kitti_loader.py:124-132
        proj_eval_label = np.zeros((proj_eval_mask.shape[0], proj_eval_mask.shape[1]), dtype=np.float32)
        proj_eval_label[proj_idx > -1] = sem_label[proj_idx[proj_idx > -1]]
        proj_train_label = np.zeros((proj_eval_mask.shape[0], proj_eval_mask.shape[1]), dtype=np.float32)
        proj_train_label[proj_idx > -1] = weak_label[proj_idx[proj_idx > -1]]
        # --- wss_sem_kitti_loader.py:159-172
        proj_range_tensor = torch.from_numpy(proj_range)
        proj_xyz_tensor = torch.from_numpy(proj_pointcloud[..., :3])
        proj_intensity_tensor = torch.from_numpy(proj_pointcloud[..., 3])
        proj_intensity_tensor = proj_intensity_tensor.ne(-1).float() * proj_intensity_tensor
        proj_feature_tensor = torch.cat([proj_range_tensor.unsqueeze(0), proj_xyz_tensor.permute(2, 0, 1),
                                         proj_intensity_tensor.unsqueeze(0)], 0)
        # --- trainer.py:555-567,599-608 (batch of one)
        input_feature = proj_feature_tensor.unsqueeze(0).clone()
        train_label = torch.from_numpy(proj_train_label).unsqueeze(0).long()
        eval_label = torch.from_numpy(proj_eval_label).unsqueeze(0).long()
        eval_mask = eval_label.gt(0)
        if normalise:
            feature_mean = torch.Tensor(mean).unsqueeze(0).unsqueeze(2).unsqueeze(2)
            feature_std = torch.Tensor(stds).unsqueeze(0).unsqueeze(2).unsqueeze(2)
            input_feature[:, 0:5] = ((input_feature[:, 0:5] - feature_mean) / feature_std
                                     * eval_mask.unsqueeze(1).expand_as(input_feature[:, 0:5]))
        cases[name] = dict(points=pts, sem_label=sem_label, weak_label=weak_label, H=H, W=W,
                           fov_up=shp.fov_up, fov_down=shp.fov_down, normalise=normalise,
                           img_mean=np.asarray(mean, np.float32), img_std=np.asarray(stds, np.float32),
                           feature=input_feature[0].numpy(), train_label=train_label[0].numpy(),
                           eval_label=eval_label[0].numpy(), proj_range=proj_range, proj_idx=proj_idx)
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "assemble.npz"), **flat)
    print("assemble: ok")


def gold_unproject(pcp):
    """trainer.py:714-728 with the reference's IOUEval (metrics/iou_eval.py) on two scans."""
    from coarse3d_b200 import synth
    RP = pcp.dataset.preprocess.projection.RangeProjection
    shp, H, W, C = synth.KITTI, 32, 256, 20
    evaluator = pcp.metrics.IOUEval(n_classes=C, device=torch.device("cpu"), ignore=[0])
    g = torch.Generator().manual_seed(61)
    argmax_2d = torch.randint(0, C, (2, H, W), generator=g)
    px, py, lab, un, offs = [], [], [], [], [0]
    for ii in range(2):
        pts, full, _ = synth.make_scan(shp, 61 + ii, 4000 + 500 * ii)
        rp = RP(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=H, proj_w=W)
        rp.doProjection(pts)
        uproj_x_idx = torch.from_numpy(rp.cached_data["uproj_x_idx"]).long()   # trainer.py:618-619
        uproj_y_idx = torch.from_numpy(rp.cached_data["uproj_y_idx"]).long()
        unproj_full_labels = torch.from_numpy(full).long()
        unproj_argmax = argmax_2d[ii, uproj_y_idx, uproj_x_idx]                # trainer.py:719
        evaluator.addBatch(unproj_argmax, unproj_full_labels)                  # trainer.py:728
        px.append(uproj_x_idx), py.append(uproj_y_idx), lab.append(unproj_full_labels)
        un.append(unproj_argmax), offs.append(offs[-1] + len(pts))
    np.savez_compressed(os.path.join(OUT, "unproject.npz"), **{
        "two_scans/argmax_2d": argmax_2d.numpy(), "two_scans/px": torch.cat(px).numpy(),
        "two_scans/py": torch.cat(py).numpy(), "two_scans/labels": torch.cat(lab).numpy(),
        "two_scans/offsets": np.asarray(offs, np.int32), "two_scans/nclasses": np.asarray(C),
        "two_scans/unproj_argmax": torch.cat(un).numpy(),
        "two_scans/conf_matrix": evaluator.conf_matrix.numpy()})
    print("unproject: ok")


def gold_entropy_select(pcp):
    """Trainer.entropy_based_selection (trainer.py:447-518) executed from the reference's
    trainer module; the Exp(1) draws of its torch.multinomial calls are replayed from the
    seed (multinomial without replacement == topk(w / q), verified below)."""
    sys.path.insert(0, os.path.join(REF, "tasks", "weak_segmentation"))
    import importlib
    trainer = importlib.import_module("trainer")
    w = torch.rand(3000)
    torch.manual_seed(5); a = torch.multinomial(w, 400, replacement=False)
    torch.manual_seed(5); b = torch.topk(w / torch.empty_like(w).exponential_(1), 400)[1]
    assert torch.equal(a, b), "torch.multinomial(replacement=False) is no longer topk(w / Exp(1))"
    cases = {}
    for name, B, C, H, W, ratio, seed in [("small", 2, 6, 8, 64, 0.37, 71), ("kitti_like", 2, 20, 16, 256, 0.21, 72)]:
        g = torch.Generator().manual_seed(seed)
        output = torch.softmax(torch.randn(B, C, H, W, generator=g) * 1.5, 1)
        eval_mask = torch.rand(B, H, W, generator=g) < 0.8
        full = torch.randint(1, C, (B, H, W), generator=g)
        wss_mask = (torch.rand(B, H, W, generator=g) < 0.02) & eval_mask
        train_label = full * wss_mask
        train_label[1][train_label[1] == 2] = 0          # a class absent from scan 1's weak labels
        wss_mask = train_label.gt(0)                     # trainer.py:602
        fake = types.SimpleNamespace(settings=types.SimpleNamespace(ignore_cls=0, n_classes=C))
        draws = []
        real = torch.multinomial

        def rec(wt, n, replacement=False):
            state = torch.get_rng_state()
            r = real(wt, n, replacement=replacement)
            after = torch.get_rng_state()
            torch.set_rng_state(state)
            draws.append(torch.empty_like(wt).exponential_(1))   # the same draws multinomial made
            torch.set_rng_state(after)
            return r

        torch.multinomial = rec
        try:
            torch.manual_seed(seed)
            label, mask = trainer.Trainer.entropy_based_selection(
                fake, output.clone(), wss_mask, eval_mask, train_label, ratio)
        finally:
            torch.multinomial = real
        # place the recorded draws at [b, cls] in the loop's order (trainer.py:473-496)
        noise = torch.ones(B, C, H * W)
        it = iter(draws)
        pseudo = torch.max(output, dim=1)[1]
        pseudo[eval_mask == False] = 0                                           # noqa: E712
        for b_ in range(B):
            for cls in torch.unique(train_label[b_]):
                if cls == 0:
                    continue
                cm = (pseudo[b_] == cls) * (eval_mask[b_] > 0)
                if cm.sum() == 0 or int(cm.sum() * ratio) < 1:
                    continue
                noise[b_, int(cls)] = next(it)
        assert next(it, None) is None
        cases[name] = dict(output=output.numpy(), wss_mask=wss_mask.numpy(), eval_mask=eval_mask.numpy(),
                           train_label=train_label.numpy(), select_ratio=np.float64(ratio),
                           noise=noise.numpy(), pseudo_label=label.numpy(), new_wss_mask=mask.numpy())
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "entropy_select.npz"), **flat)
    print("entropy_select: ok", {k: int(c["new_wss_mask"].sum()) for k, c in cases.items()})


def gold_lovasz(pcp):
    """Lovasz_softmax (pc_processor/loss/lovasz_softmax.py:160-179) executed from the
    reference, as the trainer builds it (trainer.py:362-364: ignore=ignore_cls,
    per_image=False, softmax=False); loss and autograd gradient.  Random probabilities have
    no equal errors, so the unstable torch.sort leaves nothing undefined."""
    Lov = pcp.loss.Lovasz_softmax
    cases = {}
    for name, B, C, H, W, frac, classes, seed in [("weak", 2, 20, 16, 256, 0.02, "present", 81),
                                                  ("all_classes", 1, 6, 8, 64, 0.3, "all", 82),
                                                  ("one_class", 2, 5, 8, 32, 0.05, "present", 83)]:
        g = torch.Generator().manual_seed(seed)
        probs = torch.softmax(torch.randn(B, C, H, W, generator=g) * 1.5, 1).requires_grad_(True)
        labels = torch.randint(1, C, (B, H, W), generator=g) * (torch.rand(B, H, W, generator=g) < frac)
        if name == "one_class":
            labels = (labels > 0).long() * 3
        crit = Lov(classes=classes, ignore=0, per_image=False, softmax=False)
        loss = crit(probs, labels)
        loss.backward()
        e = (torch.nn.functional.one_hot(labels, C).permute(0, 3, 1, 2).float() - probs.detach()).abs()
        v = e.permute(0, 2, 3, 1)[labels != 0]
        assert all(len(torch.unique(v[:, c])) == v.shape[0] for c in range(C)), "tied errors in fixture"
        cases[name] = dict(probs=probs.detach().numpy(), labels=labels.numpy(), ignore=np.int64(0),
                           classes_all=np.int64(classes == "all"), loss=loss.detach().numpy(),
                           grad=probs.grad.numpy())
    flat = {f"{k}/{f}": np.asarray(v) for k, c in cases.items() for f, v in c.items()}
    np.savez_compressed(os.path.join(OUT, "lovasz.npz"), **flat)
    print("lovasz: ok", {k: float(c["loss"]) for k, c in cases.items()})


if __name__ == "__main__":
    pcp = import_reference()
    gold_projection(pcp)
    gold_knn(pcp)
    gold_loss(pcp)
    gold_ema(pcp)
    gold_assemble(pcp)
    gold_unproject(pcp)
    gold_entropy_select(pcp)
    gold_lovasz(pcp)
