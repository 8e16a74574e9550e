"""The drop-in against the REAL reference package (not a fake module tree).

The reference is Python and cannot travel to the GPU box, so these tests run where a
reference tree is mounted (`/root/reference` in the build container, or `baseline/_ref`) and
skip elsewhere.  The CPU part checks that `install()` rebinds every hot-path name of the real
`pc_processor` and of the task's `Trainer`, with call-compatible signatures; the GPU part (needs
both a reference tree and a CUDA device) drives the call sequences of the real callers:
wss_sem_kitti_loader.py:117-147 (two doProjection calls + cached_data), trainer.py:366-375 and
:680-686 (criterion construction, `.cuda()`, keyword call, backward), and the prototype block of
salsanext_proto.py:497-527 through the real model class's rebound method."""
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import baselines  # noqa: E402

REF = baselines.find_reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference tree on this machine")


def _real_package(cpu):
    saved_t, saved_m = torch.Tensor.cuda, torch.nn.Module.cuda
    pcp = baselines.import_reference(REF, cpu=cpu)
    return pcp, (saved_t, saved_m)


def _reference_signatures():
    """Signatures of the reference's own symbols, read before install() rebinds them."""
    spec = importlib.util.spec_from_file_location(
        "_ref_salsanext", os.path.join(REF, "pc_processor", "models", "salsanext_proto.py"))
    return spec


@needs_ref
def test_install_rebinds_the_real_package():
    import coarse3d_b200
    from coarse3d_b200.pc_processor.models import momentum_update, prototype_learning
    pcp, saved = _real_package(cpu=True)
    try:
        ref_pl = inspect.signature(pcp.models.salsanext_proto.SalsaNextProto.prototype_learning)
        ref_mu = inspect.signature(pcp.models.salsanext_proto.momentum_update)
        ref_loss_init = inspect.signature(pcp.loss.contrast_pixel_loss.ContrastMEMLoss.__init__)
        ref_loss_fwd = inspect.signature(pcp.loss.contrast_pixel_loss.ContrastMEMLoss.forward)
        ref_rp_init = inspect.signature(pcp.dataset.preprocess.projection.RangeProjection.__init__)
        ref_rp_call = inspect.signature(pcp.dataset.preprocess.projection.RangeProjection.doProjection)
        ref_knn_init = inspect.signature(pcp.postproc.knn.KNN.__init__)
        ref_knn_fwd = inspect.signature(pcp.postproc.knn.KNN.forward)
        ref_lov_init = inspect.signature(pcp.loss.lovasz_softmax.Lovasz_softmax.__init__)

        coarse3d_b200.install(pcp)

        def same_leading(ref_sig, new_sig):
            """every reference parameter exists, same position, same default"""
            ref_p, new_p = list(ref_sig.parameters.values()), list(new_sig.parameters.values())
            assert len(new_p) >= len(ref_p)
            for r, n in zip(ref_p, new_p):
                assert r.name == n.name and r.default == n.default, (r, n)

        for mod, cls in (("salsanext_proto", "SalsaNextProto"), ("rangenet_proto", "RangeNetProto"),
                         ("squeezesegv3_Proto", "SqueezeSegV3Proto")):
            m = getattr(pcp.models, mod)
            assert getattr(m, cls).prototype_learning is prototype_learning
            assert m.momentum_update is momentum_update
            assert getattr(pcp.models, cls) is getattr(m, cls)          # models/__init__.py re-export
        same_leading(ref_pl, inspect.signature(prototype_learning))
        same_leading(ref_mu, inspect.signature(momentum_update))
        same_leading(ref_loss_init, inspect.signature(pcp.loss.ContrastMEMLoss.__init__))
        same_leading(ref_loss_fwd, inspect.signature(pcp.loss.ContrastMEMLoss.forward))
        same_leading(ref_rp_init, inspect.signature(pcp.dataset.preprocess.RangeProjection.__init__))
        same_leading(ref_rp_call, inspect.signature(pcp.dataset.preprocess.RangeProjection.doProjection))
        same_leading(ref_knn_init, inspect.signature(pcp.postproc.KNN.__init__))
        same_leading(ref_knn_fwd, inspect.signature(pcp.postproc.KNN.forward))
        same_leading(ref_lov_init, inspect.signature(pcp.loss.Lovasz_softmax.__init__))
        # the loaders resolve the class at call time through the module attribute
        assert pcp.dataset.preprocess.projection.RangeProjection is pcp.dataset.preprocess.RangeProjection
        assert pcp.loss.contrast_pixel_loss.ContrastMEMLoss is pcp.loss.ContrastMEMLoss
        # trainer.py:366-371: the criterion the task builds
        crit = pcp.loss.ContrastMEMLoss(ignore_label=0, temperature=0.07, num_anchor=512)
        assert isinstance(crit, torch.nn.Module) and crit.num_anchor == 512
        # wss_sem_kitti_loader.py:76-84: the projection the loader builds
        rp = pcp.dataset.preprocess.projection.RangeProjection(
            fov_up=3, fov_down=-25, fov_left=-180, fov_right=180, proj_h=64, proj_w=2048)
        assert rp.cached_data == {} and rp.proj_w == 2048
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = saved


@needs_ref
def test_install_rebinds_the_real_trainer():
    """tasks/weak_segmentation/trainer.py `Trainer.entropy_based_selection` (:447-518)."""
    import coarse3d_b200
    from coarse3d_b200.trainer_ops import entropy_based_selection
    pcp, saved = _real_package(cpu=True)
    try:
        tdir = os.path.join(REF, "tasks", "weak_segmentation")
        sys.path.insert(0, tdir)
        try:
            spec = importlib.util.spec_from_file_location("_ref_trainer", os.path.join(tdir, "trainer.py"))
            trainer = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(trainer)
        finally:
            sys.path.remove(tdir)
        ref_sig = inspect.signature(trainer.Trainer.entropy_based_selection)
        coarse3d_b200.install(pcp, trainer_cls=trainer.Trainer)
        assert trainer.Trainer.entropy_based_selection is entropy_based_selection
        assert list(ref_sig.parameters) == list(inspect.signature(entropy_based_selection).parameters)
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = saved


@needs_ref
@pytest.mark.gpu
def test_real_callers_run_on_the_b200(cuda_device):
    """The real callers' statement sequences with the rebound operators, checked against the
    reference's own operators (imported a second time, unpatched, by file path)."""
    import types
    import coarse3d_b200
    from coarse3d_b200 import synth
    pcp, saved = _real_package(cpu=False)
    try:
        def load(rel, name):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        ref_proj = load("pc_processor/dataset/preprocess/projection.py", "_ref_projection")
        coarse3d_b200.install(pcp)
        shp = synth.KITTI
        pts, full, weak = synth.make_scan(shp, 7, 20000)
        H, W = 32, 512
        kw = dict(fov_up=shp.fov_up, fov_down=shp.fov_down, fov_left=-180, fov_right=180, proj_h=H, proj_w=W)
        projection = pcp.dataset.preprocess.projection.RangeProjection(**kw)        # loader :76-84
        want = ref_proj.RangeProjection(**kw)
        # ---- wss_sem_kitti_loader.py:117-147
        got = projection.doProjection(pts)
        exp = want.doProjection(pts)
        for a, b in zip(got, exp):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        for k in ("uproj_x_idx", "uproj_y_idx", "uproj_depth"):
            assert np.array_equal(projection.cached_data[k], want.cached_data[k])
        depth_temp = np.linalg.norm(pts[:, :3], 2, axis=1)                          # :134-140
        far = np.nonzero(weak == 0)[0]
        depth_temp[far] = (10000.0 + np.arange(far.size)).astype(np.float32)        # tie-free far depths
        got2 = projection.doProjection(pts, depth_temp)
        exp2 = want.doProjection(pts, depth_temp)
        for a, b in zip(got2, exp2):
            assert np.array_equal(a, b)
        proj_idx = got[2]
        label_img = np.zeros((H, W), np.int64)
        label_img[proj_idx > -1] = weak[proj_idx[proj_idx > -1]]                     # :124-132

        # ---- model.forward's prototype block through the REAL class's rebound method (:497-527)
        C, M, D = shp.n_classes, 4, 32
        g = torch.Generator().manual_seed(3)
        feat_2d = torch.randn(1, D, H, W, generator=g).cuda()
        model = types.SimpleNamespace(
            prototypes=torch.nn.Parameter(torch.nn.functional.normalize(torch.randn(C, M, D, generator=g), dim=-1).cuda(),
                                          requires_grad=False),
            nclasses=C, ignore_label=0, sub_proto_size=M, proto_mom=0.999, deterministic=True)
        ln_d, ln_c = torch.nn.LayerNorm(D).cuda(), torch.nn.LayerNorm(C).cuda()
        label = torch.from_numpy(label_img)[None].cuda()
        with torch.no_grad():
            out_feat = pcp.models.salsanext_proto.l2_normalize(ln_d(feat_2d.permute(0, 2, 3, 1).reshape(-1, D)))
            sim = torch.einsum("nd,kmd->nmk", out_feat, model.prototypes)
            nearest = ln_c(torch.amax(sim, dim=1)).view(1, H, W, C).permute(0, 3, 1, 2).contiguous()
            logits, target = pcp.models.SalsaNextProto.prototype_learning(
                model, out_feat, nearest, label.view(-1), (label > 0).view(-1), sim)
        assert logits.shape == (H * W, M * C) and target.shape == (H * W,)

        # ---- trainer.py:366-375 and :675-690
        crit = pcp.loss.ContrastMEMLoss(ignore_label=0, temperature=0.07, num_anchor=512).cuda()
        pred_2d = torch.softmax(torch.randn(1, C, H, W, generator=g), 1).cuda()
        feat_2d.requires_grad_(True)
        proto_queue = model.prototypes.detach().unsqueeze(0)
        contrast_loss = crit(feats=feat_2d, output=pred_2d, labels=label, keep_mask=label.gt(0),
                             proto_queue=proto_queue)
        total_loss = (torch.tensor(0.0).cuda() + 0.1 * contrast_loss).mean()
        total_loss.backward()
        assert torch.isfinite(contrast_loss) and feat_2d.grad is not None
        assert int((feat_2d.grad.abs().sum(1) > 0).sum()) > 0
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = saved
