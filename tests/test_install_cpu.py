"""CPU: the drop-in surface -- install() rebinds the reference's names, the mirror classes
keep the reference's constructor signatures, attributes and assertions."""
import inspect
import types

import pytest


def _fake_pc_processor():
    def mod(name):
        return types.ModuleType(name)
    pcp = mod("pc_processor")
    pcp.dataset = mod("pc_processor.dataset")
    pcp.dataset.preprocess = mod("pc_processor.dataset.preprocess")
    pcp.dataset.preprocess.projection = mod("pc_processor.dataset.preprocess.projection")
    pcp.loss = mod("pc_processor.loss")
    pcp.loss.contrast_pixel_loss = mod("pc_processor.loss.contrast_pixel_loss")
    pcp.postproc = mod("pc_processor.postproc")
    pcp.postproc.knn = mod("pc_processor.postproc.knn")
    for m in (pcp.dataset.preprocess.projection, pcp.dataset.preprocess):
        m.RangeProjection = object
    for m in (pcp.loss.contrast_pixel_loss, pcp.loss):
        m.ContrastMEMLoss = object
    for m in (pcp.postproc.knn, pcp.postproc):
        m.KNN = object
    return pcp


def test_install_rebinds_every_alias():
    import coarse3d_b200
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    from coarse3d_b200.pc_processor.loss import ContrastMEMLoss
    from coarse3d_b200.pc_processor.postproc import KNN
    pcp = coarse3d_b200.install(_fake_pc_processor())
    assert pcp.dataset.preprocess.projection.RangeProjection is RangeProjection   # loaders' import path
    assert pcp.dataset.preprocess.RangeProjection is RangeProjection              # preprocess/__init__.py:2
    assert pcp.loss.ContrastMEMLoss is ContrastMEMLoss                             # trainer.py:366
    assert pcp.loss.contrast_pixel_loss.ContrastMEMLoss is ContrastMEMLoss
    assert pcp.postproc.KNN is KNN and pcp.postproc.knn.KNN is KNN


def test_signatures_match_the_reference():
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    from coarse3d_b200.pc_processor.loss import ContrastMEMLoss
    from coarse3d_b200.pc_processor.models import PrototypeBank, momentum_update
    from coarse3d_b200.pc_processor.postproc import KNN
    p = inspect.signature(RangeProjection.__init__).parameters
    assert list(p)[1:7] == ["fov_up", "fov_down", "proj_w", "proj_h", "fov_left", "fov_right"]  # projection.py:7-15
    assert [p[k].default for k in list(p)[1:7]] == [3, -25, 512, 64, -180, 180]
    assert list(inspect.signature(RangeProjection.doProjection).parameters)[1:] == ["pointcloud", "depth"]
    p = inspect.signature(ContrastMEMLoss.__init__).parameters                      # contrast_pixel_loss.py:9-16
    assert [(k, p[k].default) for k in list(p)[1:]] == [("ignore_label", 0), ("temperature", 0.1),
                                                        ("base_temperature", 0.07), ("num_anchor", 50),
                                                        ("is_debug", False)]
    f = list(inspect.signature(ContrastMEMLoss.forward).parameters)[1:6]
    assert f == ["feats", "output", "labels", "keep_mask", "proto_queue"]           # :27-34
    assert list(inspect.signature(KNN.__init__).parameters)[1:3] == ["params", "nclasses"]      # knn.py:37
    assert list(inspect.signature(KNN.forward).parameters)[1:] == ["proj_range", "unproj_range", "proj_argmax",
                                                                   "px", "py"]       # knn.py:54
    assert list(inspect.signature(momentum_update).parameters) == ["old_value", "new_value", "momentum", "debug"]
    bank = PrototypeBank(nclasses=6, sub_proto_size=4, proj_dim=8)
    assert bank.prototypes.shape == (6, 4, 8) and not bank.prototypes.requires_grad    # salsanext_proto.py:322-324
    assert "prototypes" in bank.state_dict() and "feat_norm.weight" in bank.state_dict()


def test_constructor_behaviour_without_a_gpu():
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    from coarse3d_b200.pc_processor.postproc import KNN
    with pytest.raises(AssertionError):
        RangeProjection(fov_up=-1)            # projection.py:17-21
    with pytest.raises(AssertionError):
        RangeProjection(fov_left=1)           # projection.py:22-26
    rp = RangeProjection(fov_up=3, fov_down=-25, proj_h=64, proj_w=2048)
    assert abs(rp.fov_vert - 0.4886921905584123) < 1e-12 and rp.cached_data == {}
    k = KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), 20)
    assert (k.knn, k.search, k.sigma, k.cutoff, k.nclasses) == (5, 5, 1.0, 1.0, 20)


def test_install_rebinds_trainer_selection():
    import coarse3d_b200
    from coarse3d_b200.trainer_ops import entropy_based_selection

    class Trainer:
        def entropy_based_selection(self, output, wss_mask, eval_mask, train_label, select_ratio):
            raise AssertionError("reference path")

    coarse3d_b200.install(_fake_pc_processor(), trainer_cls=Trainer)
    assert Trainer.entropy_based_selection is entropy_based_selection
    assert list(inspect.signature(entropy_based_selection).parameters) == [
        "self", "output", "wss_mask", "eval_mask", "train_label", "select_ratio"]      # trainer.py:447-454
