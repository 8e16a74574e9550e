"""GPU parity: c3d_unproject_confusion_batch against the oracle
(trainer.py:714-724 + IOUEval.addBatch, iou_eval.py:35-58) and the golden vectors produced
by the reference's IOUEval.  Integer work: exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


from conftest import load_golden
from oracle import unproject as ounp


def _reference(argmax_2d, px, py, offs, labels, C):
    return ounp.unproject_confusion(argmax_2d, px, py, offs, labels, C)


def test_matches_reference_golden(cuda_device):
    from coarse3d_b200 import ops
    g = load_golden("unproject")["two_scans"]
    C = int(g["nclasses"])
    u, conf = ops.unproject_confusion_batch(
        torch.from_numpy(g["argmax_2d"]).cuda(), torch.from_numpy(g["px"]).cuda(),
        torch.from_numpy(g["py"]).cuda(), torch.from_numpy(g["offsets"]).cuda(), C,
        labels=torch.from_numpy(g["labels"]).cuda())
    assert np.array_equal(u.cpu().numpy(), g["unproj_argmax"])
    assert np.array_equal(conf.cpu().numpy(), g["conf_matrix"])


@pytest.mark.parametrize("adt,pdt,ldt", [(torch.int64, torch.int32, torch.int32),
                                         (torch.int64, torch.int64, torch.int64),
                                         (torch.int32, torch.int32, torch.int64)])
def test_matches_reference_statements(cuda_device, adt, pdt, ldt):
    from coarse3d_b200 import ops, synth
    shp, B = synth.KITTI, 3
    pts, offs, full, _ = synth.make_batch(shp, B, seed0=900, ragged=True)
    C = shp.n_classes
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    pr = ops.project_batch(P, O, ops.Fov.from_degrees(shp.fov_up, shp.fov_down), shp.proj_h, shp.proj_w)
    g = torch.Generator().manual_seed(1)
    argmax = torch.randint(0, C, (B, shp.proj_h, shp.proj_w), generator=g)
    px, py = pr.uproj_x_idx.cpu(), pr.uproj_y_idx.cpu()
    labels = torch.from_numpy(full)
    want_u, want_c = _reference(argmax, px, py, offs, labels, C)
    conf = torch.zeros((C, C), dtype=torch.int64, device="cuda")
    for _ in range(2):  # accumulates across calls like the evaluator across iterations
        u, conf = ops.unproject_confusion_batch(argmax.to(adt).cuda(), px.to(pdt).cuda(), py.to(pdt).cuda(),
                                                O, C, labels=labels.to(ldt).cuda(), conf_matrix=conf)
    assert u.dtype == adt and torch.equal(u.cpu().long(), want_u)
    assert torch.equal(conf.cpu(), 2 * want_c)
    assert int(conf.sum()) == 2 * pts.shape[0]


def test_gather_only_feeds_knn(cuda_device):
    from coarse3d_b200 import ops, synth
    shp, B = synth.POSS, 2
    pts, offs, _, _ = synth.make_batch(shp, B, seed0=910)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    pr = ops.project_batch(P, O, ops.Fov.from_degrees(shp.fov_up, shp.fov_down), shp.proj_h, shp.proj_w)
    am = torch.randint(0, shp.n_classes, pr.proj_idx.shape, device="cuda")
    u, conf = ops.unproject_confusion_batch(am, pr.uproj_x_idx, pr.uproj_y_idx, O, shp.n_classes)
    assert conf is None
    scan = torch.repeat_interleave(torch.arange(B, device="cuda"), (O[1:] - O[:-1]).long())
    assert torch.equal(u, am[scan, pr.uproj_y_idx.long(), pr.uproj_x_idx.long()])
