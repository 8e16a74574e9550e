"""GPU parity: c3d_unproject_confusion_batch against the reference's own
statements -- trainer.py:714-724 (per-scan fancy-index gather) and IOUEval.addBatch
(iou_eval.py:35-58), restated with torch-CPU ops in the test.  Integer work: exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(argmax_2d, px, py, offs, labels, C):
    conf = torch.zeros((C, C)).long()                        # iou_eval.py:29-31
    unproj = []
    for ii in range(argmax_2d.shape[0]):
        lo, hi = int(offs[ii]), int(offs[ii + 1])
        u = argmax_2d[ii, py[lo:hi].long(), px[lo:hi].long()]  # trainer.py:719
        unproj.append(u)
        x_row, y_row = u.reshape(-1).long(), labels[lo:hi].reshape(-1).long()   # iou_eval.py:44-45
        idxs = torch.stack([x_row, y_row], dim=0)
        conf = conf.index_put_(tuple(idxs), torch.ones(idxs.shape[-1]).long(), accumulate=True)  # :56-58
    return torch.cat(unproj), conf


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
