"""GPU parity: c3d_project_assemble_batch (projection fused with the
label-image / input-feature assembly of its caller) against the oracle and the
reference golden vectors.  Everything is bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import assemble as oasm
from oracle import projection as oproj

pytestmark = pytest.mark.gpu

ASM = load_golden("assemble")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.mark.parametrize("case", sorted(ASM))
def test_matches_reference_golden(cuda_device, case):
    from coarse3d_b200 import ops
    g = ASM[case]
    H, W, n = int(g["H"]), int(g["W"]), g["points"].shape[0]
    norm = bool(g["normalise"])
    fov = ops.Fov.from_degrees(float(g["fov_up"]), float(g["fov_down"]))
    out = ops.project_assemble_batch(
        torch.from_numpy(g["points"]).cuda(), torch.tensor([0, n], dtype=torch.int32).cuda(), fov, H, W,
        sem_label=torch.from_numpy(g["sem_label"]).cuda(), weak_label=torch.from_numpy(g["weak_label"]).cuda(),
        img_mean=torch.from_numpy(g["img_mean"]).cuda() if norm else None,
        img_std=torch.from_numpy(g["img_std"]).cuda() if norm else None)
    assert out.train_label.dtype == torch.int64 and out.feature.shape == (1, 5, H, W)
    assert np.array_equal(out.proj_idx[0].cpu().numpy(), g["proj_idx"])
    assert np.array_equal(out.train_label[0].cpu().numpy(), g["train_label"])
    assert np.array_equal(out.eval_label[0].cpu().numpy(), g["eval_label"])
    assert np.array_equal(_bits(out.proj_range[0].cpu().numpy()), _bits(g["proj_range"]))
    assert np.array_equal(_bits(out.feature[0].cpu().numpy()), _bits(g["feature"]))


@pytest.mark.parametrize("ldt", [np.int32, np.uint8])
@pytest.mark.parametrize("normalise", [False, True])
def test_batched_matches_oracle(cuda_device, normalise, ldt):
    from coarse3d_b200 import ops, synth
    shp, B = synth.KITTI, 3
    pts, offs, full, weak = synth.make_batch(shp, B, seed0=700, ragged=True)
    pts[::53, 3] = -1.0
    mean = np.asarray([12.12, 10.88, 0.23, -1.04, 0.21], np.float32)
    std = np.asarray([12.32, 11.47, 6.91, 0.86, 0.16], np.float32)
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    bufs = ops.ProjectionBuffers(B, pts.shape[0], 4, shp.proj_h, shp.proj_w, "cuda")
    for rep in range(2):  # second call runs on the z-buffer the first one left clean
        out = ops.project_assemble_batch(
            torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(), fov, shp.proj_h, shp.proj_w,
            sem_label=torch.from_numpy(full.astype(ldt)).cuda(),
            weak_label=torch.from_numpy(weak.astype(ldt)).cuda(),
            img_mean=torch.from_numpy(mean).cuda() if normalise else None,
            img_std=torch.from_numpy(std).cuda() if normalise else None, buffers=bufs)
        ofov = oproj.Fov(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=shp.proj_h, proj_w=shp.proj_w)
        for b in range(B):
            lo, hi = int(offs[b]), int(offs[b + 1])
            o = oasm.assemble(pts[lo:hi], ofov, full[lo:hi], weak[lo:hi],
                              mean if normalise else None, std if normalise else None)
            assert np.array_equal(_bits(out.feature[b].cpu().numpy()), _bits(o["feature"]))
            assert np.array_equal(out.train_label[b].cpu().numpy(), o["train_label"])
            assert np.array_equal(out.eval_label[b].cpu().numpy(), o["eval_label"])
            assert np.array_equal(out.proj_idx[b].cpu().numpy(), o["proj_idx"])
            assert np.array_equal(out.uproj_x_idx[lo:hi].cpu().numpy(), o["uproj_x_idx"])


def test_consistent_with_plain_projection(cuda_device):
    """Full-size property: the fused pass equals project_batch + gathers done in torch."""
    from coarse3d_b200 import ops, synth
    shp, B = synth.KITTI, 8
    pts, offs, full, weak = synth.make_batch(shp, B, seed0=1000)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    pr = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w)
    out = ops.project_assemble_batch(P, O, fov, shp.proj_h, shp.proj_w,
                                     weak_label=torch.from_numpy(weak.astype(np.int32)).cuda())
    assert out.eval_label is None
    assert torch.equal(out.proj_idx, pr.proj_idx) and torch.equal(out.proj_range, pr.proj_range)
    valid = pr.proj_idx >= 0
    gidx = (pr.proj_idx.long() + O[:-1].long().view(B, 1, 1))[valid]
    want = torch.zeros_like(out.train_label)
    want[valid] = torch.from_numpy(weak).cuda()[gidx]
    assert torch.equal(out.train_label, want)
    assert torch.equal(out.feature[:, 0], pr.proj_range)
    assert torch.equal(out.feature[:, 1:4], pr.proj_pointcloud[..., :3].permute(0, 3, 1, 2))
