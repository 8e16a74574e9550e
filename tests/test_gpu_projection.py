"""GPU parity: c3d_project_batch (through the reference-shaped RangeProjection and
the batched tensor API) against the CPU oracle and the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import projection as oproj

pytestmark = pytest.mark.gpu

PROJ = load_golden("projection")
FIELDS = ("proj_pointcloud", "proj_range", "proj_idx", "proj_mask", "uproj_x_idx", "uproj_y_idx",
          "uproj_depth")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _run_dropin(points, depth, fov_up, fov_down, H, W):
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    rp = RangeProjection(fov_up=fov_up, fov_down=fov_down, proj_h=H, proj_w=W)
    pc, rng, idx, mask = rp.doProjection(points, depth)
    return dict(proj_pointcloud=pc, proj_range=rng, proj_idx=idx, proj_mask=mask,
                uproj_x_idx=rp.cached_data["uproj_x_idx"], uproj_y_idx=rp.cached_data["uproj_y_idx"],
                uproj_depth=rp.cached_data["uproj_depth"])


@pytest.mark.parametrize("case", sorted(PROJ))
def test_dropin_matches_oracle_and_golden(cuda_device, case):
    g = PROJ[case]
    H, W = int(g["H"]), int(g["W"])
    depth = g["depth"] if bool(g["has_depth"]) else None
    out = _run_dropin(g["points"], depth, float(g["fov_up"]), float(g["fov_down"]), H, W)
    fov = oproj.Fov(fov_up=float(g["fov_up"]), fov_down=float(g["fov_down"]), proj_h=H, proj_w=W)
    o = oproj.project(g["points"], fov, depth)
    for k in FIELDS:  # bit-exact against the oracle, dtype included
        assert out[k].dtype == o[k].dtype, k
        assert out[k].shape == o[k].shape, k
        assert np.array_equal(_bits(out[k]), _bits(o[k])), k
    # against the reference's own output: only boundary-ambiguous points may differ
    bad = (out["uproj_x_idx"] != g["uproj_x_idx"]) | (out["uproj_y_idx"] != g["uproj_y_idx"])
    if bad.any():
        amb = oproj.pixel_is_boundary_ambiguous(g["points"], fov, depth)
        assert not (bad & ~amb).any()
    elif int(g["n_depth_ties"]) == 0:
        for k in FIELDS:
            assert np.array_equal(_bits(out[k]), _bits(g[k])), k


@pytest.mark.parametrize("f64_only", [False, True])
@pytest.mark.parametrize("shape_name,batch", [("kitti", 3), ("nuscenes", 5), ("poss", 2)])
def test_batched_ragged_matches_oracle(cuda_device, shape_name, batch, f64_only):
    from coarse3d_b200 import ops, synth
    shp = synth.SHAPES[shape_name]
    pts, offs, _, _ = synth.make_batch(shp, batch, seed0=100, ragged=True)
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    out = ops.project_batch(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(), fov,
                            shp.proj_h, shp.proj_w, exact_f64=f64_only)
    torch.cuda.synchronize()
    assert int(out.flags.item()) == 0
    ofov = oproj.Fov(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=shp.proj_h, proj_w=shp.proj_w)
    for b in range(batch):
        lo, hi = int(offs[b]), int(offs[b + 1])
        o = oproj.project(pts[lo:hi], ofov)
        assert np.array_equal(out.uproj_x_idx[lo:hi].cpu().numpy(), o["uproj_x_idx"])
        assert np.array_equal(out.uproj_y_idx[lo:hi].cpu().numpy(), o["uproj_y_idx"])
        assert np.array_equal(_bits(out.uproj_depth[lo:hi].cpu().numpy()), _bits(o["uproj_depth"]))
        assert np.array_equal(out.proj_idx[b].cpu().numpy(), o["proj_idx"])
        assert np.array_equal(out.proj_mask[b].cpu().numpy(), o["proj_mask"])
        assert np.array_equal(_bits(out.proj_range[b].cpu().numpy()), _bits(o["proj_range"]))
        assert np.array_equal(_bits(out.proj_pointcloud[b].cpu().numpy()), _bits(o["proj_pointcloud"]))


@pytest.mark.parametrize("c_in", [3, 5])
def test_generic_channel_count(cuda_device, c_in):
    from coarse3d_b200 import synth
    pts, _, _ = synth.make_scan(synth.KITTI, 7, 5000)
    if c_in == 3:
        pts = np.ascontiguousarray(pts[:, :3])
    else:
        pts = np.concatenate([pts, pts[:, 3:4] * 2], 1)
    out = _run_dropin(pts, None, 3.0, -25.0, 32, 256)
    o = oproj.project(pts, oproj.Fov(proj_h=32, proj_w=256))
    for k in FIELDS:
        assert np.array_equal(_bits(out[k]), _bits(o[k])), k


def test_depth_ties_follow_the_rule(cuda_device):
    from coarse3d_b200 import synth
    pts, _, _ = synth.make_scan(synth.KITTI, 8, 4000)
    pts = np.concatenate([pts, pts[::-1], pts[::3]], 0)  # exact duplicates => depth ties
    out = _run_dropin(pts, None, 3.0, -25.0, 16, 128)
    o = oproj.project(pts, oproj.Fov(proj_h=16, proj_w=128))
    for k in FIELDS:
        assert np.array_equal(_bits(out[k]), _bits(o[k])), k


def test_point_zero_mask_quirk_and_empty_scan(cuda_device):
    pts = np.array([[5, 1, 0, 1], [5, 1, 0, 2]], np.float32)  # same pixel, same depth
    out = _run_dropin(pts, None, 3.0, -25.0, 4, 8)
    assert out["proj_idx"].max() == 0 and out["proj_mask"].sum() == 0  # projection.py:113
    empty = _run_dropin(np.zeros((0, 4), np.float32), None, 3.0, -25.0, 4, 8)
    assert (empty["proj_idx"] == -1).all() and (empty["proj_range"] == -1).all()
    assert empty["uproj_x_idx"].shape == (0,)


def test_zero_depth_is_loud(cuda_device):
    pts = np.zeros((3, 4), np.float32)
    with pytest.raises(ValueError):
        _run_dropin(pts, None, 3.0, -25.0, 4, 8)


def test_constructor_asserts(cuda_device):
    from coarse3d_b200.pc_processor.dataset.preprocess import RangeProjection
    with pytest.raises(AssertionError):
        RangeProjection(fov_up=-1)
    with pytest.raises(AssertionError):
        RangeProjection(fov_right=-1)


def test_full_size_properties_and_hybrid_equals_f64(cuda_device, monkeypatch):
    """BASELINE config sizes (8 KITTI scans): size-independent properties, and the
    hybrid fast path must give exactly the fp64 path's pixels on every point."""
    from coarse3d_b200 import ops, synth
    shp = synth.KITTI
    pts, offs, _, _ = synth.make_batch(shp, 8, seed0=1000)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    ref = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, exact_f64=True)
    out = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w)
    torch.cuda.synchronize()
    for a, b in zip(out[:7], ref[:7]):
        assert torch.equal(a, b)
    HW = shp.proj_h * shp.proj_w
    scan = torch.repeat_interleave(torch.arange(8, device="cuda"), (O[1:] - O[:-1]).long())
    lin = scan * HW + out.uproj_y_idx.long() * shp.proj_w + out.uproj_x_idx.long()
    # range image == per-pixel minimum depth; empty pixels are -1
    mn = torch.full((8 * HW,), float("inf"), device="cuda").scatter_reduce(
        0, lin, out.uproj_depth, reduce="amin")
    mn[mn == float("inf")] = -1
    assert torch.equal(mn.view(8, shp.proj_h, shp.proj_w), out.proj_range)
    # winners: re-projecting the winning point lands in its own pixel, with that depth
    valid = out.proj_idx >= 0
    gidx = (out.proj_idx.long() + O[:-1].long().view(8, 1, 1))[valid]
    assert torch.equal(lin[gidx], torch.nonzero(valid.view(-1)).view(-1))
    assert torch.equal(out.uproj_depth[gidx], out.proj_range[valid])
    assert torch.equal(P[gidx], out.proj_pointcloud[valid])
    assert torch.equal(out.proj_mask, (out.proj_idx > 0).int())
    assert (out.proj_pointcloud[~valid] == -1).all()


def test_estimate_guard_band_on_adversarial_points(cuda_device, monkeypatch):
    """Points constructed to sit on / next to pixel boundaries, on the axes, at extreme
    magnitudes and outside the vertical field of view: the fast-transcendental + guard band
    path must give exactly the exact-chain pixels (exact_f64=True) and the oracle's."""
    from coarse3d_b200 import ops
    rng = np.random.default_rng(11)
    H, W, up, down = 64, 2048, 3.0, -25.0
    fov_v = np.deg2rad(up - down)
    n = 200_000
    # yaw exactly on column boundaries (+- a few ulp), pitch exactly on row boundaries
    col = rng.integers(0, W, n)
    yaw = (col / W) * 2 * np.pi - np.pi
    row = rng.integers(0, H, n)
    pitch = (1.0 - row / H) * fov_v - abs(np.deg2rad(down))
    on_row = rng.random(n) < 0.5
    pitch = np.where(on_row, pitch, np.deg2rad(rng.uniform(down - 20, up + 60, n)))
    r = np.exp(rng.uniform(np.log(0.3), np.log(200), n))
    x = r * np.cos(pitch) * np.cos(-yaw)
    y = r * np.cos(pitch) * np.sin(-yaw)
    z = r * np.sin(pitch)
    pts = np.stack([x, y, z, rng.random(n)], 1).astype(np.float32)
    for k in range(1, 4):   # nudge by a few ulp either way
        sel = slice(k, n, 7)
        pts[sel, :3] = np.nextafter(pts[sel, :3], np.float32(np.inf if k % 2 else -np.inf))
    special = np.array([[1, 0, 0, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, -1, 0, 0], [-1, -0.0, 0, 0],
                        [1e-15, 1e-16, 1e-17, 0], [1e-15, -1e-15, 1e-16, 0], [3e18, 1e18, -1e18, 0],
                        [1e25, -1e25, 1e24, 0], [0, 0, 1, 0], [0, 0, -1, 0], [1e-3, 1e-3, 5, 0],
                        [5, 5, -5, 0], [1, 1, 1.4, 0], [1, -1, -1.39, 0]], np.float32)
    pts = np.concatenate([special, pts], 0)
    P = torch.from_numpy(pts).cuda()
    O = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device="cuda")
    fov = ops.Fov.from_degrees(up, down)
    ref = [t.clone() for t in ops.project_batch(P, O, fov, H, W, exact_f64=True)]
    out = ops.project_batch(P, O, fov, H, W)
    for a, b in zip(out, ref):
        assert torch.equal(a, b)
    o = oproj.project(pts, oproj.Fov(fov_up=up, fov_down=down, proj_h=H, proj_w=W))
    assert np.array_equal(out.uproj_x_idx.cpu().numpy(), o["uproj_x_idx"])
    assert np.array_equal(out.uproj_y_idx.cpu().numpy(), o["uproj_y_idx"])
    assert np.array_equal(out.proj_idx[0].cpu().numpy(), o["proj_idx"])


def test_fused_persistent_kernel_equals_two_kernel_form(cuda_device):
    """c3d_project_batch flag bit 2: one persistent launch with a ring z-buffer (work queue,
    per-scan counters) must give the two-kernel form's outputs bit for bit, ragged batches and
    more scans than ring slots included, and leave the workspace reusable."""
    from coarse3d_b200 import ops, synth
    shp = synth.NUSCENES
    pts, offs, _, _ = synth.make_batch(shp, 11, seed0=300, ragged=True)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    ref = [t.clone() for t in ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w)]
    bufs = ops.ProjectionBuffers(11, pts.shape[0], 4, shp.proj_h, shp.proj_w, "cuda")
    for _ in range(3):     # repeated calls reuse the ring left clean by the previous one
        out = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs, fused_kernel=True)
        for a, b in zip(out[:7], ref[:7]):
            assert torch.equal(a, b)
    out = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs)   # and back
    for a, b in zip(out[:7], ref[:7]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape,batch,ragged", [("kitti", 5, True), ("nuscenes", 37, True), ("poss", 3, False)])
def test_cluster_dsmem_form_equals_two_kernel_form(cuda_device, shape, batch, ragged):
    """c3d_project_batch flag bit 3: the scan's z-buffer lives in the distributed shared memory of
    an 8-CTA thread-block cluster (no global z-buffer traffic).  Must give the two-kernel form's
    outputs bit for bit: ragged batches, more scans than resident clusters, depth override, both
    transcendental modes, an empty scan, and leave the global workspace as it was."""
    from coarse3d_b200 import ops, synth
    shp = synth.SHAPES[shape]
    pts, offs, _, _ = synth.make_batch(shp, batch, seed0=900, ragged=ragged)
    offs = offs.copy()
    if batch > 2:                      # scan 1 becomes empty: its points go to scan 2
        offs[2] = offs[1]
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    fov = ops.Fov.from_degrees(shp.fov_up, shp.fov_down)
    g = torch.Generator().manual_seed(3)
    depth = (torch.from_numpy(np.linalg.norm(pts[:, :3], axis=1)) * (0.5 + torch.rand(pts.shape[0], generator=g))).float().cuda()
    assert ops.lib.c3d_project_cluster_supported(4, shp.proj_h, shp.proj_w) == 1
    for kw in ({}, {"depth": depth}, {"exact_f64": True}):
        ref = [t.clone() for t in ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, cluster_kernel=False, **kw)]
        bufs = ops.ProjectionBuffers(batch, pts.shape[0], 4, shp.proj_h, shp.proj_w, "cuda")
        for rep in range(2):
            out = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs, cluster_kernel=True, **kw)
            for a, b in zip(out[:7], ref[:7]):
                assert a.dtype == b.dtype and torch.equal(a, b), (kw.keys(), rep)
        out = ops.project_batch(P, O, fov, shp.proj_h, shp.proj_w, buffers=bufs, cluster_kernel=False, **kw)
        for a, b in zip(out[:7], ref[:7]):
            assert torch.equal(a, b)
