"""GPU parity: c3d_knn_batch against the CPU oracle and the reference golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import knn as oknn
from oracle import projection as oproj

pytestmark = pytest.mark.gpu

KNN = load_golden("knn")


@pytest.mark.parametrize("case", sorted(KNN))
def test_dropin_matches_golden(cuda_device, case):
    from coarse3d_b200.pc_processor.postproc import KNN as KNNModule
    g = KNN[case]
    params = dict(knn=int(g["knn"]), search=int(g["search"]), sigma=float(g["sigma"]),
                  cutoff=float(g["cutoff"]))
    mod = KNNModule(params, int(g["nclasses"]))
    out = mod(torch.from_numpy(g["proj_range"]).cuda(), torch.from_numpy(g["unproj_range"]).cuda(),
              torch.from_numpy(g["proj_argmax"]).cuda(), torch.from_numpy(g["px"]).cuda(),
              torch.from_numpy(g["py"]).cuda())
    assert out.dtype == torch.int64 and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), g["out"])


def test_even_window_raises(cuda_device):
    from coarse3d_b200.pc_processor.postproc import KNN as KNNModule
    mod = KNNModule(dict(knn=3, search=4, sigma=1.0, cutoff=1.0), 5)
    z = torch.zeros((4, 4), device="cuda")
    with pytest.raises(ValueError):
        mod(z, torch.zeros(1, device="cuda"), z.long(), torch.zeros(1, device="cuda").long(),
            torch.zeros(1, device="cuda").long())


def _make_batch(shape_name, batch, seed0, quantize=None):
    from coarse3d_b200 import synth
    shp = synth.SHAPES[shape_name]
    pts, offs, _, _ = synth.make_batch(shp, batch, seed0=seed0, ragged=True)
    fov = oproj.Fov(fov_up=shp.fov_up, fov_down=shp.fov_down, proj_h=shp.proj_h, proj_w=shp.proj_w)
    rng = np.random.default_rng(seed0)
    per = []
    for b in range(batch):
        p = pts[offs[b]:offs[b + 1]]
        d = None
        if quantize:  # coarse depths => many exactly equal distances (tie rule)
            d = (np.round(oproj.depth_of(p) / quantize) * quantize + quantize).astype(np.float32)
        o = oproj.project(p, fov, d)
        argmax = rng.integers(0, shp.n_classes, (shp.proj_h, shp.proj_w))
        per.append((o, argmax))
    return shp, offs, per


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("shape_name,batch,k,s,sigma,cutoff,quant", [
    ("kitti", 2, 5, 5, 1.0, 1.0, None),
    ("poss", 3, 5, 5, 1.0, 1.0, None),     # BASELINE config 4: 40x1800, 5x5
    ("nuscenes", 4, 3, 3, 2.0, 0.0, None),  # cutoff disabled: inf slots vote
    ("nuscenes", 2, 7, 7, 1.5, 2.0, 0.5),  # quantised depths: distance ties
    ("kitti", 1, 9, 5, 1.0, 0.7, 0.25),
])
def test_batched_matches_oracle(cuda_device, shape_name, batch, k, s, sigma, cutoff, quant, idt):
    from coarse3d_b200 import ops
    shp, offs, per = _make_batch(shape_name, batch, 300, quant)
    cat = lambda key, dt: torch.from_numpy(np.concatenate([o[key] for o, _ in per]).astype(dt)).cuda()
    npdt = np.int64 if idt == torch.int64 else np.int32
    proj_range = torch.from_numpy(np.stack([o["proj_range"] for o, _ in per])).cuda()
    proj_argmax = torch.from_numpy(np.stack([a for _, a in per]).astype(npdt)).cuda()
    out = ops.knn_batch(proj_range, proj_argmax, cat("uproj_depth", np.float32),
                        cat("uproj_x_idx", npdt), cat("uproj_y_idx", npdt),
                        torch.from_numpy(offs).cuda(), k, s, sigma, cutoff, shp.n_classes)
    assert out.dtype == idt
    # the binned form (c3d_knn_sort_points + records): same labels at the original point indices,
    # also with a co-scheduled fill and the uint8 output
    px, py, ur = cat("uproj_x_idx", npdt), cat("uproj_y_idx", npdt), cat("uproj_depth", np.float32)
    O = torch.from_numpy(offs).cuda()
    rec = ops.knn_sort_points(ur, px, py, O, shp.proj_h, shp.proj_w)
    assert sorted(rec[:, 3].view(torch.int32).tolist()) == list(range(px.numel()))     # a permutation
    seg = rec[:, 2].view(torch.int32).long() * ((shp.proj_w + 31) // 32) + (rec[:, 1].view(torch.int32).long() >> 5)
    for b in range(batch):                      # scan-major, bins ascending inside a scan
        sb = seg[offs[b]:offs[b + 1]]
        assert bool((sb[1:] >= sb[:-1]).all())
        idx = rec[offs[b]:offs[b + 1], 3].view(torch.int32)
        assert int(idx.min()) >= offs[b] and int(idx.max()) < offs[b + 1]
    fillbuf = torch.ones(3001 * 4, device="cuda")
    out2 = ops.knn_batch(proj_range, proj_argmax, ur, px, py, O, k, s, sigma, cutoff, shp.n_classes,
                         records=rec, cofill=fillbuf)
    assert torch.equal(out2, out) and float(fillbuf.abs().max()) == 0.0
    out3 = ops.knn_batch(proj_range, proj_argmax, ur, px, py, O, k, s, sigma, cutoff, shp.n_classes,
                         records=rec, out_uint8=True)
    assert torch.equal(out3.to(idt), out)
    out = out.cpu().numpy()
    for b, (o, argmax) in enumerate(per):
        want = oknn.knn_vote(o["proj_range"], o["uproj_depth"], argmax, o["uproj_x_idx"],
                             o["uproj_y_idx"], k, s, sigma, cutoff, shp.n_classes)
        got = out[offs[b]:offs[b + 1]]
        assert np.array_equal(got, want), (b, int((got != want).sum()))
        assert got.min() >= 1 and got.max() <= shp.n_classes - 1


def test_uniform_labels_are_a_fixed_point(cuda_device):
    """Size-independent property at full KITTI size: if every pixel predicts class c,
    every point is labelled c (the centre slot always votes with distance 0)."""
    from coarse3d_b200 import ops, synth
    shp = synth.KITTI
    pts, offs, _, _ = synth.make_batch(shp, 4, seed0=2000)
    P, O = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    pr = ops.project_batch(P, O, ops.Fov.from_degrees(shp.fov_up, shp.fov_down), shp.proj_h, shp.proj_w)
    for c in (1, 7, shp.n_classes - 1):
        am = torch.full(pr.proj_idx.shape, c, dtype=torch.int32, device="cuda")
        out = ops.knn_batch(pr.proj_range, am, pr.uproj_depth, pr.uproj_x_idx, pr.uproj_y_idx, O,
                            5, 5, 1.0, 1.0, shp.n_classes)
        assert (out == c).all()


@pytest.mark.parametrize("W", [101, 64, 7])
@pytest.mark.parametrize("k,s", [(5, 5), (3, 3), (12, 5), (9, 3), (20, 7), (5, 9)])
def test_random_images_any_width(cuda_device, W, k, s):
    """Widths that are not a multiple of 4 take the scalar-gather path; tiny images put
    every window across the border; large k exercises the generic top-k capacity."""
    from coarse3d_b200 import ops
    rng = np.random.default_rng(W * 100 + k)
    H, P, C = 9, 3000, 11
    proj_range = rng.uniform(1, 30, (H, W)).astype(np.float32)
    proj_range = (np.round(proj_range * 2) / 2).astype(np.float32)      # ties
    proj_range[rng.random((H, W)) < 0.3] = -1.0                          # empty pixels
    argmax = rng.integers(0, C, (H, W))
    px, py = rng.integers(0, W, P), rng.integers(0, H, P)
    ur = (np.round(rng.uniform(1, 30, P) * 2) / 2).astype(np.float32)
    for cutoff in (1.0, 0.0):
        want = oknn.knn_vote(proj_range, ur, argmax, px, py, k, s, 1.0, cutoff, C)
        out = ops.knn_batch(torch.from_numpy(proj_range[None]).cuda(), torch.from_numpy(argmax[None]).cuda(),
                            torch.from_numpy(ur).cuda(), torch.from_numpy(px).cuda(),
                            torch.from_numpy(py).cuda(), torch.tensor([0, P], dtype=torch.int32).cuda(),
                            k, s, 1.0, cutoff, C)
        assert np.array_equal(out.cpu().numpy(), want), int((out.cpu().numpy() != want).sum())


@pytest.mark.parametrize("nfill,W", [(16, 64), (4096 * 7 + 16, 64), (1 << 22, 101), (999 * 16, 7),
                                     (8192 * 10, 64), (8192 * 10 + 48, 64)])
def test_co_scheduled_fill_zeroes_exactly_its_buffer(cuda_device, nfill, W):
    """The vote kernel can carry a zero fill (the loss's dense gradient in the step
    pipeline): labels are unchanged, every byte of the buffer is zero, guards untouched."""
    from coarse3d_b200 import ops
    rng = np.random.default_rng(nfill % 1000 + W)
    H, P, C = 9, 2500, 11
    proj_range = rng.uniform(1, 30, (H, W)).astype(np.float32)
    proj_range[rng.random((H, W)) < 0.3] = -1.0
    argmax = rng.integers(0, C, (H, W))
    px, py = rng.integers(0, W, P), rng.integers(0, H, P)
    ur = rng.uniform(1, 30, P).astype(np.float32)
    args = (torch.from_numpy(proj_range[None]).cuda(), torch.from_numpy(argmax[None]).cuda(),
            torch.from_numpy(ur).cuda(), torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda(),
            torch.tensor([0, P], dtype=torch.int32).cuda(), 5, 5, 1.0, 1.0, C)
    plain = ops.knn_batch(*args)
    guard = 64
    buf = torch.full((nfill + 2 * guard,), 0xAB, dtype=torch.uint8, device="cuda")
    out = ops.knn_batch(*args, cofill=buf[guard:guard + nfill])
    assert torch.equal(out, plain)
    assert int(buf[guard:guard + nfill].max()) == 0
    assert bool((buf[:guard] == 0xAB).all()) and bool((buf[guard + nfill:] == 0xAB).all())
    with pytest.raises(ValueError):
        ops.knn_batch(*args, cofill=buf[guard + 4:guard + 4 + 32])       # not 16 B aligned


def test_co_scheduled_fill_without_points(cuda_device):
    from coarse3d_b200 import ops
    z = torch.zeros((1, 4, 8), device="cuda")
    e = torch.zeros((0,), device="cuda")
    buf = torch.ones(4096, device="cuda")
    ops.knn_batch(z, z.long(), e, e.long(), e.long(), torch.zeros(2, dtype=torch.int32, device="cuda"),
                  5, 5, 1.0, 1.0, 5, cofill=buf)
    assert float(buf.abs().max()) == 0.0


def test_uint8_output_option(cuda_device):
    """label_is_i64 bit 1: uint8 labels out, whatever the dtype of proj_argmax; same values."""
    from coarse3d_b200 import ops
    rng = np.random.default_rng(5)
    H, W, P, C = 16, 128, 3000, 20
    proj_range = rng.uniform(1, 30, (2, H, W)).astype(np.float32)
    proj_range[rng.random((2, H, W)) < 0.3] = -1.0
    argmax = rng.integers(0, C, (2, H, W))
    px, py = rng.integers(0, W, 2 * P), rng.integers(0, H, 2 * P)
    ur = rng.uniform(1, 30, 2 * P).astype(np.float32)
    offs = torch.tensor([0, P, 2 * P], dtype=torch.int32).cuda()
    for adt in (torch.int64, torch.int32):
        args = (torch.from_numpy(proj_range).cuda(), torch.from_numpy(argmax).to(adt).cuda(),
                torch.from_numpy(ur).cuda(), torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda(), offs,
                5, 5, 1.0, 1.0, C)
        want = ops.knn_batch(*args)
        got = ops.knn_batch(*args, out_uint8=True)
        assert got.dtype == torch.uint8 and torch.equal(got.long(), want.long())
    with pytest.raises(ValueError):
        ops.knn_batch(*args[:-1], 300, out_uint8=True)
