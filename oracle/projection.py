"""Oracle: spherical range projection + min-depth z-buffer (TEST INFRASTRUCTURE).

Restates `RangeProjection` from the reference,
pc_processor/dataset/preprocess/projection.py:4-115, in numpy.

Rules fixed where the reference is undefined or machine dependent:

* Transcendentals.  The reference evaluates `np.arctan2` / `np.arcsin` on
  float32 arrays (projection.py:54-55).  numpy dispatches those to SIMD kernels
  that are not correctly rounded and differ between CPUs (in the build
  container 38 % of arctan2 results are 1-2 ulp off).  The oracle uses the
  *correctly rounded* float32 value (float64 libm result rounded to float32),
  i.e. the IEEE evaluation of the reference's own formula.  Everything after
  the transcendental (projection.py:62-85) is IEEE float32 and is restated
  operation by operation.  `pixel_is_boundary_ambiguous` reports the points
  whose pixel could legitimately differ under a <=2 ulp perturbation of the
  angle; the golden test allows reference/oracle disagreement only there.
* Z-buffer ties.  projection.py:94 uses the default unstable `np.argsort`, so
  among points of equal depth in one pixel the winner is undefined.  Rule:
  minimum depth wins, then minimum point index (what a stable sort would give).
  Depths are ordered by the monotone integer key of their float32 bits so the
  order is total (-0.0 < +0.0).
* `depth == 0` gives NaN pixel coordinates and crashes the reference; the
  oracle raises ValueError.
"""
import numpy as np

F32 = np.float32


class Fov:
    """Constructor arithmetic of RangeProjection.__init__ (projection.py:17-39)."""

    def __init__(self, fov_up=3, fov_down=-25, proj_w=512, proj_h=64,
                 fov_left=-180, fov_right=180):
        assert fov_up >= 0 and fov_down <= 0, \
            "require fov_up >= 0 and fov_down <= 0, while fov_up/fov_down is {}/{}".format(fov_up, fov_down)
        assert fov_right >= 0 and fov_left <= 0, \
            "require fov_right >= 0 and fov_left <= 0, while fov_right/fov_left is {}/{}".format(fov_right, fov_left)
        self.fov_up = fov_up / 180.0 * np.pi
        self.fov_down = fov_down / 180.0 * np.pi
        self.fov_vert = abs(self.fov_up) + abs(self.fov_down)
        self.fov_left = fov_left / 180.0 * np.pi
        self.fov_right = fov_right / 180.0 * np.pi
        self.fov_hori = abs(self.fov_left) + abs(self.fov_right)
        self.proj_w = proj_w
        self.proj_h = proj_h


def depth_of(points):
    """projection.py:47 -- np.linalg.norm(.., 2, axis=1) on float32 is bitwise
    sqrt((x*x + y*y) + z*z) in float32 (SURVEY.md 8a1 [probed])."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    return np.sqrt((x * x + y * y) + z * z)


def angles(points, depth):
    """projection.py:54-55 with correctly rounded float32 transcendentals."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    yaw = -(np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(F32))
    q = z / depth  # float32 division
    with np.errstate(invalid="ignore"):
        pitch = np.arcsin(q.astype(np.float64)).astype(F32)
    return yaw, pitch


def pixel_coords(yaw, pitch, fov):
    """projection.py:62-85, one IEEE float32 operation per reference operation.
    Python-float constants are weak scalars in numpy 2 => rounded to float32."""
    W, H = fov.proj_w, fov.proj_h
    fx = (yaw + F32(abs(fov.fov_left))) / F32(fov.fov_hori)
    fy = F32(1.0) - (pitch + F32(abs(fov.fov_down))) / F32(fov.fov_vert)
    fx = fx * F32(W)
    fy = fy * F32(H)
    px = np.maximum(np.minimum(F32(W - 1), np.floor(fx)), F32(0))
    py = np.maximum(np.minimum(F32(H - 1), np.floor(fy)), F32(0))
    if np.isnan(px).any() or np.isnan(py).any():
        raise ValueError("NaN pixel coordinate (depth == 0 or |z| > depth)")
    return px.astype(np.int32), py.astype(np.int32)


def depth_key(depth):
    """Monotone uint32 key of float32 bits: a < b  <=>  key(a) < key(b)."""
    bits = np.ascontiguousarray(depth, dtype=F32).view(np.uint32)
    neg = (bits >> 31).astype(bool)
    return np.where(neg, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)


def project(points, fov, depth=None):
    """RangeProjection.doProjection (projection.py:43-115).

    Returns a dict with the four returned images and the three cached per-point
    arrays: proj_pointcloud (H,W,C) f32, proj_range (H,W) f32, proj_idx (H,W)
    i32, proj_mask (H,W) i32, uproj_x_idx (N,) i32, uproj_y_idx (N,) i32,
    uproj_depth (N,) f32.
    """
    points = np.ascontiguousarray(points, dtype=F32)
    n, c = points.shape
    H, W = fov.proj_h, fov.proj_w
    depth = depth_of(points) if depth is None else np.asarray(depth, dtype=F32)
    yaw, pitch = angles(points, depth)
    px, py = pixel_coords(yaw, pitch, fov)

    # projection.py:92-99 with the tie rule above: process in decreasing
    # (depth, index) order so that the last write per pixel is the minimum.
    idx = np.arange(n, dtype=np.int64)
    order = np.lexsort((idx, depth_key(depth)))[::-1]
    proj_range = np.full((H, W), -1, dtype=F32)
    proj_range[py[order], px[order]] = depth[order]
    proj_pc = np.full((H, W, c), -1, dtype=F32)
    proj_pc[py[order], px[order]] = points[order]
    proj_idx = np.full((H, W), -1, dtype=np.int32)
    proj_idx[py[order], px[order]] = idx[order].astype(np.int32)
    proj_mask = (proj_idx > 0).astype(np.int32)  # projection.py:113 (drops point 0)
    return {
        "proj_pointcloud": proj_pc, "proj_range": proj_range, "proj_idx": proj_idx,
        "proj_mask": proj_mask, "uproj_x_idx": px, "uproj_y_idx": py,
        "uproj_depth": depth.copy(),
    }


def pixel_is_boundary_ambiguous(points, fov, depth=None, ulps=2):
    """True for points whose (px, py) changes when yaw / pitch move by up to
    `ulps` float32 ulps -- the only points on which a non correctly rounded
    arctan2/arcsin (numpy SIMD, CUDA libm) may legitimately disagree."""
    points = np.ascontiguousarray(points, dtype=F32)
    depth = depth_of(points) if depth is None else np.asarray(depth, dtype=F32)
    yaw, pitch = angles(points, depth)
    px0, py0 = pixel_coords(yaw, pitch, fov)
    amb = np.zeros(points.shape[0], dtype=bool)
    for s in range(1, ulps + 1):
        for sign in (-1.0, 1.0):
            y2, p2 = yaw.copy(), pitch.copy()
            for _ in range(s):
                y2 = np.nextafter(y2, F32(sign * np.inf))
                p2 = np.nextafter(p2, F32(sign * np.inf))
            px, _ = pixel_coords(y2, pitch, fov)
            _, py = pixel_coords(yaw, p2, fov)
            amb |= (px != px0) | (py != py0)
    return amb
