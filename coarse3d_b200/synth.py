"""Seeded synthetic scans shaped like the datasets COARSE3D trains on.

There is no dataset on the benchmark box, so bench.py and the tests draw
LiDAR-like scans from a beam model.  Shapes follow the
reference configs: SemanticKITTI 64x2048, fov +3/-25 deg
(tasks/weak_segmentation/config_semantic_kitti.yaml:133-141); the nuScenes- and
POSS-shaped variants are BASELINE.json's configs 3 and 4.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ScanShape:
    name: str
    n_points: int
    proj_h: int
    proj_w: int
    fov_up: float
    fov_down: float
    n_classes: int
    label_ratio: float


KITTI = ScanShape("kitti", 120_000, 64, 2048, 3.0, -25.0, 20, 1e-3)
NUSCENES = ScanShape("nuscenes", 34_000, 32, 1024, 15.0, -35.0, 17, 1e-4)
POSS = ScanShape("poss", 72_000, 40, 1800, 15.0, -25.0, 14, 1e-3)
SHAPES = {s.name: s for s in (KITTI, NUSCENES, POSS)}


def make_scan(shape: ScanShape, seed: int, n_points: int = None, sensor_order: bool = False):
    """One scan: points (N,4) f32 [x,y,z,intensity], full labels (N,) i64 in
    [1, C-1], weak labels (N,) i64 (0 = unlabelled).  Points come in random order (BASELINE's
    synthetic spec); sensor_order=True sorts them by (beam, azimuth) the way a
    spinning LiDAR emits them (what the .bin files of the real datasets hold)."""
    rng = np.random.default_rng(seed)
    n = shape.n_points if n_points is None else n_points
    beams = shape.proj_h
    beam = rng.integers(0, beams, n)
    pitch_deg = shape.fov_up - (beam + 0.5) * (shape.fov_up - shape.fov_down) / beams
    pitch = np.deg2rad(pitch_deg + rng.normal(0.0, 0.05, n))
    yaw = rng.uniform(-np.pi, np.pi, n)
    rng_m = np.clip(rng.lognormal(2.3, 0.8, n), 1.0, 80.0)
    xyz = np.stack([rng_m * np.cos(pitch) * np.cos(yaw),
                    rng_m * np.cos(pitch) * np.sin(yaw),
                    rng_m * np.sin(pitch)], 1)
    inten = rng.uniform(0.0, 1.0, n)
    points = np.concatenate([xyz, inten[:, None]], 1).astype(np.float32)
    c1 = shape.n_classes - 1
    sector = np.floor((yaw + np.pi) / (2 * np.pi) * 8).astype(np.int64)
    band = np.minimum((rng_m / 10.0).astype(np.int64), 7)
    full = (sector * 3 + band) % c1 + 1
    weak = np.zeros(n, dtype=np.int64)
    k = max(1, int(round(shape.label_ratio * n)))
    pick = rng.choice(n, k, replace=False)
    weak[pick] = full[pick]
    if sensor_order:
        order = np.lexsort((yaw, beam))
        points, full, weak = points[order], full[order], weak[order]
    return points, full.astype(np.int64), weak


def make_batch(shape: ScanShape, batch: int, seed0: int, ragged: bool = False, sensor_order: bool = False):
    """CSR batch: points (sum N, 4) f32, offsets (B+1,) i32, full / weak labels."""
    pts, fulls, weaks, offs = [], [], [], [0]
    for i in range(batch):
        n = shape.n_points
        if ragged:
            n = int(n * (0.7 + 0.3 * ((seed0 + i) % 7) / 6.0))
        p, f, w = make_scan(shape, seed0 + i, n, sensor_order)
        pts.append(p), fulls.append(f), weaks.append(w)
        offs.append(offs[-1] + n)
    return (np.concatenate(pts, 0), np.asarray(offs, dtype=np.int32),
            np.concatenate(fulls, 0), np.concatenate(weaks, 0))
