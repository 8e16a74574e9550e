"""Oracle: KNN range-image label vote (TEST INFRASTRUCTURE).

Restates `KNN.forward` and `get_gaussian_kernel` from the reference,
pc_processor/postproc/knn.py:11-142, without materialising the unfolds.

Rules fixed where the reference is undefined:

* `torch.topk(k, largest=False, sorted=False)` (knn.py:110-111) picks arbitrary
  members among equal distances.  Rule: the k smallest by (distance, window
  slot), slot = row-major index in the S x S window.  +inf distances compare
  equal to each other and are therefore taken in slot order.
* Vote argmax (knn.py:137) returns the first maximum (torch semantics), i.e.
  the smallest class id among equally voted classes; an all-zero vote gives 1.
"""
import math

import numpy as np
import torch

F32 = np.float32


def gaussian_kernel(kernel_size=3, sigma=2):
    """knn.py:11-33, same torch ops in the same order (bit-identical)."""
    x_coord = torch.arange(kernel_size)
    x_grid = x_coord.repeat(kernel_size).view(kernel_size, kernel_size)
    y_grid = x_grid.t()
    xy_grid = torch.stack([x_grid, y_grid], dim=-1).float()
    mean = (kernel_size - 1) / 2.
    variance = sigma ** 2.
    g = (1. / (2. * math.pi * variance)) * \
        torch.exp(-torch.sum((xy_grid - mean) ** 2., dim=-1) / (2 * variance))
    g = g / torch.sum(g)
    return g.view(kernel_size, kernel_size)


def inv_gauss_weights(search, sigma):
    """knn.py:102-104: (1 - G) flattened row-major, float32."""
    return (1 - gaussian_kernel(search, sigma)).reshape(-1).numpy().astype(F32)


def knn_vote(proj_range, unproj_range, proj_argmax, px, py, knn, search, sigma,
             cutoff, nclasses):
    """KNN.forward (knn.py:54-142) for one scan.  Returns (P,) int64 labels."""
    if search % 2 == 0:
        raise ValueError("Nearest neighbor kernel must be odd number")  # knn.py:72-73
    proj_range = np.asarray(proj_range, dtype=F32)
    unproj_range = np.asarray(unproj_range, dtype=F32)
    proj_argmax = np.asarray(proj_argmax)
    px = np.asarray(px).astype(np.int64)
    py = np.asarray(py).astype(np.int64)
    H, W = proj_range.shape
    P = unproj_range.shape[0]
    S = search
    pad = (S - 1) // 2
    S2 = S * S

    # knn.py:79-85,114-117: zero-padded unfold, gathered at (py, px)
    rng_p = np.zeros((H + 2 * pad, W + 2 * pad), dtype=F32)
    rng_p[pad:pad + H, pad:pad + W] = proj_range
    cls_p = np.zeros((H + 2 * pad, W + 2 * pad), dtype=np.int64)
    cls_p[pad:pad + H, pad:pad + W] = proj_argmax
    win = np.empty((P, S2), dtype=F32)
    cls = np.empty((P, S2), dtype=np.int64)
    for dy in range(S):
        for dx in range(S):
            win[:, dy * S + dx] = rng_p[py + dy, px + dx]
            cls[:, dy * S + dx] = cls_p[py + dy, px + dx]

    win[win < 0] = np.inf                       # knn.py:90
    center = (S2 - 1) // 2
    win[:, center] = unproj_range               # knn.py:93-94
    with np.errstate(invalid="ignore"):
        d = np.abs(win - unproj_range[:, None])     # knn.py:97
        d = d * inv_gauss_weights(S, sigma)[None, :]  # knn.py:107

    # knn.py:110-111 with the tie rule: stable sort == (distance, slot) order
    sel = np.argsort(d, axis=1, kind="stable")[:, :knn]
    sel_cls = np.take_along_axis(cls, sel, axis=1)           # knn.py:120-121
    if cutoff > 0:                                           # knn.py:124-127
        sel_d = np.take_along_axis(d, sel, axis=1)
        sel_cls = np.where(sel_d > F32(cutoff), nclasses, sel_cls)

    votes = np.zeros((P, nclasses + 1), dtype=np.int32)      # knn.py:131-134
    np.add.at(votes, (np.repeat(np.arange(P), knn), sel_cls.reshape(-1)), 1)
    return votes[:, 1:-1].argmax(axis=1).astype(np.int64) + 1  # knn.py:137
