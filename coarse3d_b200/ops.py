"""Tensor-level batched API over the C ABI (device tensors in, device tensors out).

These are the throughput entry points: a CSR batch of scans per call, all work
enqueued on torch's current CUDA stream, no host synchronisation.  The
reference-shaped classes in `coarse3d_b200.pc_processor` are thin adapters over
these functions.
"""
import ctypes
import math
from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import check, lib


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(**tensors):
    for name, t in tensors.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("coarse3d_b200: `%s` must be a CUDA tensor (no CPU path)" % name)
        if not t.is_contiguous():
            raise ValueError("coarse3d_b200: `%s` must be contiguous" % name)


def _cofill_args(cofill):
    """(pointer, bytes) of a carried fill: a contiguous CUDA tensor (or a flat slice of one)
    that the call's kernels zero on the side while they do their own work."""
    if cofill is None or cofill.numel() == 0:
        return None, 0
    _need_cuda(cofill=cofill)
    n = cofill.numel() * cofill.element_size()
    if n % 16 or cofill.data_ptr() % 16:
        raise ValueError("cofill must be contiguous, 16 B aligned and a multiple of 16 B")
    return ctypes.c_void_p(cofill.data_ptr()), n


# --------------------------------------------------------------------- a1 --
class Fov(NamedTuple):
    """The angles RangeProjection.__init__ stores (projection.py:29-35), radians."""
    abs_fov_left: float
    fov_hori: float
    abs_fov_down: float
    fov_vert: float

    @staticmethod
    def from_degrees(fov_up=3, fov_down=-25, fov_left=-180, fov_right=180):
        up = fov_up / 180.0 * math.pi
        down = fov_down / 180.0 * math.pi
        left = fov_left / 180.0 * math.pi
        right = fov_right / 180.0 * math.pi
        return Fov(abs(left), abs(left) + abs(right), abs(down), abs(up) + abs(down))


class Projection(NamedTuple):
    proj_pointcloud: torch.Tensor  # (B,H,W,C) f32
    proj_range: torch.Tensor       # (B,H,W) f32
    proj_idx: torch.Tensor         # (B,H,W) i32, index within the scan
    proj_mask: torch.Tensor        # (B,H,W) i32
    uproj_x_idx: torch.Tensor      # (sum N,) i32
    uproj_y_idx: torch.Tensor      # (sum N,) i32
    uproj_depth: torch.Tensor      # (sum N,) f32
    flags: torch.Tensor            # (1,) i32; bit 0: NaN pixel coordinate


class ProjectionBuffers:
    """Pre-allocated outputs + z-buffer scratch, reusable across calls (and
    stable addresses for CUDA-graph capture)."""

    def __init__(self, batch, total_points, c_in, proj_h, proj_w, device):
        f32, i32 = torch.float32, torch.int32
        self.batch, self.total, self.c_in, self.H, self.W = batch, total_points, c_in, proj_h, proj_w
        self.proj_pointcloud = torch.empty((batch, proj_h, proj_w, c_in), dtype=f32, device=device)
        self.proj_range = torch.empty((batch, proj_h, proj_w), dtype=f32, device=device)
        self.proj_idx = torch.empty((batch, proj_h, proj_w), dtype=i32, device=device)
        self.proj_mask = torch.empty((batch, proj_h, proj_w), dtype=i32, device=device)
        self.uproj_x_idx = torch.empty((total_points,), dtype=i32, device=device)
        self.uproj_y_idx = torch.empty((total_points,), dtype=i32, device=device)
        self.uproj_depth = torch.empty((total_points,), dtype=f32, device=device)
        self.flags = torch.zeros((1,), dtype=i32, device=device)   # sticky: bit 0 stays set once a
        # call saw a NaN pixel coordinate (depth == 0); `flags.zero_()` re-arms it
        nbytes = lib.c3d_project_workspace_bytes(batch, proj_h, proj_w)
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        self.clean = False  # "two" / "fused": the form whose call has left the z-buffer reset


PROJECT_CLUSTER_DEFAULT = False   # set by the measurements in DESIGN.md 3.1


def project_batch(points, offsets, fov: Fov, proj_h, proj_w, depth=None,
                  buffers: ProjectionBuffers = None, exact_f64=False, fused_kernel=False,
                  cofill=None, cluster_kernel=None) -> Projection:
    """RangeProjection.doProjection for a CSR batch (projection.py:43-115).

    points (sum N, C>=3) f32, offsets (B+1,) i32, optional depth (sum N,) f32;
    all CUDA.  proj_idx holds indices local to each scan.  exact_f64=True evaluates the angles
    of EVERY point in fp64 (the default does so only inside the guard band of a pixel boundary;
    both give the same pixels, the flag exists to prove that).  cluster_kernel: the form that keeps
    each scan's z-buffer in the distributed shared memory of a thread-block cluster (None = where
    it applies: C == 4 and H*W*8 <= 8 x 200 KB; bit-identical results).
    """
    _need_cuda(points=points, offsets=offsets, depth=depth)
    if points.dtype != torch.float32 or points.dim() != 2:
        raise ValueError("points must be (N, C) float32")
    if offsets.dtype != torch.int32:
        raise ValueError("offsets must be int32")
    if depth is not None and (depth.dtype != torch.float32 or depth.numel() != points.shape[0]):
        raise ValueError("depth must be (N,) float32")
    batch = offsets.numel() - 1
    total, c_in = points.shape
    b = buffers
    if b is None:
        b = ProjectionBuffers(batch, total, c_in, proj_h, proj_w, points.device)
    elif (b.batch, b.total, b.c_in, b.H, b.W) != (batch, total, c_in, proj_h, proj_w):
        raise ValueError("ProjectionBuffers shape mismatch")
    want_cluster = PROJECT_CLUSTER_DEFAULT if cluster_kernel is None else bool(cluster_kernel)
    cluster = (want_cluster and not fused_kernel and total > 0 and points.data_ptr() % 16 == 0 and
               bool(lib.c3d_project_cluster_supported(c_in, proj_h, proj_w)))
    form = "fused" if fused_kernel else "two"
    keep_clean = b.clean                            # the cluster form does not touch the workspace
    was_clean, b.clean = (b.clean == form), False   # a failed call may leave a dirty z-buffer; the two
    # forms initialise different parts of the workspace, so "clean" holds per form
    check(lib.c3d_project_batch(
        _p(points), c_in, _p(offsets), batch, total, _p(depth),
        fov.abs_fov_left, fov.fov_hori, fov.abs_fov_down, fov.fov_vert, proj_h, proj_w,
        _p(b.proj_range), _p(b.proj_pointcloud), _p(b.proj_idx), _p(b.proj_mask),
        _p(b.uproj_x_idx), _p(b.uproj_y_idx), _p(b.uproj_depth), _p(b.workspace),
        (1 if was_clean else 0) | (2 if exact_f64 else 0) | (4 if fused_kernel else 0) | (8 if cluster else 0),
        _p(b.flags), *_cofill_args(cofill), _stream()))
    b.clean = keep_clean if cluster else form
    return Projection(b.proj_pointcloud, b.proj_range, b.proj_idx, b.proj_mask,
                      b.uproj_x_idx, b.uproj_y_idx, b.uproj_depth, b.flags)


# --------------------------------------------------------------------- f1 --
class Assembled(NamedTuple):
    feature: torch.Tensor      # (B,5,H,W) f32 [range,x,y,z,intensity], normalised if mean/std given
    train_label: Optional[torch.Tensor]  # (B,H,W) i64
    eval_label: Optional[torch.Tensor]   # (B,H,W) i64
    proj_range: torch.Tensor   # (B,H,W) f32
    proj_idx: torch.Tensor     # (B,H,W) i32
    uproj_x_idx: torch.Tensor  # (sum N,) i32
    uproj_y_idx: torch.Tensor
    uproj_depth: torch.Tensor
    flags: torch.Tensor


def project_assemble_batch(points, offsets, fov: Fov, proj_h, proj_w, sem_label=None,
                           weak_label=None, img_mean=None, img_std=None, depth=None,
                           buffers: ProjectionBuffers = None, fused_kernel=False, cofill=None) -> Assembled:
    """Projection fused with its caller (loader :124-172, trainer :600-608): label images
    and the 5-channel network input straight from the z-buffer winners.

    points (sum N, 4) f32; sem_label / weak_label (sum N,) int32 (the loaders' dtype) or uint8
    (both the same); img_mean / img_std (5,) f32.
    """
    _need_cuda(points=points, offsets=offsets, depth=depth, sem_label=sem_label,
               weak_label=weak_label, img_mean=img_mean, img_std=img_std)
    if points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] != 4:
        raise ValueError("points must be (N, 4) float32")
    ldt = None
    for name, t in (("sem_label", sem_label), ("weak_label", weak_label)):
        if t is None:
            continue
        if t.dtype not in (torch.int32, torch.uint8) or t.numel() != points.shape[0] or \
                (ldt is not None and t.dtype != ldt):
            raise ValueError("%s must be (N,) int32 or uint8 (both labels the same dtype)" % name)
        ldt = t.dtype
    batch, total = offsets.numel() - 1, points.shape[0]
    b = buffers
    if b is None:
        b = ProjectionBuffers(batch, total, 4, proj_h, proj_w, points.device)
    elif (b.batch, b.total, b.c_in, b.H, b.W) != (batch, total, 4, proj_h, proj_w):
        raise ValueError("ProjectionBuffers shape mismatch")
    dev = points.device
    feature = torch.empty((batch, 5, proj_h, proj_w), dtype=torch.float32, device=dev)
    train = torch.empty((batch, proj_h, proj_w), dtype=torch.int64, device=dev) if weak_label is not None else None
    evall = torch.empty((batch, proj_h, proj_w), dtype=torch.int64, device=dev) if sem_label is not None else None
    form = "fused" if fused_kernel else "two"
    was_clean, b.clean = (b.clean == form), False
    check(lib.c3d_project_assemble_batch(
        _p(points), _p(offsets), batch, total, _p(depth), _p(sem_label), _p(weak_label),
        1 if ldt == torch.uint8 else 0, _p(img_mean), _p(img_std), fov.abs_fov_left, fov.fov_hori, fov.abs_fov_down, fov.fov_vert,
        proj_h, proj_w, _p(feature), _p(train), _p(evall), _p(b.proj_range), _p(b.proj_idx),
        _p(b.uproj_x_idx), _p(b.uproj_y_idx), _p(b.uproj_depth), _p(b.workspace),
        (1 if was_clean else 0) | (4 if fused_kernel else 0), _p(b.flags),
        *_cofill_args(cofill), _stream()))
    b.clean = form
    return Assembled(feature, train, evall, b.proj_range, b.proj_idx, b.uproj_x_idx, b.uproj_y_idx,
                     b.uproj_depth, b.flags)


# --------------------------------------------------------------------- a4 --
def gaussian_kernel(kernel_size=3, sigma=2):
    """get_gaussian_kernel (knn.py:11-33): same torch ops, CPU, float32."""
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


def knn_batch(proj_range, proj_argmax, unproj_range, px, py, offsets, knn, search, sigma,
              cutoff, nclasses, inv_gauss=None, out=None, cofill=None, out_uint8=False, records=None):
    """KNN.forward for a CSR batch (knn.py:54-142).

    proj_range (B,H,W) f32; px, py (sum N,) int64 (the reference's dtype) or
    int32 (project_batch's output); proj_argmax (B,H,W) int64 or int32; offsets
    (B+1,) i32.  Returns (sum N,) labels with proj_argmax's dtype, or uint8 with out_uint8=True
    (opt-in: class ids < 256, an eighth of the reference's int64 bytes on the way to the host).
    `cofill`: a contiguous CUDA tensor zeroed by the same kernel (TMA bulk stores from a
    shared-memory zero page while the vote keeps the ALU busy; the step pipeline passes the
    loss's dense gradient buffer).
    `records`: the binned per-point records of `knn_sort_points` (same points, same results): the
    vote then reads them instead of unproj_range / px / py, and a warp's gathers share cache lines.
    """
    if search % 2 == 0:
        raise ValueError("Nearest neighbor kernel must be odd number")  # knn.py:72-73
    _need_cuda(proj_range=proj_range, proj_argmax=proj_argmax, unproj_range=unproj_range,
               px=px, py=py, offsets=offsets)
    idt, ldt = px.dtype, proj_argmax.dtype
    if idt not in (torch.int64, torch.int32) or py.dtype != idt:
        raise ValueError("px, py must share dtype int64 or int32")
    if ldt not in (torch.int64, torch.int32):
        raise ValueError("proj_argmax must be int64 or int32")
    if proj_range.dtype != torch.float32 or unproj_range.dtype != torch.float32:
        raise ValueError("ranges must be float32")
    if proj_range.dim() != 3 or proj_argmax.shape != proj_range.shape:
        raise ValueError("proj_range / proj_argmax must be (B,H,W)")
    B, H, W = proj_range.shape
    if offsets.numel() != B + 1 or offsets.dtype != torch.int32:
        raise ValueError("offsets must be (B+1,) int32")
    total = unproj_range.numel()
    if inv_gauss is None:
        inv_gauss = (1 - gaussian_kernel(search, sigma)).reshape(-1).to(proj_range.device)
    if out is None:
        out = torch.empty((total,), dtype=torch.uint8 if out_uint8 else ldt, device=proj_range.device)
    elif out.dtype != (torch.uint8 if out_uint8 else ldt) or \
            (out.numel() != total and not (records is not None and out.numel() > total)):
        # with records the labels go to the ORIGINAL point indices, which may belong to a larger
        # batch than the scans of this call (two-launch vote): `out` is then the whole batch's
        raise ValueError("out must be (sum N,) %s" % (torch.uint8 if out_uint8 else ldt))
    nfill = 0
    if cofill is not None:
        _need_cuda(cofill=cofill)
        nfill = cofill.numel() * cofill.element_size()
        if not cofill.is_contiguous() or nfill % 16 or cofill.data_ptr() % 16:
            raise ValueError("cofill must be contiguous, 16 B aligned and a multiple of 16 B")
    if records is not None:
        _need_cuda(records=records)
        if records.dtype != torch.float32 or records.shape != (total, 4):
            raise ValueError("records must be the (sum N, 4) float32 tensor of knn_sort_points")
    check(lib.c3d_knn_batch(
        _p(proj_range), _p(proj_argmax), _p(unproj_range), _p(px if records is None else records), _p(py),
        _p(offsets), B, total,
        H, W, int(knn), int(search), float(cutoff), int(nclasses), _p(inv_gauss),
        2 if records is not None else (1 if idt == torch.int64 else 0),
        (1 if ldt == torch.int64 else 0) | (2 if out_uint8 else 0), _p(out),
        _p(cofill) if nfill else None, nfill, _stream()))
    return out


def knn_sort_workspace(batch, total, proj_h, proj_w, device):
    n = lib.c3d_knn_sort_workspace_bytes(int(batch), int(total), int(proj_h), int(proj_w))
    return torch.empty((max(n, 256),), dtype=torch.uint8, device=device)


def knn_sort_points(unproj_range, px, py, offsets, proj_h, proj_w, workspace=None, out=None):
    """Bin the points of every scan by (row, 32-pixel column segment) for `knn_batch(records=)`:
    returns (sum N, 4) float32 records {range, x, y, original index} (ints bit-cast), scan-major,
    so that the vote's warps touch shared cache lines.  Three small kernels, no sync."""
    _need_cuda(unproj_range=unproj_range, px=px, py=py, offsets=offsets)
    if px.dtype not in (torch.int64, torch.int32) or py.dtype != px.dtype:
        raise ValueError("px, py must share dtype int64 or int32")
    if unproj_range.dtype != torch.float32 or offsets.dtype != torch.int32:
        raise ValueError("unproj_range must be float32, offsets int32")
    total, batch = unproj_range.numel(), offsets.numel() - 1
    if workspace is None:
        workspace = knn_sort_workspace(batch, total, proj_h, proj_w, px.device)
    if out is None:
        out = torch.empty((total, 4), dtype=torch.float32, device=px.device)
    check(lib.c3d_knn_sort_points(_p(unproj_range), _p(px), _p(py), _p(offsets), batch, total, int(proj_h),
                                  int(proj_w), 1 if px.dtype == torch.int64 else 0, _p(workspace), _p(out),
                                  _stream()))
    return out


# --------------------------------------------------------------------- f2 --
def unproject_confusion_batch(proj_argmax, px, py, offsets, nclasses, labels=None, conf_matrix=None,
                              want_unproj=True, flags=None):
    """Per-point prediction `argmax_2d[b, py, px]` (trainer.py:714-724) and, with `labels`,
    IOUEval.addBatch (iou_eval.py:35-58) accumulated into `conf_matrix` (C,C) int64
    (rows = prediction, columns = ground truth).  Returns (unproj_argmax or None, conf_matrix)."""
    _need_cuda(proj_argmax=proj_argmax, px=px, py=py, offsets=offsets, labels=labels,
               conf_matrix=conf_matrix)
    def is64(t, name):
        if t.dtype not in (torch.int64, torch.int32):
            raise ValueError("%s must be int64 or int32" % name)
        return 1 if t.dtype == torch.int64 else 0
    if py.dtype != px.dtype:
        raise ValueError("px and py must share a dtype")
    B, H, W = proj_argmax.shape
    total = px.numel()
    dev = proj_argmax.device
    if labels is not None and conf_matrix is None:
        conf_matrix = torch.zeros((nclasses, nclasses), dtype=torch.int64, device=dev)
    if conf_matrix is not None and (conf_matrix.dtype != torch.int64 or conf_matrix.shape != (nclasses, nclasses)):
        raise ValueError("conf_matrix must be (C, C) int64")
    out = torch.empty((total,), dtype=proj_argmax.dtype, device=dev) if want_unproj else None
    if flags is None:
        flags = torch.zeros((1,), dtype=torch.int32, device=dev)
    check(lib.c3d_unproject_confusion_batch(
        _p(proj_argmax), _p(px), _p(py), _p(labels), _p(offsets), B, total, H, W, int(nclasses),
        is64(proj_argmax, "proj_argmax"), is64(px, "px"), 0 if labels is None else is64(labels, "labels"),
        _p(out), _p(conf_matrix if labels is not None else None), _p(flags), _stream()))
    return out, conf_matrix


# --------------------------------------------------------------------- f3 --
def entropy_select_batch(output, wss_mask, eval_mask, train_label, select_ratio, ignore_cls=0,
                         noise=None, seed=None, workspace=None):
    """Trainer.entropy_based_selection (trainer.py:447-518) for the whole batch in three
    launches.  output (B,C,H,W) f32 probs; wss_mask / eval_mask (B,H,W) bool; train_label
    (B,H,W) int64.  noise (B,C,H*W) f32 injects the Exp(1) draws of the reference's
    multinomial calls; otherwise Philox draws from `seed`.
    Returns (pseudo_label (B,H,W) int64, new_wss_mask (B,H,W) bool)."""
    _need_cuda(output=output, wss_mask=wss_mask, eval_mask=eval_mask, train_label=train_label, noise=noise)
    if output.dtype != torch.float32 or train_label.dtype != torch.int64:
        raise ValueError("output must be float32 and train_label int64")
    if wss_mask.dtype != torch.bool or eval_mask.dtype != torch.bool:
        raise ValueError("wss_mask / eval_mask must be bool")
    B, C, H, W = output.shape
    if noise is not None and (noise.dtype != torch.float32 or noise.shape != (B, C, H * W)):
        raise ValueError("noise must be (B, C, H*W) float32")
    if workspace is None:
        n = lib.c3d_entropy_select_workspace_bytes(B, C, H * W)
        workspace = torch.empty((n,), dtype=torch.uint8, device=output.device)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
    label = torch.empty((B, H, W), dtype=torch.int64, device=output.device)
    mask = torch.empty((B, H, W), dtype=torch.bool, device=output.device)
    check(lib.c3d_entropy_select_batch(
        _p(output), _p(train_label), _p(wss_mask), _p(eval_mask), B, C, H, W, int(ignore_cls),
        float(select_ratio), _p(noise), int(seed), _p(workspace), _p(label), _p(mask), _stream()))
    return label, mask


# --------------------------------------------------------------------- f4 --
LOVASZ_MAX_VALID = 1 << 24      # radix path (24 pixel bits in the sort keys); <= 32768 stays in shared memory / all-pairs


def lovasz_info(workspace):
    """(valid pixels, classes averaged, flags) of the last Lovasz forward.  Synchronises."""
    host = (ctypes.c_int32 * 4)()
    check(lib.c3d_lovasz_info(_p(workspace), ctypes.cast(host, ctypes.c_void_p), _stream()))
    return int(host[0]), int(host[1]), int(host[2])


class _LovaszFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, labels, ignore, classes_all, class_mask, max_valid, workspace):
        B, C, H, W = probs.shape
        loss = torch.empty((), dtype=torch.float32, device=probs.device)
        check(lib.c3d_lovasz_forward(_p(probs), _p(labels), B, C, H, W, int(ignore), int(classes_all),
                                     int(class_mask), int(max_valid), _p(workspace), _p(loss), _stream()))
        ctx.args = (B, C, H, W, int(classes_all), int(class_mask), int(max_valid), workspace)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        B, C, H, W, classes_all, class_mask, max_valid, workspace = ctx.args
        grad = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
        check(lib.c3d_lovasz_backward(B, C, H, W, classes_all, class_mask, max_valid, _p(workspace),
                                      _p(grad_out.contiguous().float()), _p(grad), 0, _stream()))
        return grad, None, None, None, None, None, None


def lovasz_softmax(probs, labels, ignore=None, classes="present", max_valid=None,
                   workspace=None):
    """lovasz_softmax(per_image=False) of lovasz_softmax.py:67-98 on (B,C,H,W) probabilities and
    (B,H,W) int64 labels; `classes`: 'present', 'all' or a list of class ids (:117).  Returns
    (0-dim loss with autograd to `probs`, workspace).
    `max_valid` = capacity in valid (labelled) pixels; it selects the sort path (<= 16384: one CTA
    per class in shared memory; <= 32768: all-pairs ranks; more, up to 2^24: device-wide radix
    sort) and sizes the workspace.  None: the valid pixels are counted first (one host read, as
    the reference's own `.sum() == 0` checks do) and the next power of two is taken.  More valid
    pixels than an explicit `max_valid`: NaN loss + flag (`lovasz_info`)."""
    _need_cuda(probs=probs, labels=labels)
    if probs.dtype != torch.float32 or probs.dim() != 4:
        raise ValueError("probas must be (B, C, H, W) float32")
    if labels.dtype != torch.int64 or labels.shape != (probs.shape[0],) + probs.shape[2:]:
        raise ValueError("labels must be (B, H, W) int64")
    C = probs.shape[1]
    class_mask = 0
    if isinstance(classes, (list, tuple)):          # an explicit class list (lovasz_softmax.py:117)
        for c in classes:
            if not 0 <= int(c) < C:
                raise ValueError("class %r outside [0, %d)" % (c, C))
            class_mask |= 1 << int(c)
        mode = 2
    elif classes in ("present", "all"):
        mode = 1 if classes == "all" else 0
    else:
        raise ValueError("classes must be 'present', 'all' or a list of class ids")
    if max_valid is None:
        lab_ok = (labels >= 0) & (labels < C)
        if ignore is not None:
            lab_ok &= labels != int(ignore)
        n_valid = int(lab_ok.sum())
        max_valid = 1024
        while max_valid < min(n_valid, 32768):
            max_valid *= 2                       # shared-memory / all-pairs paths: power-of-two capacities
        if n_valid > 32768:                      # radix path: the sort handles C * capacity keys
            max_valid = (n_valid + 4095) // 4096 * 4096
        max_valid = min(max_valid, max(labels.numel(), 1))
        workspace = None
    if workspace is None:
        n = lib.c3d_lovasz_workspace_bytes(C, int(max_valid))
        if n == 0:
            raise ValueError("Lovasz: max_valid must be in [1, %d] and n_classes <= 64" % LOVASZ_MAX_VALID)
        workspace = torch.empty((n,), dtype=torch.uint8, device=probs.device)
    loss = _LovaszFn.apply(probs.contiguous(), labels.contiguous(), -1 if ignore is None else int(ignore),
                           mode, class_mask, int(max_valid), workspace)
    return loss, workspace


# --------------------------------------------------------------------- a2 --
class ProtoLossConfig(NamedTuple):
    ignore_label: int = 0
    temperature: float = 0.1
    base_temperature: float = 0.07
    num_anchor: int = 50


FLAG_NO_ANCHOR, FLAG_BAD_KEEP, FLAG_KEEP_ROWS, FLAG_BAD_LABEL = 1, 2, 4, 8


def proto_loss_workspace(batch, n_classes, hw, dim, sub_protos, num_anchor, device):
    n = lib.c3d_proto_loss_workspace_bytes(batch, n_classes, hw, dim, sub_protos, num_anchor)
    if n == 0:
        raise ValueError("bad prototype-loss shape")
    return torch.empty((n,), dtype=torch.uint8, device=device)


def proto_loss_info(workspace):
    """(T segments, labelled pixels, flags) of the last forward.  Synchronises."""
    host = (ctypes.c_int32 * 4)()
    check(lib.c3d_proto_loss_info(_p(workspace), ctypes.cast(host, ctypes.c_void_p), _stream()))
    return int(host[0]), int(host[1]), int(host[2])


def proto_loss_rows(workspace, batch, dim, hw, n_classes, sub_protos, num_anchor):
    """Labelled-pixel slots of the last forward: (pix, cls, cnt) int32 tensors,
    sorted by (class, scan, pixel).  Synchronises (reads the slot count)."""
    _, n_lab, _ = proto_loss_info(workspace)
    dev = workspace.device
    pix, cls, cnt = (torch.empty((max(n_lab, 1),), dtype=torch.int32, device=dev) for _ in range(3))
    if n_lab:
        check(lib.c3d_proto_loss_rows(_p(workspace), batch, dim, hw, n_classes, sub_protos, num_anchor, n_lab,
                                      _p(pix), _p(cls), _p(cnt), _stream()))
    return pix[:n_lab], cls[:n_lab], cnt[:n_lab]


def proto_loss_forward_raw(feats, probs, labels, keep_mask, queue, cfg, keep, seed, workspace, loss_out,
                           need_grad=True, phases=3, tensor_cores=False):
    """c3d_proto_loss_forward on pre-validated device tensors (no autograd, no allocation).
    phases: 1 = selection only, 2 = rows only (after a phase 1), 3 = both."""
    B, D, H, W = feats.shape
    C, M, _ = queue.shape
    check(lib.c3d_proto_loss_forward_phase(
        _p(feats), _p(probs), _p(labels), _p(keep_mask), _p(queue), B, D, H, W, C, M,
        int(cfg.ignore_label), float(cfg.temperature), float(cfg.base_temperature),
        int(cfg.num_anchor), _p(keep), 0 if keep is None else keep.shape[0], int(seed),
        (1 if need_grad else 0) | (2 if tensor_cores else 0), int(phases), _p(workspace), _p(loss_out),
        _stream()))
    return loss_out


def zero_fill(t):
    """Zero a contiguous float32 CUDA tensor with the library's streaming fill kernel."""
    _need_cuda(t=t)
    check(lib.c3d_zero_fill(_p(t), t.numel() * t.element_size(), _stream()))
    return t


def zero_fill_background(t, mode=0, ctas_per_sm=1, page_bytes=8192, inflight=4):
    """Zero a contiguous CUDA tensor with the minimal-footprint persistent fill kernel
    (c3d_zero_fill_background): meant to be launched first, on its own stream, and to run under
    the other kernels of a step."""
    _need_cuda(t=t)
    check(lib.c3d_zero_fill_background(_p(t), t.numel() * t.element_size(), int(mode), int(ctas_per_sm),
                                       int(page_bytes), int(inflight), _stream()))
    return t


def zero_fill_daemon(t, ctrl, max_per_sm=1, launch_per_sm=4, page_bytes=8192, chunk_pages=4, debug=None):
    """c3d_zero_fill_daemon: the placement-proof background fill.  ctrl: >= 1 KB uint8/int32
    CUDA scratch; debug: optional int64 (launch_per_sm*148, 4) tensor."""
    _need_cuda(t=t, ctrl=ctrl, debug=debug)
    check(lib.c3d_zero_fill_daemon(_p(t), t.numel() * t.element_size(), int(max_per_sm), int(launch_per_sm),
                                   int(page_bytes), int(chunk_pages), _p(ctrl), _p(debug), _stream()))
    return t


def delay(ns):
    check(lib.c3d_delay(int(ns), _stream()))


def proto_loss_backward_raw(shape, cfg, n_classes, sub_protos, workspace, grad_out, grad_feats,
                            grad_is_zeroed=False):
    """c3d_proto_loss_backward: writes the dense (B,D,H,W) gradient into grad_feats.
    grad_is_zeroed=True skips the zero fill (the caller ran `zero_fill(grad_feats)`)."""
    B, D, H, W = shape
    check(lib.c3d_proto_loss_backward(
        B, D, H, W, n_classes, sub_protos, int(cfg.num_anchor), _p(workspace), _p(grad_out),
        _p(grad_feats), 1 if grad_is_zeroed else 0, _stream()))
    return grad_feats


class _ProtoLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, probs, labels, keep_mask, queue, cfg, keep, seed, workspace):
        loss = torch.empty((), dtype=torch.float32, device=feats.device)
        need_grad = feats.requires_grad
        proto_loss_forward_raw(feats, probs, labels, keep_mask, queue, cfg, keep, seed, workspace, loss,
                               need_grad=need_grad)
        ctx.workspace, ctx.cfg, ctx.cm, ctx.shape = workspace, cfg, (queue.shape[0], queue.shape[1]), feats.shape
        ctx.prefill = None
        if need_grad and PREFILL_GRAD:
            # start the dense zero fill now, on a side stream, so that it overlaps
            # whatever runs between this forward and the backward pass
            grad = torch.empty_like(feats)
            side = _side_stream(feats.device)
            side.wait_stream(torch.cuda.current_stream(feats.device))
            grad.record_stream(side)
            with torch.cuda.stream(side):
                zero_fill(grad)
                ev = torch.cuda.Event()
                ev.record(side)
            ctx.prefill = (grad, ev)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        C, M = ctx.cm
        grad_out = grad_out.contiguous().float()
        if ctx.prefill is not None:
            grad, ev = ctx.prefill
            ctx.prefill = None   # sole owner: autograd can adopt the 0.5 GB buffer instead of cloning it
            torch.cuda.current_stream(grad.device).wait_event(ev)
            proto_loss_backward_raw(ctx.shape, ctx.cfg, C, M, ctx.workspace, grad_out, grad,
                                    grad_is_zeroed=True)
        else:
            grad = torch.empty(ctx.shape, dtype=torch.float32, device=grad_out.device)
            proto_loss_backward_raw(ctx.shape, ctx.cfg, C, M, ctx.workspace, grad_out, grad)
        return grad, None, None, None, None, None, None, None, None


PREFILL_GRAD = True
_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


def proto_loss(feats, probs, labels, keep_mask, proto_queue, cfg: ProtoLossConfig, keep=None,
               seed=None, workspace=None):
    """ContrastMEMLoss.forward (contrast_pixel_loss.py:27-75) on device tensors.

    feats (B,D,H,W) f32 [grad], probs (B,C,H,W) f32, labels (B,H,W) i64,
    keep_mask (B,H,W) bool or None, proto_queue (C,M,D) f32.  `keep` (T,A) i64
    injects the sampled anchors; otherwise they are drawn on the device from
    `seed` (default: drawn from torch's global CPU generator).
    Returns (loss 0-dim tensor, workspace).
    """
    _need_cuda(feats=feats, probs=probs, labels=labels, keep_mask=keep_mask,
               proto_queue=proto_queue, keep=keep)
    if feats.dtype != torch.float32 or probs.dtype != torch.float32 or proto_queue.dtype != torch.float32:
        raise ValueError("feats / probs / proto_queue must be float32")
    if labels.dtype != torch.int64:
        raise ValueError("labels must be int64")
    if keep_mask is not None and keep_mask.dtype not in (torch.bool, torch.uint8):
        raise ValueError("keep_mask must be bool")
    if keep is not None and (keep.dtype != torch.int64 or keep.dim() != 2
                             or keep.shape[1] != cfg.num_anchor):
        raise ValueError("keep must be (T, num_anchor) int64")
    B, D, H, W = feats.shape
    C, M, D2 = proto_queue.shape
    if D2 != D or probs.shape != (B, C, H, W) or labels.shape != (B, H, W):
        raise ValueError("shape mismatch between feats / probs / labels / proto_queue")
    if workspace is None:
        workspace = proto_loss_workspace(B, C, H * W, D, M, cfg.num_anchor, feats.device)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    loss = _ProtoLossFn.apply(feats, probs, labels, keep_mask, proto_queue, cfg, keep, seed, workspace)
    return loss, workspace


# --------------------------------------------------------------------- a3 --
EMA_FLAG_NO_ROWS, EMA_FLAG_BAD_LABEL, EMA_FLAG_OVERFLOW = 1, 8, 16
ASSIGN_ARGMAX, ASSIGN_GUMBEL_INJECTED, ASSIGN_GUMBEL_DEVICE = 0, 1, 2


class EmaAccum(NamedTuple):
    packed: torch.Tensor        # (C*M*D + C*M,) f32: feature sums then counts
    proto_target: Optional[torch.Tensor]  # (B*H*W,) f32 or None
    workspace: torch.Tensor


def proto_ema_info(workspace):
    """(non-empty segments, labelled rows, flags) of the last accumulate.  Synchronises."""
    host = (ctypes.c_int32 * 4)()
    check(lib.c3d_proto_ema_info(_p(workspace), ctypes.cast(host, ctypes.c_void_p), _stream()))
    return int(host[0]), int(host[1]), int(host[2])


def proto_ema_accumulate(embedding, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                         ignore_label=0, ln_eps=1e-5, gumbel=None, assign_mode=None, seed=None,
                         max_rows=None, want_target=False, workspace=None, packed=None) -> EmaAccum:
    """Per-(class, sub-prototype) feature sums and counts of this rank's scans
    (salsanext_proto.py:497-510 + :340-377), as the all-reduce payload."""
    _need_cuda(embedding=embedding, label=label, prototypes=prototypes, ln_d_w=ln_d_w,
               ln_d_b=ln_d_b, ln_c_w=ln_c_w, ln_c_b=ln_c_b, gumbel=gumbel)
    if embedding.dtype != torch.float32 or prototypes.dtype != torch.float32:
        raise ValueError("embedding / prototypes must be float32")
    if label.dtype != torch.int64:
        raise ValueError("label must be int64")
    B, D, H, W = embedding.shape
    C, M, D2 = prototypes.shape
    if D2 != D or label.numel() != B * H * W:
        raise ValueError("shape mismatch between embedding / label / prototypes")
    if assign_mode is None:
        assign_mode = ASSIGN_GUMBEL_INJECTED if gumbel is not None else ASSIGN_GUMBEL_DEVICE
    if max_rows is None:
        max_rows = min(B * H * W, 1 << 17)
    if gumbel is not None and (gumbel.dtype != torch.float32 or gumbel.dim() != 2 or gumbel.shape[1] != M):
        raise ValueError("gumbel must be (rows, M) float32 in (class, pixel) row order")
    if workspace is None:
        n = lib.c3d_proto_ema_workspace_bytes(B, C, H * W, D, M, max_rows)
        if n == 0:
            raise ValueError("bad EMA shape")
        workspace = torch.empty((n,), dtype=torch.uint8, device=embedding.device)
    if packed is None:
        packed = torch.empty((C * M * D + C * M,), dtype=torch.float32, device=embedding.device)
    target = torch.empty((B * H * W,), dtype=torch.float32, device=embedding.device) if want_target else None
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if assign_mode == ASSIGN_GUMBEL_DEVICE else 0
    check(lib.c3d_proto_ema_accumulate(
        _p(embedding), _p(label), _p(prototypes), _p(ln_d_w), _p(ln_d_b), _p(ln_c_w), _p(ln_c_b),
        float(ln_eps), B, D, H, W, C, M, int(ignore_label), int(max_rows), _p(gumbel),
        int(assign_mode), int(seed), _p(workspace), _p(packed), _p(target), _stream()))
    return EmaAccum(packed, target, workspace)


def proto_ema_accumulate_dense(out_feat, nearest, label, feat_proto_sim, ignore_label=0, gumbel=None,
                               assign_mode=None, seed=None, max_rows=None, want_target=True,
                               workspace=None, packed=None) -> EmaAccum:
    """The accumulation with `prototype_learning`'s own arguments (salsanext_proto.py:337-339):
    out_feat (n, D), nearest (B, C, H, W), label (n,) or (B, H, W), feat_proto_sim (n, M, C)."""
    _need_cuda(out_feat=out_feat, nearest=nearest, label=label, feat_proto_sim=feat_proto_sim, gumbel=gumbel)
    if out_feat.dtype != torch.float32 or nearest.dtype != torch.float32 or feat_proto_sim.dtype != torch.float32:
        raise ValueError("out_feat / nearest_proto_distance / feat_proto_sim must be float32")
    if label.dtype != torch.int64:
        raise ValueError("label must be int64")
    if nearest.dim() != 4 or out_feat.dim() != 2 or feat_proto_sim.dim() != 3:
        raise ValueError("expected out_feat (n, D), nearest (B, C, H, W), feat_proto_sim (n, M, C)")
    B, C, H, W = nearest.shape
    n, D = out_feat.shape
    M = feat_proto_sim.shape[1]
    if n != B * H * W or feat_proto_sim.shape != (n, M, C) or label.numel() != n:
        raise ValueError("shape mismatch between out_feat / nearest / label / feat_proto_sim")
    if assign_mode is None:
        assign_mode = ASSIGN_GUMBEL_INJECTED if gumbel is not None else ASSIGN_GUMBEL_DEVICE
    if max_rows is None:
        max_rows = min(n, 1 << 17)
    if gumbel is not None and (gumbel.dtype != torch.float32 or gumbel.dim() != 2 or gumbel.shape[1] != M):
        raise ValueError("gumbel must be (rows, M) float32 in (class, pixel) row order")
    if workspace is None:
        nb = lib.c3d_proto_ema_workspace_bytes(B, C, H * W, D, M, max_rows)
        if nb == 0:
            raise ValueError("bad EMA shape")
        workspace = torch.empty((nb,), dtype=torch.uint8, device=out_feat.device)
    if packed is None:
        packed = torch.empty((C * M * D + C * M,), dtype=torch.float32, device=out_feat.device)
    target = torch.empty((n,), dtype=torch.float32, device=out_feat.device) if want_target else None
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if assign_mode == ASSIGN_GUMBEL_DEVICE else 0
    check(lib.c3d_proto_ema_accumulate_dense(
        _p(out_feat), _p(nearest), _p(label), _p(feat_proto_sim), B, D, H, W, C, M, int(ignore_label),
        int(max_rows), _p(gumbel), int(assign_mode), int(seed), _p(workspace), _p(packed), _p(target),
        _stream()))
    return EmaAccum(packed, target, workspace)


def proto_ema_apply(prototypes, packed, momentum, ignore_label=0, out=None, normalised_out=None,
                    seed_counters=None):
    """normalise sums -> EMA where count != 0 -> renormalise (salsanext_proto.py:379-394).
    normalised_out (C,M,D): also receives F.normalize(out), the form the bank's readers use;
    seed_counters (2,) int64: device step counters of the fused step ([1] is advanced here)."""
    _need_cuda(prototypes=prototypes, packed=packed)
    C, M, D = prototypes.shape
    if packed.numel() != C * M * D + C * M or packed.dtype != torch.float32:
        raise ValueError("packed must be (C*M*D + C*M,) float32")
    if out is None:
        out = torch.empty_like(prototypes)
    check(lib.c3d_proto_ema_apply(_p(prototypes), _p(packed), C, M, D, int(ignore_label),
                                  float(momentum), _p(out), _p(normalised_out), _p(seed_counters), _stream()))
    return out


def bank_normalise(prototypes, out=None):
    """F.normalize(prototypes, dim=-1) with the library's own kernel (bitwise what
    c3d_proto_ema_apply's `normalised_out` and the operators' internal normalisation produce)."""
    _need_cuda(prototypes=prototypes, out=out)
    if out is None:
        out = torch.empty_like(prototypes)
    D = prototypes.shape[-1]
    check(lib.c3d_proto_bank_normalise(_p(prototypes), prototypes.numel() // D, D, _p(out), _stream()))
    return out


STEP_SPLIT, STEP_SAMPLE, STEP_ACCUMULATE, STEP_LOSS_ROWS = 1, 2, 4, 8


def proto_step_workspace(batch, n_classes, hw, dim, sub_protos, num_anchor, max_rows, device):
    n = lib.c3d_proto_step_workspace_bytes(batch, n_classes, hw, dim, sub_protos, num_anchor, int(max_rows))
    if n == 0:
        raise ValueError("bad prototype-step shape")
    return torch.empty((n,), dtype=torch.uint8, device=device)


def proto_step_raw(phases, feats, probs, labels, keep_mask, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                   cfg, workspace, packed, loss_out, max_rows, ln_eps=1e-5, keep=None,
                   gumbel=None, assign_mode=ASSIGN_GUMBEL_DEVICE, seed=0, need_grad=True, proto_target=None,
                   bank_n=None, seed_counters=None, tensor_cores=False, cofill=None):
    """c3d_proto_step on pre-validated device tensors (no autograd, no allocation): the phases of
    the fused EMA-update + loss step (STEP_* bit mask) on one shared label split."""
    B, D, H, W = feats.shape
    C, M, _ = prototypes.shape
    check(lib.c3d_proto_step(
        _p(feats), _p(probs), _p(labels), _p(keep_mask), _p(prototypes), _p(ln_d_w), _p(ln_d_b), _p(ln_c_w),
        _p(ln_c_b), float(ln_eps), B, D, H, W, C, M, int(cfg.ignore_label), float(cfg.temperature),
        float(cfg.base_temperature), int(cfg.num_anchor), _p(keep), 0 if keep is None else keep.shape[0],
        _p(gumbel), int(assign_mode), int(seed), int(max_rows),
        (1 if need_grad else 0) | (2 if tensor_cores else 0), int(phases),
        _p(bank_n), _p(seed_counters), _p(workspace), _p(packed), _p(proto_target), _p(loss_out),
        *_cofill_args(cofill), _stream()))


def launch_count():
    """Kernel launches enqueued by the library since load."""
    return _lib.launch_count()


class profile:
    """Per-kernel device timing (CUDA events around each launch inside the library).

        with ops.profile("fill_zero_kernel") as prof: ...steps...
        ms, n = prof.read("fill_zero_kernel")
    """

    def __init__(self, kernel_name=""):
        self.kernel_name = kernel_name

    def __enter__(self):
        check(lib.c3d_profile_reset())
        check(lib.c3d_profile_enable(self.kernel_name.encode()))
        return self

    def __exit__(self, *exc):
        check(lib.c3d_profile_enable(None))
        return False

    @staticmethod
    def read(kernel_name=""):
        ms, n = ctypes.c_double(0), ctypes.c_longlong(0)
        check(lib.c3d_profile_read(kernel_name.encode(), ctypes.cast(ctypes.byref(ms), ctypes.c_void_p),
                                   ctypes.cast(ctypes.byref(n), ctypes.c_void_p)))
        return ms.value, n.value

    @staticmethod
    def names():
        buf = ctypes.create_string_buffer(4096)
        check(lib.c3d_profile_names(ctypes.cast(buf, ctypes.c_void_p), 4096))
        return [s for s in buf.value.decode().split(",") if s]

    @staticmethod
    def timeline():
        buf = ctypes.create_string_buffer(1 << 20)
        check(lib.c3d_profile_timeline(ctypes.cast(buf, ctypes.c_void_p), 1 << 20))
        rows = []
        for line in buf.value.decode().splitlines():
            name, a, b = line.split(",")
            rows.append((name, float(a), float(b)))
        return rows

    @staticmethod
    def all():
        return {n: profile.read(n) for n in profile.names()}
