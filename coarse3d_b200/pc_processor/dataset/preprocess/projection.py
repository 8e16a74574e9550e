"""RangeProjection with the reference's interface, running on the B200.

Mirrors pc_processor/dataset/preprocess/projection.py:4-115 (constructor
arguments, assertions, `doProjection(pointcloud, depth=None)` returning four
numpy arrays, the `cached_data` dict).  The arithmetic runs in
`c3d_project_batch`; this class only moves numpy arrays to and from the device.
`doProjectionBatch` is the throughput entry point (device tensors, CSR batch).
"""
import numpy as np
import torch

from coarse3d_b200 import ops


class RangeProjection(object):
    """project 3d point cloud to 2d data with range projection"""

    def __init__(self, fov_up=3, fov_down=-25, proj_w=512, proj_h=64, fov_left=-180,
                 fov_right=180, device=None):
        # check params (projection.py:17-26)
        assert (
            fov_up >= 0 and fov_down <= 0
        ), "require fov_up >= 0 and fov_down <= 0, while fov_up/fov_down is {}/{}".format(
            fov_up, fov_down)
        assert (
            fov_right >= 0 and fov_left <= 0
        ), "require fov_right >= 0 and fov_left <= 0, while fov_right/fov_left is {}/{}".format(
            fov_right, fov_left)
        # params of fov angles (projection.py:29-35)
        self.fov_up = fov_up / 180.0 * np.pi
        self.fov_down = fov_down / 180.0 * np.pi
        self.fov_vert = abs(self.fov_up) + abs(self.fov_down)
        self.fov_left = fov_left / 180.0 * np.pi
        self.fov_right = fov_right / 180.0 * np.pi
        self.fov_hori = abs(self.fov_left) + abs(self.fov_right)
        self.proj_w = proj_w
        self.proj_h = proj_h
        self.cached_data = {}
        self.device = torch.device("cuda" if device is None else device)
        self._one = None  # offsets tensor cache for single scans

    @property
    def fov(self):
        return ops.Fov(abs(self.fov_left), self.fov_hori, abs(self.fov_down), self.fov_vert)

    def doProjectionBatch(self, points, offsets, depth=None, buffers=None):
        """CSR batch on the device: see ops.project_batch."""
        return ops.project_batch(points, offsets, self.fov, self.proj_h, self.proj_w, depth,
                                 buffers)

    def doProjectionAssembleBatch(self, points, offsets, sem_label=None, weak_label=None,
                                  img_mean=None, img_std=None, depth=None, buffers=None):
        """CSR batch on the device, fused with what the loaders / trainer build from the
        projection (wss_sem_kitti_loader.py:124-172, trainer.py:600-608): the 5-channel
        input (optionally normalised) and the int64 train / eval label images.  See
        ops.project_assemble_batch."""
        return ops.project_assemble_batch(points, offsets, self.fov, self.proj_h, self.proj_w,
                                          sem_label, weak_label, img_mean, img_std, depth, buffers)

    def doProjection(self, pointcloud: np.ndarray, depth: np.ndarray = None):
        self.cached_data = {}
        pts = torch.from_numpy(np.ascontiguousarray(pointcloud, dtype=np.float32))
        n = pts.shape[0]
        pts = pts.to(self.device, non_blocking=True)
        d = None
        if depth is not None:
            d = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).to(
                self.device, non_blocking=True)
        offsets = torch.tensor([0, n], dtype=torch.int32, device=self.device)
        out = ops.project_batch(pts, offsets, self.fov, self.proj_h, self.proj_w, d)
        if int(out.flags.item()) & 1:
            # the reference dies here with an out-of-range fancy index
            raise ValueError("RangeProjection: NaN pixel coordinate (a point has depth 0)")
        self.cached_data["uproj_x_idx"] = out.uproj_x_idx.cpu().numpy()
        self.cached_data["uproj_y_idx"] = out.uproj_y_idx.cpu().numpy()
        self.cached_data["uproj_depth"] = out.uproj_depth.cpu().numpy()
        return (out.proj_pointcloud[0].cpu().numpy(), out.proj_range[0].cpu().numpy(),
                out.proj_idx[0].cpu().numpy(), out.proj_mask[0].cpu().numpy())
