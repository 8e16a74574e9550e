"""KNN post-processing with the reference's interface, running on the B200.

Mirrors pc_processor/postproc/knn.py:36-142: `KNN(params, nclasses)` with
params keys knn / search / sigma / cutoff; `forward(proj_range, unproj_range,
proj_argmax, px, py)` -> (P,) int64 labels; ValueError for an even window.
`forward_batch` takes a CSR batch of scans in one launch.
"""
import torch
import torch.nn as nn

from coarse3d_b200 import ops
from coarse3d_b200.ops import gaussian_kernel as get_gaussian_kernel  # noqa: F401  (knn.py:11)


class KNN(nn.Module):
    def __init__(self, params, nclasses, verbose=False):
        super().__init__()
        self.knn = params["knn"]
        self.search = params["search"]
        self.sigma = params["sigma"]
        self.cutoff = params["cutoff"]
        self.nclasses = nclasses
        if verbose:  # the reference prints this banner unconditionally (knn.py:39-52)
            print("*" * 80)
            print("Cleaning point-clouds with kNN post-processing")
            print("kNN parameters:")
            print("knn:", self.knn)
            print("search:", self.search)
            print("sigma:", self.sigma)
            print("cutoff:", self.cutoff)
            print("nclasses:", self.nclasses)
            print("*" * 80)
        self._inv_gauss = {}

    def _weights(self, device):
        key = (str(device), self.search, self.sigma)
        if key not in self._inv_gauss:
            self._inv_gauss[key] = (1 - ops.gaussian_kernel(self.search, self.sigma)).reshape(-1).to(device)
        return self._inv_gauss[key]

    def forward_batch(self, proj_range, unproj_range, proj_argmax, px, py, offsets, out_uint8=False):
        """CSR batch of scans in one launch.  out_uint8=True returns uint8 class ids instead of
        the reference's int64 (opt-in: an eighth of the bytes when the labels go to the host)."""
        if self.search % 2 == 0:
            raise ValueError("Nearest neighbor kernel must be odd number")
        return ops.knn_batch(proj_range, proj_argmax, unproj_range, px, py, offsets, self.knn,
                             self.search, self.sigma, self.cutoff, self.nclasses,
                             inv_gauss=self._weights(proj_range.device), out_uint8=out_uint8)

    def forward(self, proj_range, unproj_range, proj_argmax, px, py):
        ''' Un-batched, like the reference (knn.py:55-58). '''
        if self.search % 2 == 0:
            raise ValueError("Nearest neighbor kernel must be odd number")
        if not proj_range.is_cuda:
            raise RuntimeError("coarse3d_b200 KNN runs on CUDA tensors only (no CPU path)")
        idt = torch.int64 if px.dtype != torch.int32 else torch.int32
        offsets = torch.tensor([0, unproj_range.shape[0]], dtype=torch.int32, device=proj_range.device)
        out = ops.knn_batch(
            proj_range.contiguous().float()[None], proj_argmax.contiguous().to(idt)[None],
            unproj_range.contiguous().float(), px.contiguous().to(idt), py.contiguous().to(idt),
            offsets, self.knn, self.search, self.sigma, self.cutoff, self.nclasses,
            inv_gauss=self._weights(proj_range.device))
        return out.long()
