// f2 -- un-projection gather + confusion matrix for a CSR batch of scans.
//
// Replaces, per scan, `unproj_argmax = argmax_2d[ii, uproj_y_idx[ii], uproj_x_idx[ii]]`
// (reference tasks/weak_segmentation/trainer.py:714-724) and `IOUEval.addBatch`
// (pc_processor/metrics/iou_eval.py:35-58: conf[pred, gt] += 1 through a CPU index_put).
// One thread per point: coalesced reads of px / py / label, one gathered class read, the
// per-point prediction written back (the input of the KNN post-processing), and a
// per-CTA shared-memory histogram flushed with 64-bit integer atomics (exact, order
// independent).  HBM-bound: 4+4+4(+8) B read and 4(8) B written per point.
#include "common.cuh"

namespace c3d {

__global__ void __launch_bounds__(256)
unproject_confusion_kernel(const void* __restrict__ proj_argmax, const void* __restrict__ px_,
                           const void* __restrict__ py_, const void* __restrict__ labels,
                           const int32_t* __restrict__ offsets, int batch, int total, int H, int W,
                           int C, int argmax64, int pxy64, int label64, void* __restrict__ unproj,
                           unsigned long long* __restrict__ conf, int32_t* __restrict__ flags) {
  extern __shared__ int32_t smem_i[];
  int32_t* s_off = smem_i;                 // [batch + 1]
  int32_t* s_hist = smem_i + batch + 1;    // [C * C] (only if conf)
  __shared__ int s_b0;
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  if (conf) for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  // each CTA owns one contiguous chunk of points and keeps its histogram in shared
  // memory for the whole chunk: C*C global atomics per CTA, not per 256 points
  const int chunk = (((total + gridDim.x - 1) / gridDim.x) + 255) & ~255;
  const int g_begin = blockIdx.x * chunk, g_end = min(g_begin + chunk, total);
  if (threadIdx.x == 0) s_b0 = scan_of(s_off, batch, min(g_begin, total - 1));
  __syncthreads();
  int b = s_b0;
  for (int g = g_begin + threadIdx.x; g < g_end; g += blockDim.x) {
    while (g >= s_off[b + 1]) ++b;
    const int x = pxy64 ? (int)reinterpret_cast<const long long*>(px_)[g] : reinterpret_cast<const int*>(px_)[g];
    const int y = pxy64 ? (int)reinterpret_cast<const long long*>(py_)[g] : reinterpret_cast<const int*>(py_)[g];
    int pred = 0;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const size_t o = (size_t)b * H * W + (size_t)y * W + x;
      pred = argmax64 ? (int)__ldg(reinterpret_cast<const long long*>(proj_argmax) + o)
                      : __ldg(reinterpret_cast<const int*>(proj_argmax) + o);
    } else {
      atomicOr(flags, 1);  // pixel index outside the image
    }
    if (unproj) {
      if (argmax64) reinterpret_cast<long long*>(unproj)[g] = pred;
      else reinterpret_cast<int*>(unproj)[g] = pred;
    }
    if (conf) {
      const int gt = label64 ? (int)reinterpret_cast<const long long*>(labels)[g]
                             : reinterpret_cast<const int*>(labels)[g];
      if (pred >= 0 && pred < C && gt >= 0 && gt < C) atomicAdd(&s_hist[pred * C + gt], 1);
      else atomicOr(flags, 2);  // class outside [0, C)
    }
  }
  if (conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
      const int v = s_hist[i];
      if (v) atomicAdd(conf + i, (unsigned long long)v);  // rows = pred, cols = gt (:55-58)
    }
  }
}

}  // namespace c3d

using namespace c3d;

extern "C" int c3d_unproject_confusion_batch(
    const void* proj_argmax, const void* px, const void* py, const void* labels,
    const int32_t* offsets, int batch, int64_t total_points, int proj_h, int proj_w, int nclasses,
    int argmax_is_i64, int pxy_is_i64, int label_is_i64, void* unproj_argmax, int64_t* conf_matrix,
    int32_t* status_flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size");
  C3D_REQUIRE(nclasses >= 1 && nclasses <= 96, "nclasses must be in [1, 96]");
  C3D_REQUIRE(proj_argmax && offsets && status_flags, "null pointer argument");
  C3D_REQUIRE(unproj_argmax || conf_matrix, "nothing to compute");
  C3D_REQUIRE(!conf_matrix || labels, "the confusion matrix needs per-point labels");
  if (total_points == 0) return C3D_OK;
  C3D_REQUIRE(px && py, "null per-point pointer");
  const int threads = 256;
  long long blocks = (total_points + threads - 1) / threads;
  const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  const size_t smem = ((size_t)(batch + 1) + (conf_matrix ? (size_t)nclasses * nclasses : 0)) * 4;
  KernelTimer kt__("unproject_confusion_kernel", stream);
  unproject_confusion_kernel<<<grid, threads, smem, stream>>>(
      proj_argmax, px, py, labels, offsets, batch, (int)total_points, proj_h, proj_w, nclasses,
      argmax_is_i64, pxy_is_i64, label_is_i64, unproj_argmax,
      reinterpret_cast<unsigned long long*>(conf_matrix), status_flags);
  return check_launch("unproject_confusion_kernel");
}
