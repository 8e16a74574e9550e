// a4 -- KNN range-image label vote for a CSR batch of scans.
//
// Replaces KNN.forward (reference pc_processor/postproc/knn.py:54-142) without
// materialising the two (1, S*S, H*W) unfolds or the three (1, S*S, P) gathers.
//
// One thread per point: the S x S window of the range image is read through
// the read-only path (the image is <= 0.5 MB per scan and stays L1/L2
// resident), the k nearest slots are selected in registers by (distance, slot)
// -- the tie rule fixed by the oracle -- and only those k class labels are
// fetched.  The vote is an O(k^2) register comparison; no (C+1)-wide one-hot.
//
// Compiled with -fmad=false (|a-b| * w must round like torch's separate ops).
#include <math_constants.h>

#include "common.cuh"

namespace c3d {

template <typename IdxT, typename LabT, int S, int KMAX>
__global__ void __launch_bounds__(256)
knn_vote_kernel(const float* __restrict__ proj_range, const LabT* __restrict__ proj_argmax,
                const float* __restrict__ unproj_range, const IdxT* __restrict__ px_,
                const IdxT* __restrict__ py_, const int32_t* __restrict__ offsets, int batch,
                int total, int H, int W, int knn, float cutoff, int nclasses,
                const float* __restrict__ inv_gauss, LabT* __restrict__ out) {
  constexpr int S2 = S * S;
  constexpr int PAD = (S - 1) / 2;
  extern __shared__ int32_t s_off[];
  __shared__ float s_w[S2];
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  for (int i = threadIdx.x; i < S2; i += blockDim.x) s_w[i] = inv_gauss[i];
  __syncthreads();
  const int HW = H * W;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    const int b = scan_of(s_off, batch, g);
    const float r = __ldcs(unproj_range + g);
    const int x0 = (int)__ldcs(px_ + g), y0 = (int)__ldcs(py_ + g);
    const float* img = proj_range + (size_t)b * HW;

    float d[S2];
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = y0 + dy - PAD;
      const bool yin = (y >= 0) && (y < H);
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        const int x = x0 + dx - PAD;
        float v = 0.0f;  // F.unfold zero padding (knn.py:79-81)
        if (yin && x >= 0 && x < W) v = __ldg(img + y * W + x);
        if (v < 0.0f) v = CUDART_INF_F;              // knn.py:90
        if (dy == PAD && dx == PAD) v = r;           // knn.py:93-94
        d[dy * S + dx] = fabsf(v - r) * s_w[dy * S + dx];  // knn.py:97,107
      }
    }

    // k smallest by (distance, slot), ascending
    int sel_cls[KMAX];
    float prev_d = -CUDART_INF_F;
    int prev_s = -1;
    const LabT* cimg = proj_argmax + (size_t)b * HW;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      sel_cls[j] = 0;
      if (j < knn) {
        float best_d = CUDART_INF_F;
        int best_s = -1;
#pragma unroll
        for (int s = 0; s < S2; ++s) {
          const bool after = (d[s] > prev_d) || (d[s] == prev_d && s > prev_s);
          const bool better = (best_s < 0) ? true : (d[s] < best_d);
          if (after && better) { best_d = d[s]; best_s = s; }
        }
        if (best_s >= 0) {
          prev_d = best_d; prev_s = best_s;
          const int y = y0 + best_s / S - PAD, x = x0 + best_s % S - PAD;
          int c = 0;  // zero padding => class 0 (knn.py:114-116)
          if (y >= 0 && y < H && x >= 0 && x < W) c = (int)__ldg(cimg + y * W + x);
          if (cutoff > 0.0f && best_d > cutoff) c = nclasses;  // knn.py:124-127
          sel_cls[j] = c;
        }
      }
    }

    // vote over classes 1..C-1, first maximum wins (knn.py:131-137)
    int best_c = 1, best_n = 0;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int c = sel_cls[j];
      if (j < knn && c >= 1 && c < nclasses) {
        int n = 0;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) n += (i < knn && sel_cls[i] == c) ? 1 : 0;
        if (n > best_n || (n == best_n && c < best_c)) { best_n = n; best_c = c; }
      }
    }
    out[g] = (LabT)best_c;
  }
}

template <typename IdxT, typename LabT>
int launch_knn(const float* proj_range, const void* proj_argmax, const float* unproj_range,
               const void* px, const void* py, const int32_t* offsets, int batch, int total,
               int H, int W, int knn, int search, float cutoff, int nclasses,
               const float* inv_gauss, void* out, cudaStream_t stream) {
  const int threads = 256;
  const int grid = wave_grid(total, threads, 8);
  const size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
  KernelTimer timer("knn_vote_kernel", stream);
#define LAUNCH_KNN(S_, K_)                                                                  \
  knn_vote_kernel<IdxT, LabT, S_, K_><<<grid, threads, smem, stream>>>(                     \
      proj_range, (const LabT*)proj_argmax, unproj_range, (const IdxT*)px, (const IdxT*)py, \
      offsets, batch, total, H, W, knn, cutoff, nclasses, inv_gauss, (LabT*)out)
  if (search == 3) { LAUNCH_KNN(3, 9); }
  else if (search == 5 && knn <= 8) { LAUNCH_KNN(5, 8); }
  else if (search == 5) { LAUNCH_KNN(5, 25); }
  else if (search == 7 && knn <= 8) { LAUNCH_KNN(7, 8); }
  else if (search == 7 && knn <= 16) { LAUNCH_KNN(7, 16); }
  else if (search == 9 && knn <= 16) { LAUNCH_KNN(9, 16); }
  else {
    set_error("unsupported KNN window/k: search=%d knn=%d (search in {3,5,7,9})", search, knn);
    return C3D_UNSUPPORTED;
  }
#undef LAUNCH_KNN
  return check_launch("knn_vote_kernel");
}

}  // namespace c3d

using namespace c3d;

extern "C" int c3d_knn_batch(const float* proj_range, const void* proj_argmax,
                             const float* unproj_range, const void* px, const void* py,
                             const int32_t* offsets, int batch, int64_t total_points, int proj_h,
                             int proj_w, int knn, int search, float cutoff, int nclasses,
                             const float* inv_gauss, int pxy_is_i64, int label_is_i64,
                             void* out_labels, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(search % 2 == 1, "Nearest neighbor kernel must be odd number");  // knn.py:72-73
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(knn >= 1 && knn <= search * search, "knn must be in [1, search^2]");
  C3D_REQUIRE(nclasses >= 2, "nclasses must be >= 2");
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size");
  C3D_REQUIRE(proj_range && proj_argmax && offsets && inv_gauss, "null pointer argument");
  if (total_points == 0) return C3D_OK;
  C3D_REQUIRE(unproj_range && px && py && out_labels, "null per-point pointer");
#define KNN_ARGS proj_range, proj_argmax, unproj_range, px, py, offsets, batch, (int)total_points, \
                 proj_h, proj_w, knn, search, cutoff, nclasses, inv_gauss, out_labels, stream
  if (pxy_is_i64 && label_is_i64) return launch_knn<long long, long long>(KNN_ARGS);
  if (pxy_is_i64) return launch_knn<long long, int>(KNN_ARGS);
  if (label_is_i64) return launch_knn<int, long long>(KNN_ARGS);
  return launch_knn<int, int>(KNN_ARGS);
#undef KNN_ARGS
}
