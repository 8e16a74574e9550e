// a1 -- spherical range projection + min-depth z-buffer for a CSR batch of scans.
//
// Replaces RangeProjection.doProjection (reference
// pc_processor/dataset/preprocess/projection.py:43-115).  Compiled with
// -fmad=false: every float32 operation after the transcendentals must round
// exactly like the reference's separate numpy ufuncs.
//
// Two kernels, both HBM-bound:
//   project_points : 1 thread / point.  128-bit load of (x,y,z,i); depth, yaw,
//                    pitch, pixel; writes uproj_x/y/depth; 64-bit atomicMin of
//                    (depth_key << 32 | point_index) into the z-buffer.
//   resolve_pixels : 1 thread / pixel.  Decodes the winner, gathers its point,
//                    writes range / pointcloud / idx / mask with coalesced
//                    (128-bit for the pointcloud) stores.
//
// Transcendentals.  The oracle's rule is the correctly rounded float32
// arctan2 / arcsin.  Evaluating both in fp64 for every point costs ~150 DFMA
// per point and would make the kernel FP64-pipe bound (B200: 64 DFMA/clk/SM),
// so the kernel first evaluates atan2f/asinf (<= 2 ulp) and only re-evaluates
// in fp64 when the scaled coordinate lies within a proven error bound of a
// pixel boundary (the only case where the floor could differ): ~0.5 % of
// points.  C3D_PROJECT_F64_ONLY=1 forces the fp64 path for every point.

#include "common.cuh"

namespace c3d {

__device__ __forceinline__ uint32_t depth_key(float d) {
  uint32_t b = __float_as_uint(d);
  return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_depth(uint32_t k) {
  return __uint_as_float((k >> 31) ? (k & 0x7fffffffu) : ~k);
}

struct ProjParams {
  float abs_left, fov_hori, abs_down, fov_vert;  // float32-rounded Python floats
  float wf, hf, wmax, hmax;
  float tol_x, tol_y;  // hybrid guard bands, in pixels
  float sx, sy;        // fast path: W / fov_hori, H / fov_vert
  int H, W;
};

template <bool kHybrid>
__device__ __forceinline__ void pixel_of(float x, float y, float q, const ProjParams& p,
                                         int& px, int& py, bool& nan) {
  float yaw, pitch, fx, fy;
  if (kHybrid) {
    // fast estimate: <= 2 ulp transcendentals and one multiply by a pre-divided
    // scale instead of the reference's divide-then-multiply (error inside the guard band)
    yaw = -atan2f(y, x);
    pitch = asinf(q);
    fx = (yaw + p.abs_left) * p.sx;
    fy = p.hf - (pitch + p.abs_down) * p.sy;
    // !(a > b) also catches NaN
    bool near_x = !(fabsf(fx - rintf(fx)) > p.tol_x);
    bool near_y = !(fabsf(fy - rintf(fy)) > p.tol_y);
    if (near_x) {  // exact chain (projection.py:62-64,73): correctly rounded angle, IEEE ops
      yaw = -(float)atan2((double)y, (double)x);
      fx = ((yaw + p.abs_left) / p.fov_hori) * p.wf;
    }
    if (near_y) {  // projection.py:66-68,74
      pitch = (float)asin((double)q);
      fy = (1.0f - (pitch + p.abs_down) / p.fov_vert) * p.hf;
    }
  } else {
    yaw = -(float)atan2((double)y, (double)x);
    pitch = (float)asin((double)q);
    fx = ((yaw + p.abs_left) / p.fov_hori) * p.wf;
    fy = (1.0f - (pitch + p.abs_down) / p.fov_vert) * p.hf;
  }
  nan = (fx != fx) || (fy != fy);
  px = (int)fmaxf(fminf(p.wmax, floorf(fx)), 0.0f);
  py = (int)fmaxf(fminf(p.hmax, floorf(fy)), 0.0f);
}

template <bool kHybrid, bool kC4>
__global__ void __launch_bounds__(256)
project_points_kernel(const float* __restrict__ points, int c_in,
                      const int32_t* __restrict__ offsets, int batch, int total,
                      const float* __restrict__ depth_override, ProjParams p,
                      int32_t* __restrict__ upx, int32_t* __restrict__ upy,
                      float* __restrict__ udepth, unsigned long long* __restrict__ zbuf,
                      int32_t* __restrict__ flags) {
  extern __shared__ int32_t s_off[];
  __shared__ int s_b0;
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  if (threadIdx.x == 0) s_b0 = scan_of(s_off, batch, min((int)(blockIdx.x * blockDim.x), total - 1));
  __syncthreads();
  const int HW = p.H * p.W;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  int b = s_b0;
  while (g >= s_off[b + 1]) ++b;  // a CTA spans at most a few scans
  float x, y, z;
  if (kC4) {
    float4 v = __ldg(reinterpret_cast<const float4*>(points) + g);  // keep in L2 for resolve
    x = v.x; y = v.y; z = v.z;
  } else {
    const float* r = points + (size_t)g * c_in;
    x = r[0]; y = r[1]; z = r[2];
  }
  float depth = depth_override ? depth_override[g] : sqrtf((x * x + y * y) + z * z);
  float q = z / depth;
  int px, py; bool nan;
  pixel_of<kHybrid>(x, y, q, p, px, py, nan);
  upx[g] = px; upy[g] = py; udepth[g] = depth;
  unsigned long long key =
      ((unsigned long long)depth_key(depth) << 32) | (uint32_t)(g - s_off[b]);
  atomicMin(zbuf + (size_t)b * HW + py * p.W + px, key);
  if (nan) atomicOr(flags, 1);
}

// Two pixels per thread (two independent key -> gather -> store chains in flight).
// Every z-buffer entry is reset to "empty" after it is read, so the workspace is
// left clean for the next call and needs no memset.
template <bool kC4>
__global__ void __launch_bounds__(256)
resolve_pixels_kernel(const float* __restrict__ points, int c_in,
                      const int32_t* __restrict__ offsets, int HW, long long total_px,
                      unsigned long long* __restrict__ zbuf,
                      float* __restrict__ proj_range, float* __restrict__ proj_pc,
                      int32_t* __restrict__ proj_idx, int32_t* __restrict__ proj_mask) {
  const long long q0 = ((long long)blockIdx.x * blockDim.x) * 2 + threadIdx.x;
  unsigned long long key[2];
  bool in[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const long long q = q0 + j * 256;
    in[j] = q < total_px;
    key[j] = in[j] ? __ldcs(zbuf + q) : ~0ull;
  }
  float4 pt[2];
  size_t row[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    pt[j] = make_float4(-1.f, -1.f, -1.f, -1.f);
    row[j] = 0;
    if (key[j] != ~0ull) {
      const long long q = q0 + j * 256;
      row[j] = (size_t)__ldg(offsets + (int)(q / HW)) + (uint32_t)key[j];
      if (kC4) pt[j] = __ldg(reinterpret_cast<const float4*>(points) + row[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (!in[j]) continue;
    const long long q = q0 + j * 256;
    const bool valid = key[j] != ~0ull;
    const int idx = valid ? (int)(uint32_t)key[j] : -1;
    proj_range[q] = valid ? key_depth((uint32_t)(key[j] >> 32)) : -1.0f;
    proj_idx[q] = idx;
    proj_mask[q] = idx > 0;  // projection.py:113
    if (kC4) {
      st_stream(reinterpret_cast<float4*>(proj_pc) + q, pt[j]);
    } else {
      for (int c = 0; c < c_in; ++c)
        proj_pc[q * c_in + c] = valid ? points[row[j] * c_in + c] : -1.0f;
    }
    if (valid) zbuf[q] = ~0ull;
  }
}

// ---------------------------------------------------------------- f1 -------
// Resolve pass fused with the projection's caller: the label images and the 5-channel
// network input are written straight from the z-buffer winners
// (wss_sem_kitti_loader.py:124-132,159-172; trainer.py:600-608), so neither the
// (H,W,4) projected point cloud nor a CPU-side gather is needed.
__global__ void __launch_bounds__(256)
resolve_assemble_kernel(const float* __restrict__ points, const int32_t* __restrict__ offsets,
                        int HW, long long total_px, unsigned long long* __restrict__ zbuf,
                        const void* __restrict__ sem_label, const void* __restrict__ weak_label,
                        int label_is_u8,
                        const float* __restrict__ mean, const float* __restrict__ stdv,
                        float* __restrict__ proj_range, int32_t* __restrict__ proj_idx,
                        float* __restrict__ feature, long long* __restrict__ train_label,
                        long long* __restrict__ eval_label) {
  const long long q0 = ((long long)blockIdx.x * blockDim.x) * 2 + threadIdx.x;
  unsigned long long key[2]; bool in[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const long long q = q0 + j * 256;
    in[j] = q < total_px;
    key[j] = in[j] ? __ldcs(zbuf + q) : ~0ull;
  }
  float4 pt[2]; int sl[2], wl[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    pt[j] = make_float4(-1.f, -1.f, -1.f, -1.f); sl[j] = 0; wl[j] = 0;
    if (key[j] != ~0ull) {
      const long long q = q0 + j * 256;
      const size_t row = (size_t)__ldg(offsets + (int)(q / HW)) + (uint32_t)key[j];
      pt[j] = __ldg(reinterpret_cast<const float4*>(points) + row);
      if (label_is_u8) {
        if (sem_label) sl[j] = __ldg(reinterpret_cast<const uint8_t*>(sem_label) + row);
        if (weak_label) wl[j] = __ldg(reinterpret_cast<const uint8_t*>(weak_label) + row);
      } else {
        if (sem_label) sl[j] = __ldg(reinterpret_cast<const int32_t*>(sem_label) + row);
        if (weak_label) wl[j] = __ldg(reinterpret_cast<const int32_t*>(weak_label) + row);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (!in[j]) continue;
    const long long q = q0 + j * 256;
    const bool valid = key[j] != ~0ull;
    const float rng = valid ? key_depth((uint32_t)(key[j] >> 32)) : -1.0f;
    if (proj_range) proj_range[q] = rng;
    if (proj_idx) proj_idx[q] = valid ? (int)(uint32_t)key[j] : -1;
    // loader :124-132 builds float32 label images, trainer :600-601 casts them to int64
    if (eval_label) eval_label[q] = (long long)(float)sl[j];
    if (train_label) train_label[q] = (long long)(float)wl[j];
    float f[5];
    f[0] = rng; f[1] = pt[j].x; f[2] = pt[j].y; f[3] = pt[j].z;
    f[4] = ((pt[j].w != -1.0f) ? 1.0f : 0.0f) * pt[j].w;          // loader :161-164
    const int b = (int)(q / HW);
    const long long pix = q - (long long)b * HW;
    float* dst = feature + (size_t)b * 5 * HW + pix;
    if (mean) {
      const float m = (sl[j] > 0) ? 1.0f : 0.0f;                    // eval_mask, trainer :603
#pragma unroll
      for (int c = 0; c < 5; ++c) f[c] = (f[c] - mean[c]) / stdv[c] * m;  // trainer :604-608
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) dst[(size_t)c * HW] = f[c];
    if (valid) zbuf[q] = ~0ull;
  }
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_project_workspace_bytes(int batch, int proj_h, int proj_w) {
  if (batch <= 0 || proj_h <= 0 || proj_w <= 0) return 0;
  return (size_t)batch * proj_h * proj_w * sizeof(unsigned long long);
}

extern "C" int c3d_project_batch(
    const float* points, int c_in, const int32_t* offsets, int batch, int64_t total_points,
    const float* depth_override, double abs_fov_left, double fov_hori, double abs_fov_down,
    double fov_vert, int proj_h, int proj_w, float* proj_range, float* proj_pointcloud,
    int32_t* proj_idx, int32_t* proj_mask, int32_t* uproj_x_idx, int32_t* uproj_y_idx,
    float* uproj_depth, void* workspace, int workspace_flags, int32_t* status_flags,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d], got %d", kMaxBatch, batch);
  C3D_REQUIRE(c_in >= 3, "c_in must be >= 3, got %d", c_in);
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size %dx%d", proj_h, proj_w);
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE((long long)batch * proj_h * proj_w < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(fov_hori > 0 && fov_vert > 0, "field of view must be positive");
  C3D_REQUIRE(offsets && proj_range && proj_pointcloud && proj_idx && proj_mask && workspace &&
              status_flags, "null pointer argument");
  C3D_REQUIRE(total_points == 0 || (points && uproj_x_idx && uproj_y_idx && uproj_depth),
              "null per-point pointer");

  ProjParams p;
  p.abs_left = (float)abs_fov_left; p.fov_hori = (float)fov_hori;
  p.abs_down = (float)abs_fov_down; p.fov_vert = (float)fov_vert;
  p.wf = (float)proj_w; p.hf = (float)proj_h;
  p.wmax = (float)(proj_w - 1); p.hmax = (float)(proj_h - 1);
  p.H = proj_h; p.W = proj_w;
  // Guard bands: |fast - exact| chain error bounds (projection),
  // with a 2.5x margin.  atan2f/asinf <= 2 ulp (CUDA math API), |yaw| <= pi,
  // |pitch| <= pi/2, then add / divide / multiply roundings.
  p.tol_x = p.wf * 1.0e-6f;
  p.tol_y = p.hf * (2.0e-6f / p.fov_vert + 1.0e-6f);
  p.sx = p.wf / p.fov_hori;
  p.sy = p.hf / p.fov_vert;

  const long long total_px = (long long)batch * proj_h * proj_w;
  auto* zbuf = reinterpret_cast<unsigned long long*>(workspace);
  if (!(workspace_flags & 1)) {
    KernelTimer kt__("zbuf_memset", stream);
    C3D_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)total_px * sizeof(unsigned long long), stream));
  }

  const bool c4 = (c_in == 4) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(proj_pointcloud) & 15) == 0);
  const bool hybrid = !(workspace_flags & 2);
  const int threads = 256;
  if (total_points > 0) {
    int grid = (int)((total_points + threads - 1) / threads);  // short CTAs (overlap)
    size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
#define LAUNCH_PP(HY, C4)                                                                   \
  project_points_kernel<HY, C4><<<grid, threads, smem, stream>>>(                           \
      points, c_in, offsets, batch, (int)total_points, depth_override, p, uproj_x_idx,      \
      uproj_y_idx, uproj_depth, zbuf, status_flags)
    {
      KernelTimer kt__("project_points_kernel", stream);
      if (hybrid) { if (c4) LAUNCH_PP(true, true); else LAUNCH_PP(true, false); }
      else        { if (c4) LAUNCH_PP(false, true); else LAUNCH_PP(false, false); }
    }
#undef LAUNCH_PP
    int rc = check_launch("project_points_kernel");
    if (rc) return rc;
  }
  {
    int grid = (int)((total_px + 2 * threads - 1) / (2 * threads));
    KernelTimer kt__("resolve_pixels_kernel", stream);
    if (c4)
      resolve_pixels_kernel<true><<<grid, threads, 0, stream>>>(
          points, c_in, offsets, proj_h * proj_w, total_px, zbuf, proj_range, proj_pointcloud,
          proj_idx, proj_mask);
    else
      resolve_pixels_kernel<false><<<grid, threads, 0, stream>>>(
          points, c_in, offsets, proj_h * proj_w, total_px, zbuf, proj_range, proj_pointcloud,
          proj_idx, proj_mask);
    int rc = check_launch("resolve_pixels_kernel");
    if (rc) return rc;
  }
  return C3D_OK;
}

extern "C" int c3d_project_assemble_batch(
    const float* points, const int32_t* offsets, int batch, int64_t total_points,
    const float* depth_override, const void* sem_label, const void* weak_label, int label_is_u8,
    const float* img_mean, const float* img_std, double abs_fov_left, double fov_hori,
    double abs_fov_down, double fov_vert, int proj_h, int proj_w, float* feature,
    int64_t* train_label, int64_t* eval_label, float* proj_range, int32_t* proj_idx,
    int32_t* uproj_x_idx, int32_t* uproj_y_idx, float* uproj_depth, void* workspace,
    int workspace_flags, int32_t* status_flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d], got %d", kMaxBatch, batch);
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size %dx%d", proj_h, proj_w);
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE((long long)batch * proj_h * proj_w < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(fov_hori > 0 && fov_vert > 0, "field of view must be positive");
  C3D_REQUIRE(offsets && feature && workspace && status_flags, "null pointer argument");
  C3D_REQUIRE(total_points == 0 || (points && uproj_x_idx && uproj_y_idx && uproj_depth),
              "null per-point pointer");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(points) & 15) == 0, "points must be 16 B aligned (x,y,z,i rows)");
  C3D_REQUIRE((img_mean == nullptr) == (img_std == nullptr), "img_mean and img_std go together");
  C3D_REQUIRE(!img_mean || sem_label, "normalisation needs sem_label (eval_mask = eval_label > 0)");
  C3D_REQUIRE(!eval_label || sem_label, "eval_label output needs sem_label");
  C3D_REQUIRE(!train_label || weak_label, "train_label output needs weak_label");

  ProjParams p;
  p.abs_left = (float)abs_fov_left; p.fov_hori = (float)fov_hori;
  p.abs_down = (float)abs_fov_down; p.fov_vert = (float)fov_vert;
  p.wf = (float)proj_w; p.hf = (float)proj_h;
  p.wmax = (float)(proj_w - 1); p.hmax = (float)(proj_h - 1);
  p.H = proj_h; p.W = proj_w;
  p.tol_x = p.wf * 1.0e-6f;
  p.tol_y = p.hf * (2.0e-6f / p.fov_vert + 1.0e-6f);
  p.sx = p.wf / p.fov_hori;
  p.sy = p.hf / p.fov_vert;
  const long long total_px = (long long)batch * proj_h * proj_w;
  auto* zbuf = reinterpret_cast<unsigned long long*>(workspace);
  if (!(workspace_flags & 1)) {
    KernelTimer kt__("zbuf_memset", stream);
    C3D_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)total_px * sizeof(unsigned long long), stream));
  }
  const bool hybrid = !(workspace_flags & 2);
  const int threads = 256;
  if (total_points > 0) {
    int grid = (int)((total_points + threads - 1) / threads);
    size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
    KernelTimer kt__("project_points_kernel", stream);
    if (hybrid)
      project_points_kernel<true, true><<<grid, threads, smem, stream>>>(
          points, 4, offsets, batch, (int)total_points, depth_override, p, uproj_x_idx, uproj_y_idx,
          uproj_depth, zbuf, status_flags);
    else
      project_points_kernel<false, true><<<grid, threads, smem, stream>>>(
          points, 4, offsets, batch, (int)total_points, depth_override, p, uproj_x_idx, uproj_y_idx,
          uproj_depth, zbuf, status_flags);
    int rc = check_launch("project_points_kernel");
    if (rc) return rc;
  }
  {
    int grid = (int)((total_px + 2 * threads - 1) / (2 * threads));
    KernelTimer kt__("resolve_assemble_kernel", stream);
    resolve_assemble_kernel<<<grid, threads, 0, stream>>>(
        points, offsets, proj_h * proj_w, total_px, zbuf, sem_label, weak_label, label_is_u8, img_mean, img_std,
        proj_range, proj_idx, feature, (long long*)train_label, (long long*)eval_label);
    return check_launch("resolve_assemble_kernel");
  }
}
