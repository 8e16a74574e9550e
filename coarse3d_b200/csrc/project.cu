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

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

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

// L2 residency hints of the two-kernel form.  The z-buffer (8 B per pixel, 67 MB at batch 64) is
// written by the point pass and consumed by the resolve pass right after: its lines are marked
// evict-last so that the streamed per-point traffic does not push them out to DRAM and back.
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_min_u64_hint(unsigned long long* addr, unsigned long long v,
                                                 unsigned long long pol) {
  asm volatile("red.global.min.L2::cache_hint.u64 [%0], %1, %2;\n" :: "l"(addr), "l"(v), "l"(pol) : "memory");
}

constexpr int kProjZeroPage = 8192;   // zero page of the carried fill (common.cuh)

// kFill: the CTAs also carry a share of an unrelated zero fill (the loss's dense gradient).
template <bool kHybrid, bool kC4, bool kFill>
__global__ void __launch_bounds__(256)
project_points_kernel(const float* __restrict__ points, int c_in,
                      const int32_t* __restrict__ offsets, int batch, int total,
                      const float* __restrict__ depth_override, ProjParams p,
                      int32_t* __restrict__ upx, int32_t* __restrict__ upy,
                      float* __restrict__ udepth, unsigned long long* __restrict__ zbuf,
                      int32_t* __restrict__ flags, FillShare fill) {
  extern __shared__ int32_t s_off[];
  __shared__ int s_b0;
  __shared__ __align__(128) float4 s_zero[kFill ? kProjZeroPage / 16 : 1];
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  if (kFill) carrier_init(s_zero, kProjZeroPage);
  __syncthreads();
  if (threadIdx.x == 0) {
    s_b0 = scan_of(s_off, batch, min((int)(blockIdx.x * blockDim.x), total - 1));
    if (kFill) carrier_issue(fill, s_zero, kProjZeroPage, blockIdx.x, gridDim.x, 0, 1);
  }
  __syncthreads();
  const int HW = p.H * p.W;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) { if (kFill && threadIdx.x == 0) bulk_wait_read_all(); return; }
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
  __stcs(upx + g, px); __stcs(upy + g, py); __stcs(udepth + g, depth);   // written once, streamed
  unsigned long long key =
      ((unsigned long long)depth_key(depth) << 32) | (uint32_t)(g - s_off[b]);
  red_min_u64_hint(zbuf + (size_t)b * HW + py * p.W + px, key, policy_evict_last());
  if (nan) atomicOr(flags, 1);
  if (kFill && threadIdx.x == 0) bulk_wait_read_all();   // the zero page must outlive the copies reading it
}

// Two pixels per thread (two independent key -> gather -> store chains in flight).
// Every z-buffer entry is reset to "empty" after it is read, so the workspace is
// left clean for the next call and needs no memset.
template <bool kC4, bool kFill>
__global__ void __launch_bounds__(256)
resolve_pixels_kernel(const float* __restrict__ points, int c_in,
                      const int32_t* __restrict__ offsets, int HW, long long total_px,
                      unsigned long long* __restrict__ zbuf,
                      float* __restrict__ proj_range, float* __restrict__ proj_pc,
                      int32_t* __restrict__ proj_idx, int32_t* __restrict__ proj_mask, FillShare fill) {
  __shared__ __align__(128) float4 s_zero[kFill ? kProjZeroPage / 16 : 1];
  if (kFill) {
    carrier_init(s_zero, kProjZeroPage);
    __syncthreads();
    if (threadIdx.x == 0) carrier_issue(fill, s_zero, kProjZeroPage, blockIdx.x, gridDim.x, 0, 1);
  }
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
  if (kFill && threadIdx.x == 0) bulk_wait_read_all();
}

// ---------------------------------------------------------------- f1 -------
// Resolve pass fused with the projection's caller: the label images and the 5-channel
// network input are written straight from the z-buffer winners
// (wss_sem_kitti_loader.py:124-132,159-172; trainer.py:600-608), so neither the
// (H,W,4) projected point cloud nor a CPU-side gather is needed.
template <bool kFill>
__global__ void __launch_bounds__(256)
resolve_assemble_kernel(const float* __restrict__ points, const int32_t* __restrict__ offsets,
                        int HW, long long total_px, unsigned long long* __restrict__ zbuf,
                        const void* __restrict__ sem_label, const void* __restrict__ weak_label,
                        int label_is_u8,
                        const float* __restrict__ mean, const float* __restrict__ stdv,
                        float* __restrict__ proj_range, int32_t* __restrict__ proj_idx,
                        float* __restrict__ feature, long long* __restrict__ train_label,
                        long long* __restrict__ eval_label, FillShare fill) {
  __shared__ __align__(128) float4 s_zero[kFill ? kProjZeroPage / 16 : 1];
  if (kFill) {
    carrier_init(s_zero, kProjZeroPage);
    __syncthreads();
    if (threadIdx.x == 0) carrier_issue(fill, s_zero, kProjZeroPage, blockIdx.x, gridDim.x, 0, 1);
  }
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
  if (kFill && threadIdx.x == 0) bulk_wait_read_all();
}


// ------------------------------------------------------------ fused, persistent --
// One launch for both passes, z-buffer resident in L2.
//
// The two-kernel form moves ~1.7x the algorithmic bytes through DRAM at batch 64: every scan's
// 1 MB of z-buffer keys is written back after the point pass, re-read by the resolve pass and
// written back again after its reset (24 B per pixel), and the winners' points are re-fetched.
// Here a persistent grid pulls work items from one queue in which the resolve items of scan s
// follow the point items of scan s + 2:
//     P(0) P(1) | R(0) P(2) | R(1) P(3) | ... | R(B-2) | R(B-1)
// and the z-buffer is a ring of kRing scan-sized regions (8 MB for KITTI) that never leaves L2.
// R(s) waits (spin on a counter) until every P(s) item has retired, P(s) until R(s - kRing) has
// reset the region it reuses; both are earlier in the queue, hence already running: no deadlock,
// no grid-wide barrier.  DRAM traffic ~= the algorithmic 28 N + 28 HW bytes.
constexpr int kRing = 8;            // z-buffer regions (scans in flight)
constexpr int kPtItem = 1024;       // points per P item (4 per thread)
constexpr int kPxItem = 1024;       // pixels per R item (4 per thread)

struct FusedCtl {                   // device control block (zeroed per call), after the ring
  unsigned int head;                // next queue item
  unsigned int pad[15];
  // followed by pdone[B], rdone[B]
};

__device__ __forceinline__ void wait_count(unsigned int* ctr, unsigned int need) {
  if (threadIdx.x == 0) {
    while (atomicAdd(ctr, 0u) < need) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

template <bool kHybrid, bool kAssemble>
__global__ void __launch_bounds__(256)
project_fused_kernel(const float4* __restrict__ points, const int32_t* __restrict__ offsets, int batch,
                     const float* __restrict__ depth_override, ProjParams p,
                     int32_t* __restrict__ upx, int32_t* __restrict__ upy, float* __restrict__ udepth,
                     unsigned long long* __restrict__ ring, unsigned int* __restrict__ ctl,
                     int32_t* __restrict__ flags,
                     // plain outputs
                     float* __restrict__ proj_range, float4* __restrict__ proj_pc,
                     int32_t* __restrict__ proj_idx, int32_t* __restrict__ proj_mask,
                     // assemble outputs (f-1)
                     const void* __restrict__ sem_label, const void* __restrict__ weak_label, int label_is_u8,
                     const float* __restrict__ mean, const float* __restrict__ stdv,
                     float* __restrict__ feature, long long* __restrict__ train_label,
                     long long* __restrict__ eval_label) {
  extern __shared__ int32_t s_tab[];            // [batch + 1] offsets, [batch + 3] block bases
  int32_t* s_off = s_tab;
  int32_t* s_base = s_tab + batch + 1;
  __shared__ unsigned int s_item;
  const int HW = p.H * p.W;
  const int nR = (HW + kPxItem - 1) / kPxItem;
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  // queue block j = R items of scan j - 2, then P items of scan j; exclusive prefix of the sizes
  if (threadIdx.x == 0) {
    int run = 0;
    for (int j = 0; j < batch + 2; ++j) {
      s_base[j] = run;
      if (j >= 2) run += nR;
      if (j < batch) run += (s_off[j + 1] - s_off[j] + kPtItem - 1) / kPtItem;
    }
    s_base[batch + 2] = run;
  }
  __syncthreads();
  const unsigned int n_items = (unsigned int)s_base[batch + 2];
  unsigned int* pdone = ctl + 16;
  unsigned int* rdone = pdone + batch;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(&ctl[0], 1u);
    __syncthreads();
    const unsigned int item = s_item;
    if (item >= n_items) break;
    int lo = 0, hi = batch + 2;                 // s_base[lo] <= item < s_base[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((unsigned)s_base[mid] <= item) lo = mid; else hi = mid; }
    const int j = lo;
    int k = (int)item - s_base[j];
    const bool is_r = (j >= 2) && (k < nR);
    if (!is_r) {
      // ---------------- P item: kPtItem points of scan j
      if (j >= 2) k -= nR;
      const int s = j;
      if (s >= kRing) wait_count(&rdone[s - kRing], (unsigned)nR);   // the region it reuses is reset
      unsigned long long* zb = ring + (size_t)(s % kRing) * HW;
      const int g0 = s_off[s] + k * kPtItem, g_end = s_off[s + 1];
      bool nan_any = false;
#pragma unroll
      for (int u = 0; u < kPtItem / 256; ++u) {
        const int g = g0 + u * 256 + threadIdx.x;
        if (g < g_end) {
          const float4 v = __ldg(points + g);
          const float depth = depth_override ? depth_override[g] : sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
          const float q = v.z / depth;
          int px, py; bool nan;
          pixel_of<kHybrid>(v.x, v.y, q, p, px, py, nan);
          upx[g] = px; upy[g] = py; udepth[g] = depth;
          const unsigned long long key = ((unsigned long long)depth_key(depth) << 32) | (uint32_t)(g - s_off[s]);
          atomicMin(zb + py * p.W + px, key);
          nan_any |= nan;
        }
      }
      if (nan_any) atomicOr(flags, 1);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(&pdone[s], 1u);
    } else {
      // ---------------- R item: kPxItem pixels of scan j - 2
      const int s = j - 2;
      const unsigned int nP = (unsigned)((s_off[s + 1] - s_off[s] + kPtItem - 1) / kPtItem);
      wait_count(&pdone[s], nP);
      unsigned long long* zb = ring + (size_t)(s % kRing) * HW;
      const int off = s_off[s];
      constexpr int U = kPxItem / 256;
      unsigned long long key[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = k * kPxItem + u * 256 + threadIdx.x;
        key[u] = (pix < HW) ? __ldcg(zb + pix) : ~0ull;       // L2: the keys were written by other SMs
      }
      float4 pt[U]; int sl[U], wl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pt[u] = make_float4(-1.f, -1.f, -1.f, -1.f); sl[u] = 0; wl[u] = 0;
        if (key[u] != ~0ull) {
          const size_t row = (size_t)off + (uint32_t)key[u];
          pt[u] = __ldg(points + row);
          if (kAssemble) {
            if (label_is_u8) {
              if (sem_label) sl[u] = __ldg(reinterpret_cast<const uint8_t*>(sem_label) + row);
              if (weak_label) wl[u] = __ldg(reinterpret_cast<const uint8_t*>(weak_label) + row);
            } else {
              if (sem_label) sl[u] = __ldg(reinterpret_cast<const int32_t*>(sem_label) + row);
              if (weak_label) wl[u] = __ldg(reinterpret_cast<const int32_t*>(weak_label) + row);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = k * kPxItem + u * 256 + threadIdx.x;
        if (pix >= HW) continue;
        const size_t q = (size_t)s * HW + pix;
        const bool valid = key[u] != ~0ull;
        const int idx = valid ? (int)(uint32_t)key[u] : -1;
        const float rng = valid ? key_depth((uint32_t)(key[u] >> 32)) : -1.0f;
        if (!kAssemble) {
          proj_range[q] = rng;
          proj_idx[q] = idx;
          proj_mask[q] = idx > 0;  // projection.py:113
          st_stream(proj_pc + q, pt[u]);
        } else {
          if (proj_range) proj_range[q] = rng;
          if (proj_idx) proj_idx[q] = idx;
          // loader :124-132 builds float32 label images, trainer :600-601 casts them to int64
          if (eval_label) eval_label[q] = (long long)(float)sl[u];
          if (train_label) train_label[q] = (long long)(float)wl[u];
          float f[5];
          f[0] = rng; f[1] = pt[u].x; f[2] = pt[u].y; f[3] = pt[u].z;
          f[4] = ((pt[u].w != -1.0f) ? 1.0f : 0.0f) * pt[u].w;          // loader :161-164
          float* dst = feature + (size_t)s * 5 * HW + pix;
          if (mean) {
            const float m = (sl[u] > 0) ? 1.0f : 0.0f;                   // eval_mask, trainer :603
#pragma unroll
            for (int c = 0; c < 5; ++c) f[c] = (f[c] - mean[c]) / stdv[c] * m;  // trainer :604-608
          }
#pragma unroll
          for (int c = 0; c < 5; ++c) dst[(size_t)c * HW] = f[c];
        }
        if (valid) zb[pix] = ~0ull;     // the ring region is left clean for scan s + kRing / the next call
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(&rdone[s], 1u);
    }
  }
}

// ---------------------------------------------------- cluster / DSMEM form ----
// The two-kernel form moves 1.7x its algorithmic bytes at batch 64: the z-buffer (1 MB per scan)
// does not stay in L2, so its lines are fetched by the atomics, read back by the resolve pass and
// written back reset.  Here a scan's z-buffer never leaves the chip: a thread-block CLUSTER of 8
// CTAs owns one scan at a time and keeps the scan's H x W keys in its distributed shared memory
// (8 x 128 KB for 64 x 2048 pixels).  Shared-memory atomics are native for 32 bits only (a
// 64-bit atomicMin through a generic pointer into a peer's shared memory LOSES updates --
// measured: 0.5 % of the pixels kept a farther point), so the 64-bit (depth, index) minimum is
// taken in two native 32-bit steps.  Per scan:
//   reset    every CTA resets its slice of depth keys and indices;               cluster barrier
//   depth    the cluster's 8192 threads stream the scan's points (4 per thread in flight), write
//            the per-point outputs and send one red.min.u32 of the depth key to the CTA that owns
//            the pixel (DSMEM);                                                  cluster barrier
//   index    the same threads re-read their per-point outputs (L2 hits), read the pixel's final
//            depth key from its owner and, where it is their own, red.min.u32 their point index
//            (equal depths: the smallest index wins, the z-buffer's tie rule);   cluster barrier
//   resolve  every CTA resolves its own slice (the winners' points were read a moment ago by this
//            cluster: L2 hits).
// Clusters walk the scans persistently.  DRAM traffic = algorithmic bytes + the sector waste of
// the winner gather; no global z-buffer, no memset, no reset pass.  Results are bit-identical to
// the two-kernel form (same keys, same minimum, same per-point arithmetic).
constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = 1024;
constexpr int kClusterPts = 4;           // points / pixels per thread in flight
constexpr size_t kClusterSmemMax = 200 * 1024;

__device__ __forceinline__ unsigned dsmem_addr(const void* local_smem, int owner) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem);
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(owner));
  return r;
}
__device__ __forceinline__ void dsmem_red_min(unsigned addr, unsigned v) {
  asm volatile("red.relaxed.cluster.shared::cluster.min.u32 [%0], %1;\n" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned dsmem_ld(unsigned addr) {
  unsigned v;
  asm volatile("ld.relaxed.cluster.shared::cluster.u32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

template <bool kHybrid>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
project_cluster_kernel(const float4* __restrict__ points, const int32_t* __restrict__ offsets, int batch,
                       const float* __restrict__ depth_override, ProjParams p, int slice,
                       int32_t* upx, int32_t* upy, float* udepth,
                       int32_t* __restrict__ flags, float* __restrict__ proj_range,
                       float4* __restrict__ proj_pc, int32_t* __restrict__ proj_idx,
                       int32_t* __restrict__ proj_mask) {
  extern __shared__ __align__(16) unsigned s_zd[];   // [slice] depth keys, then [slice] point indices
  unsigned* s_zi = s_zd + slice;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int n_clusters = gridDim.x / kClusterCtas, cid = blockIdx.x / kClusterCtas;
  const int HW = p.H * p.W;
  const int tid = threadIdx.x;
  constexpr int kStride = kClusterCtas * kClusterThreads;
  bool any_nan = false;
  for (int b = cid; b < batch; b += n_clusters) {
    for (int i = tid; i < 2 * slice; i += kClusterThreads) s_zd[i] = 0xFFFFFFFFu;
    cluster.sync();            // every slice is reset before the first atomic lands
    const int n0 = __ldg(offsets + b), n = __ldg(offsets + b + 1) - n0;
    // ---- depth: per-point outputs + minimum depth key per pixel
    for (int i0 = rank * kClusterThreads + tid; i0 < n; i0 += kClusterPts * kStride) {
      float4 v[kClusterPts];
      float dov[kClusterPts];
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        const int i = i0 + j * kStride;
        v[j] = make_float4(1.f, 0.f, 0.f, 0.f);
        dov[j] = 1.f;
        if (i < n) {
          v[j] = __ldg(points + n0 + i);
          if (depth_override) dov[j] = __ldg(depth_override + n0 + i);
        }
      }
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        const int i = i0 + j * kStride;
        if (i >= n) continue;
        const float x = v[j].x, y = v[j].y, z = v[j].z;
        const float depth = depth_override ? dov[j] : sqrtf((x * x + y * y) + z * z);
        const float q = z / depth;
        int px, py; bool nan;
        pixel_of<kHybrid>(x, y, q, p, px, py, nan);
        any_nan |= nan;
        const int g = n0 + i;
        upx[g] = px; upy[g] = py; udepth[g] = depth;        // re-read below: plain stores (L2)
        const int pix = py * p.W + px;
        const int owner = pix / slice, off = pix - owner * slice;
        dsmem_red_min(dsmem_addr(s_zd + off, owner), depth_key(depth));
      }
    }
    cluster.sync();            // the pixels' minimum depths are final
    // ---- index: among the points at a pixel's minimum depth, the smallest index
    for (int i0 = rank * kClusterThreads + tid; i0 < n; i0 += kClusterPts * kStride) {
      int pix[kClusterPts]; unsigned dk[kClusterPts], zmin[kClusterPts];
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        const int i = i0 + j * kStride;
        pix[j] = -1;
        if (i < n) {
          const int g = n0 + i;
          pix[j] = upy[g] * p.W + upx[g];
          dk[j] = depth_key(udepth[g]);
        }
      }
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        zmin[j] = 0;
        if (pix[j] >= 0) { const int owner = pix[j] / slice; zmin[j] = dsmem_ld(dsmem_addr(s_zd + (pix[j] - owner * slice), owner)); }
      }
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        if (pix[j] >= 0 && zmin[j] == dk[j]) {
          const int owner = pix[j] / slice;
          dsmem_red_min(dsmem_addr(s_zi + (pix[j] - owner * slice), owner), (unsigned)(i0 + j * kStride));
        }
      }
    }
    cluster.sync();            // winners are final
    // ---- resolve this CTA's slice
    const int pix0 = rank * slice;
    for (int l0 = tid; l0 < slice; l0 += kClusterPts * kClusterThreads) {
      unsigned kd[kClusterPts], ki[kClusterPts];
      float4 pt[kClusterPts];
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        const int l = l0 + j * kClusterThreads;
        const bool in = l < slice && pix0 + l < HW;
        kd[j] = in ? s_zd[l] : 0xFFFFFFFFu;
        ki[j] = in ? s_zi[l] : 0xFFFFFFFFu;
        pt[j] = make_float4(-1.f, -1.f, -1.f, -1.f);
        if (ki[j] != 0xFFFFFFFFu) pt[j] = __ldg(points + n0 + ki[j]);
      }
#pragma unroll
      for (int j = 0; j < kClusterPts; ++j) {
        const int l = l0 + j * kClusterThreads;
        if (l >= slice || pix0 + l >= HW) continue;
        const size_t qd = (size_t)b * HW + pix0 + l;
        const bool valid = ki[j] != 0xFFFFFFFFu;
        const int idx = valid ? (int)ki[j] : -1;
        __stcs(proj_range + qd, valid ? key_depth(kd[j]) : -1.0f);
        __stcs(proj_idx + qd, idx);
        __stcs(proj_mask + qd, (int)(idx > 0));   // projection.py:113
        st_stream(proj_pc + qd, pt[j]);
      }
    }
    // the reset at the top of the next scan touches only this CTA's own slice; the barrier after
    // it keeps the other CTAs' next atomics behind it
  }
  if (any_nan) atomicOr(flags, 1);
  cluster.sync();              // no CTA exits while a peer may still address its shared memory
}

static int cluster_slice(int HW) { return (HW + kClusterCtas - 1) / kClusterCtas; }
static bool cluster_form_fits(int HW) { return (size_t)cluster_slice(HW) * 8 <= kClusterSmemMax; }

static int launch_cluster(const float* points, const int32_t* offsets, int batch, const float* depth_override,
                          const ProjParams& p, bool hybrid, int32_t* upx, int32_t* upy, float* udepth,
                          int32_t* status_flags, float* proj_range, float* proj_pc, int32_t* proj_idx,
                          int32_t* proj_mask, cudaStream_t stream) {
  const int HW = p.H * p.W, slice = cluster_slice(HW);
  const size_t smem = (size_t)slice * sizeof(unsigned long long);
  auto kern = hybrid ? project_cluster_kernel<true> : project_cluster_kernel<false>;
  C3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // as many clusters as can be resident at once (a cluster lives inside one GPC), at most one per scan
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClusterCtas, 1, 1); cfg.blockDim = dim3(kClusterThreads, 1, 1);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kClusterCtas; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  C3D_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) { set_error("cluster projection: no 8-CTA cluster fits on this device"); return C3D_UNSUPPORTED; }
  const int n_clusters = batch < max_clusters ? batch : max_clusters;
  KernelTimer kt__("project_cluster_kernel", stream);
  kern<<<n_clusters * kClusterCtas, kClusterThreads, smem, stream>>>(
      reinterpret_cast<const float4*>(points), offsets, batch, depth_override, p, slice, upx, upy, udepth,
      status_flags, proj_range, reinterpret_cast<float4*>(proj_pc), proj_idx, proj_mask);
  return check_launch("project_cluster_kernel");
}

static size_t fused_ctl_bytes(int batch) { return (size_t)(16 + 2 * (size_t)batch) * 4; }
static size_t fused_ring_bytes(int batch, int HW) {
  return (size_t)(batch < kRing ? batch : kRing) * HW * sizeof(unsigned long long);
}


// Launches the fused persistent kernel (points must be [N, 4], 16 B aligned).
template <bool kAssemble>
static int launch_fused(const float* points, const int32_t* offsets, int batch, long long total_points,
                        const float* depth_override, const ProjParams& p, bool hybrid, int32_t* upx,
                        int32_t* upy, float* udepth, void* workspace, int32_t* status_flags,
                        float* proj_range, float* proj_pc, int32_t* proj_idx, int32_t* proj_mask,
                        const void* sem_label, const void* weak_label, int label_is_u8, const float* mean,
                        const float* stdv, float* feature, long long* train_label, long long* eval_label,
                        cudaStream_t stream) {
  const int HW = p.H * p.W;
  auto* ring = reinterpret_cast<unsigned long long*>(workspace);
  auto* ctl = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) +
                                              (size_t)batch * HW * sizeof(unsigned long long));
  C3D_CUDA(cudaMemsetAsync(ctl, 0, fused_ctl_bytes(batch), stream));
  const long long items = (long long)batch * ((HW + kPxItem - 1) / kPxItem) +
                          (total_points + kPtItem - 1) / kPtItem + batch;
  long long grid = (long long)kNumSMs * 4;   // 64 registers x 256 threads: four CTAs per SM
  if (grid > items) grid = items;
  const size_t smem = (size_t)(2 * batch + 4) * sizeof(int32_t);
  KernelTimer kt__("project_fused_kernel", stream);
#define C3D_FUSED(HY)                                                                               \
  project_fused_kernel<HY, kAssemble><<<(unsigned)grid, 256, smem, stream>>>(                        \
      reinterpret_cast<const float4*>(points), offsets, batch, depth_override, p, upx, upy, udepth,  \
      ring, ctl, status_flags, proj_range, reinterpret_cast<float4*>(proj_pc), proj_idx, proj_mask,  \
      sem_label, weak_label, label_is_u8, mean, stdv, feature, train_label, eval_label)
  if (hybrid) C3D_FUSED(true); else C3D_FUSED(false);
#undef C3D_FUSED
  return check_launch("project_fused_kernel");
}

}  // namespace c3d

using namespace c3d;

extern "C" int c3d_project_cluster_supported(int c_in, int proj_h, int proj_w) {
  // 1 if workspace_flags bit 3 of c3d_project_batch takes the cluster form for this shape
  return c_in == 4 && proj_h > 0 && proj_w > 0 && cluster_form_fits(proj_h * proj_w);
}

extern "C" size_t c3d_project_workspace_bytes(int batch, int proj_h, int proj_w) {
  if (batch <= 0 || proj_h <= 0 || proj_w <= 0) return 0;
  // two-kernel form: one key per pixel; fused form: a ring of kRing scans + the control block
  return (size_t)batch * proj_h * proj_w * sizeof(unsigned long long) + fused_ctl_bytes(batch);
}

extern "C" int c3d_project_batch(
    const float* points, int c_in, const int32_t* offsets, int batch, int64_t total_points,
    const float* depth_override, double abs_fov_left, double fov_hori, double abs_fov_down,
    double fov_vert, int proj_h, int proj_w, float* proj_range, float* proj_pointcloud,
    int32_t* proj_idx, int32_t* proj_mask, int32_t* uproj_x_idx, int32_t* uproj_y_idx,
    float* uproj_depth, void* workspace, int workspace_flags, int32_t* status_flags,
    void* cofill_ptr, size_t cofill_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d], got %d", kMaxBatch, batch);
  C3D_REQUIRE(c_in >= 3, "c_in must be >= 3, got %d", c_in);
  C3D_REQUIRE((cofill_ptr == nullptr) == (cofill_bytes == 0) && cofill_bytes % 16 == 0 &&
              (reinterpret_cast<uintptr_t>(cofill_ptr) & 15) == 0, "carried fill: 16 B aligned pointer and size");
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
  const bool c4 = (c_in == 4) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(proj_pointcloud) & 15) == 0);
  const bool hybrid = !(workspace_flags & 2);
  const bool fused = c4 && (workspace_flags & 4);    // bit 2: the fused persistent kernel (A/B; measured slower)
  // bit 3: the cluster / DSMEM form (z-buffer in distributed shared memory; workspace untouched)
  if ((workspace_flags & 8) && c4 && !fused && total_points > 0 && cluster_form_fits(proj_h * proj_w) &&
      ((reinterpret_cast<uintptr_t>(proj_pointcloud) & 15) == 0)) {
    if (cofill_bytes) { int rc = launch_fill(cofill_ptr, cofill_bytes, stream); if (rc) return rc; }
    return launch_cluster(points, offsets, batch, depth_override, p, hybrid, uproj_x_idx, uproj_y_idx, uproj_depth,
                          status_flags, proj_range, proj_pointcloud, proj_idx, proj_mask, stream);
  }
  if (!(workspace_flags & 1)) {
    KernelTimer kt__("zbuf_memset", stream);
    const size_t nb = fused ? fused_ring_bytes(batch, proj_h * proj_w) : (size_t)total_px * sizeof(unsigned long long);
    C3D_CUDA(cudaMemsetAsync(zbuf, 0xFF, nb, stream));
  }
  // the carried fill is split between the two passes; forms that cannot carry it fill up front
  FillShare f1{nullptr, 0}, f2{nullptr, 0};
  if (cofill_bytes) {
    if (fused || !c4 || total_points == 0) {
      int rc = launch_fill(cofill_ptr, cofill_bytes, stream);
      if (rc) return rc;
    } else {
      const unsigned long long half = (cofill_bytes / 2) & ~(unsigned long long)(kProjZeroPage - 1);
      f1 = FillShare{reinterpret_cast<char*>(cofill_ptr), half};
      f2 = FillShare{reinterpret_cast<char*>(cofill_ptr) + half, cofill_bytes - half};
    }
  }
  if (fused)
    return launch_fused<false>(points, offsets, batch, total_points, depth_override, p, hybrid, uproj_x_idx,
                               uproj_y_idx, uproj_depth, workspace, status_flags, proj_range, proj_pointcloud,
                               proj_idx, proj_mask, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                               nullptr, stream);
  const int threads = 256;
  if (total_points > 0) {
    int grid = (int)((total_points + threads - 1) / threads);  // short CTAs (overlap)
    size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
#define LAUNCH_PP(HY, C4, FL)                                                               \
  project_points_kernel<HY, C4, FL><<<grid, threads, smem, stream>>>(                       \
      points, c_in, offsets, batch, (int)total_points, depth_override, p, uproj_x_idx,      \
      uproj_y_idx, uproj_depth, zbuf, status_flags, f1)
    {
      KernelTimer kt__("project_points_kernel", stream);
      if (f1.bytes) { if (hybrid) LAUNCH_PP(true, true, true); else LAUNCH_PP(false, true, true); }
      else if (hybrid) { if (c4) LAUNCH_PP(true, true, false); else LAUNCH_PP(true, false, false); }
      else             { if (c4) LAUNCH_PP(false, true, false); else LAUNCH_PP(false, false, false); }
    }
#undef LAUNCH_PP
    int rc = check_launch("project_points_kernel");
    if (rc) return rc;
  }
  {
    int grid = (int)((total_px + 2 * threads - 1) / (2 * threads));
    KernelTimer kt__("resolve_pixels_kernel", stream);
    if (f2.bytes)
      resolve_pixels_kernel<true, true><<<grid, threads, 0, stream>>>(
          points, c_in, offsets, proj_h * proj_w, total_px, zbuf, proj_range, proj_pointcloud,
          proj_idx, proj_mask, f2);
    else if (c4)
      resolve_pixels_kernel<true, false><<<grid, threads, 0, stream>>>(
          points, c_in, offsets, proj_h * proj_w, total_px, zbuf, proj_range, proj_pointcloud,
          proj_idx, proj_mask, f2);
    else
      resolve_pixels_kernel<false, false><<<grid, threads, 0, stream>>>(
          points, c_in, offsets, proj_h * proj_w, total_px, zbuf, proj_range, proj_pointcloud,
          proj_idx, proj_mask, f2);
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
    int workspace_flags, int32_t* status_flags, void* cofill_ptr, size_t cofill_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d], got %d", kMaxBatch, batch);
  C3D_REQUIRE((cofill_ptr == nullptr) == (cofill_bytes == 0) && cofill_bytes % 16 == 0 &&
              (reinterpret_cast<uintptr_t>(cofill_ptr) & 15) == 0, "carried fill: 16 B aligned pointer and size");
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
  const bool hybrid = !(workspace_flags & 2);
  const bool fused = (workspace_flags & 4) != 0;
  if (!(workspace_flags & 1)) {
    KernelTimer kt__("zbuf_memset", stream);
    const size_t nb = fused ? fused_ring_bytes(batch, proj_h * proj_w) : (size_t)total_px * sizeof(unsigned long long);
    C3D_CUDA(cudaMemsetAsync(zbuf, 0xFF, nb, stream));
  }
  FillShare f1{nullptr, 0}, f2{nullptr, 0};
  if (cofill_bytes) {
    if (fused || total_points == 0) {
      int rc = launch_fill(cofill_ptr, cofill_bytes, stream);
      if (rc) return rc;
    } else {
      const unsigned long long half = (cofill_bytes / 2) & ~(unsigned long long)(kProjZeroPage - 1);
      f1 = FillShare{reinterpret_cast<char*>(cofill_ptr), half};
      f2 = FillShare{reinterpret_cast<char*>(cofill_ptr) + half, cofill_bytes - half};
    }
  }
  if (fused)
    return launch_fused<true>(points, offsets, batch, total_points, depth_override, p, hybrid, uproj_x_idx,
                              uproj_y_idx, uproj_depth, workspace, status_flags, proj_range, nullptr, proj_idx,
                              nullptr, sem_label, weak_label, label_is_u8, img_mean, img_std, feature,
                              (long long*)train_label, (long long*)eval_label, stream);
  const int threads = 256;
  if (total_points > 0) {
    int grid = (int)((total_points + threads - 1) / threads);
    size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
    KernelTimer kt__("project_points_kernel", stream);
#define LAUNCH_PA(HY, FL)                                                                              \
  project_points_kernel<HY, true, FL><<<grid, threads, smem, stream>>>(                                \
      points, 4, offsets, batch, (int)total_points, depth_override, p, uproj_x_idx, uproj_y_idx,       \
      uproj_depth, zbuf, status_flags, f1)
    if (f1.bytes) { if (hybrid) LAUNCH_PA(true, true); else LAUNCH_PA(false, true); }
    else          { if (hybrid) LAUNCH_PA(true, false); else LAUNCH_PA(false, false); }
#undef LAUNCH_PA
    int rc = check_launch("project_points_kernel");
    if (rc) return rc;
  }
  {
    int grid = (int)((total_px + 2 * threads - 1) / (2 * threads));
    KernelTimer kt__("resolve_assemble_kernel", stream);
    if (f2.bytes)
      resolve_assemble_kernel<true><<<grid, threads, 0, stream>>>(
          points, offsets, proj_h * proj_w, total_px, zbuf, sem_label, weak_label, label_is_u8, img_mean, img_std,
          proj_range, proj_idx, feature, (long long*)train_label, (long long*)eval_label, f2);
    else
      resolve_assemble_kernel<false><<<grid, threads, 0, stream>>>(
          points, offsets, proj_h * proj_w, total_px, zbuf, sem_label, weak_label, label_is_u8, img_mean, img_std,
          proj_range, proj_idx, feature, (long long*)train_label, (long long*)eval_label, f2);
    return check_launch("resolve_assemble_kernel");
  }
}
