// Stable multi-split of labelled pixels (shared by the prototype loss and the
// EMA prototype update).
//
//   split_count_scan  labels (+ keep_mask) -> per-tile per-class counts (9 B/px,
//                  coalesced); the last CTA of each scan turns them into exclusive
//                  prefixes, the last scan's CTA builds the segment table (start, index
//                  among non-empty) from the B*C totals -- one launch
//   split_scatter  labelled pixels -> slots sorted by (segment, pixel); optional
//                  entropy weight per slot; extra CTAs L2-normalise bank rows
//
// Segment order: kClassMajor = false -> (scan, class)  [ContrastMEMLoss X_ptr order,
// contrast_pixel_loss.py:82-109]; true -> (class, scan), i.e. per class all
// labelled pixels of the batch in global pixel order [prototype_learning's
// `label == id_c` row order, salsanext_proto.py:350-365].
// No atomics decide positions: the order is deterministic.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace c3d {

constexpr int kTile = 1024;  // pixels per count/scatter CTA (4 rounds x 256 threads)
constexpr int kMaxClasses = 64;
constexpr int kFlagBadLabel = 8;
enum SplitInfo { kInfoT = 0, kInfoPl = 1, kInfoFlags = 2, kInfoDone = 3, kInfoDone2 = 4 };

// Tile CTAs of the two split kernels.  Default: one CTA per tile.  With the concurrent hint
// (c3d_set_concurrent_hint, set by the step pipeline) and up to 2048 tiles (16 KITTI scans), a
// persistent grid of C3D_SPLIT_CTAS_PER_SM (default 2) CTAs per SM walks the tiles: alone it
// is slower (39 vs 18 us at batch 8), but next to the KNN vote, whose CTAs hold every SM slot
// for tens of microseconds, it is resident almost at once instead of being placed as slots
// trickle free (step 211 -> 203 us).  Larger batches keep one CTA per tile (the per-CTA tile
// loop costs more than the placement delay: 1436 vs 1252 us at batch 64).
inline int split_grid(int nblk) {
  if (!g_concurrent_hint.load(std::memory_order_relaxed) || nblk > 2048) return nblk;
  const char* env = getenv("C3D_SPLIT_CTAS_PER_SM");
  const int per_sm = env ? atoi(env) : 2;
  if (per_sm <= 0) return nblk;
  const int g = kNumSMs * per_sm;
  return nblk < g ? nblk : g;
}

template <bool kClassMajor>
__device__ __forceinline__ int seg_index(int b, int c, int B, int C) {
  return kClassMajor ? c * B + b : b * C + c;
}

__device__ __forceinline__ int masked_class(const long long* __restrict__ labels,
                                            const uint8_t* __restrict__ keep, size_t i,
                                            int ignore_label) {
  long long l = labels[i];
  if (keep && keep[i] == 0) l = ignore_label;  // contrast_pixel_loss.py:36-38
  return (int)l;
}

// Count + scan in one launch.  Every CTA counts its tile; the last CTA of a scan
// (atomic ticket per scan) turns the scan's tile counts into exclusive prefixes per
// class, and the last scan's CTA turns the B*C totals into the segment table.
// `info` is [8 + B] ints: info[8 + b] is scan b's ticket counter (zeroed by the caller).
template <bool kClassMajor>
__global__ void __launch_bounds__(256, 8)
split_count_scan_kernel(const long long* __restrict__ labels, const uint8_t* __restrict__ keep,
                        int HW, int nbps, int B, int C, int ignore_label,
                        int32_t* __restrict__ blk_cnt, int32_t* __restrict__ seg_cnt,
                        int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_tidx,
                        int32_t* __restrict__ info) {
  __shared__ int s_cnt[kMaxClasses];
  __shared__ int s_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Tiles are dealt to a small persistent grid: next to a kernel whose CTAs hold every SM
  // slot for tens of microseconds (the KNN vote), a grid of one CTA per tile is only placed as
  // slots trickle free, a grid of two CTAs per SM is resident almost at once.
  for (int blk = blockIdx.x; blk < B * nbps; blk += gridDim.x) {
  __syncthreads();
  if (threadIdx.x < C) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int b = blk / nbps, tile = blk % nbps;
  bool bad = false;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pix = tile * kTile + r * 256 + threadIdx.x;
    if (pix < HW) {
      const int c = masked_class(labels, keep, (size_t)b * HW + pix, ignore_label);
      if (c != ignore_label) {
        if (c < 0 || c >= C) bad = true; else atomicAdd(&s_cnt[c], 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < C) blk_cnt[(size_t)blk * C + threadIdx.x] = s_cnt[threadIdx.x];
  if (bad) atomicOr(&info[kInfoFlags], kFlagBadLabel);

  // ---- last CTA of scan b: exclusive prefix of the tile counts, per class
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_flag = (atomicAdd(&info[8 + b], 1) == nbps - 1);
  __syncthreads();
  if (!s_flag) continue;
  __threadfence();
  // All tile counts of up to four of this warp's classes are loaded before any is scanned:
  // one global round trip instead of one per (class, 32-tile chunk) -- this tail is latency.
  const int nch = (nbps + 31) >> 5;
  if (nch <= 4) {
    for (int c0 = warp; c0 < C; c0 += 32) {
      int v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 8 * k;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          v[k][q] = (c < C && i < nbps) ? __ldcg(blk_cnt + ((size_t)(b * nbps + i)) * C + c) : 0;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 8 * k;
        if (c >= C) break;
        int carry = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q >= nch) break;
          const int i = q * 32 + lane;
          int incl = v[k][q];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < nbps) blk_cnt[((size_t)(b * nbps + i)) * C + c] = carry + incl - v[k][q];
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) seg_cnt[seg_index<kClassMajor>(b, c, B, C)] = carry;
      }
    }
  } else {
  for (int c = warp; c < C; c += 8) {
    int carry = 0;
    for (int base = 0; base < nbps; base += 32) {
      const int i = base + lane;
      int32_t* p = blk_cnt + ((size_t)(b * nbps + i)) * C + c;
      const int v = (i < nbps) ? __ldcg(p) : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (i < nbps) *p = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) seg_cnt[seg_index<kClassMajor>(b, c, B, C)] = carry;
  }
  }
  // ---- last scan: segment table over the B*C totals
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { info[8 + b] = 0; s_flag = (atomicAdd(&info[kInfoDone], 1) == B - 1); }
  __syncthreads();
  if (!s_flag) continue;
  __threadfence();
  if (warp == 0) {
    int carry = 0, tcarry = 0;
    const int n = B * C;
    for (int base0 = 0; base0 < n; base0 += 32 * 8) {   // eight chunks' loads in flight
      int vv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = base0 + q * 32 + lane;
        vv[q] = (i < n) ? __ldcg(seg_cnt + i) : 0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int base = base0 + q * 32;
        if (base >= n) break;
        const int i = base + lane;
        const int v = vv[q];
        const int ne = v > 0;
        int incl = v, tincl = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, incl, o);
          int u = __shfl_up_sync(0xffffffffu, tincl, o);
          if (lane >= o) { incl += t; tincl += u; }
        }
        if (i < n) {
          seg_start[i] = carry + incl - v;
          seg_tidx[i] = ne ? (tcarry + tincl - 1) : -1;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
        tcarry += __shfl_sync(0xffffffffu, tincl, 31);
      }
    }
    if (lane == 0) {
      info[kInfoT] = tcarry;
      info[kInfoPl] = carry;
      if (tcarry == 0) atomicOr(&info[kInfoFlags], 1);  // no labelled pixel
      info[kInfoDone] = 0;
    }
  }
  }  // tile loop
}

template <bool kClassMajor, bool kEntropy>
__global__ void __launch_bounds__(256)
split_scatter_kernel(const long long* __restrict__ labels, const uint8_t* __restrict__ keep,
                     const float* __restrict__ probs, int HW, int nbps, int nblk, int B, int C,
                     int ignore_label, const int32_t* __restrict__ blk_prefix,
                     const int32_t* __restrict__ seg_start, int32_t* __restrict__ pix_list,
                     int32_t* __restrict__ cls_list, float* __restrict__ w_list,
                     int32_t* __restrict__ cnt_list, const float* __restrict__ bank_src,
                     int bank_rows, int D, float* __restrict__ bank_n, int tile_ctas) {
  if ((int)blockIdx.x >= tile_ctas) {
    // bank rows: F.normalize(x, p=2, dim=-1), eps 1e-12
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = bank_rows;
    for (int k = (blockIdx.x - tile_ctas) * 8 + warp; k < rows; k += (gridDim.x - tile_ctas) * 8) {
      const float* src = bank_src + (size_t)k * D;
      float s = 0.f;
      for (int d = lane; d < D; d += 32) { float v = src[d]; s += v * v; }
      s = warp_sum(s);
      const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
      for (int d = lane; d < D; d += 32) bank_n[(size_t)k * D + d] = src[d] * inv;
    }
    return;
  }
  __shared__ int s_cnt[4][8][kMaxClasses];  // [round][warp][class] -> exclusive prefix
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int blk = blockIdx.x; blk < nblk; blk += tile_ctas) {   // persistent grid, see split_count_scan
  const int b = blk / nbps, tile = blk % nbps;
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 8 * kMaxClasses; i += 256) (&s_cnt[0][0][0])[i] = 0;
  __syncthreads();
  int cls[4], rank[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pix = tile * kTile + r * 256 + threadIdx.x;
    int c = -1;
    if (pix < HW) {
      c = masked_class(labels, keep, (size_t)b * HW + pix, ignore_label);
      if (c == ignore_label || c < 0 || c >= C) c = -1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    rank[r] = __popc(peers & ((1u << lane) - 1));
    cls[r] = c;
    if (c >= 0 && rank[r] == 0) s_cnt[r][warp][c] = __popc(peers);
  }
  __syncthreads();
  if (threadIdx.x < C) {  // exclusive prefix over (round, warp) for class threadIdx.x
    int run = 0;
    for (int i = 0; i < 32; ++i) {
      int* p = &s_cnt[i >> 3][i & 7][threadIdx.x];
      const int v = *p; *p = run; run += v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = cls[r];
    if (c < 0) continue;
    const int pix = tile * kTile + r * 256 + threadIdx.x;
    const int slot = seg_start[seg_index<kClassMajor>(b, c, B, C)] +
                     blk_prefix[(size_t)blk * C + c] + s_cnt[r][warp][c] + rank[r];
    pix_list[slot] = b * HW + pix;
    cls_list[slot] = c;
    if (kEntropy) {
      // entropy weight (contrast_pixel_loss.py:46-49)
      const float* p = probs + (size_t)b * C * HW + pix;
      float ent = 0.f;
      for (int k0 = 0; k0 < C; k0 += 8) {  // 8 strided loads in flight per pass
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k0 + j < C) ? __ldg(p + (size_t)(k0 + j) * HW) : 1.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (k0 + j < C) ent += v[j] * logf(v[j] + 1e-10f);
      }
      ent = -ent;
      w_list[slot] = expf(-(ent * ent));
      cnt_list[slot] = 0;
    }
  }
  }  // tile loop
}


}  // namespace c3d
