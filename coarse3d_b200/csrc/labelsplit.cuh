// Stable multi-split of labelled pixels, shared by the prototype loss and the EMA prototype
// update (and computed ONCE when both run on the same labels: the fused prototype step).
//
//   split_count  labels (+ keep_mask), 9 B/px coalesced -> per-tile per-class counts and a
//                tile-local staging list of the labelled pixels {pixel, class, rank within
//                (tile, class)}; the last CTA of each scan turns the scan's tile counts into
//                exclusive prefixes per class, the last scan's CTA builds the segment tables
//   split_place  staged pixels -> slots sorted by (class, scan, pixel); optional entropy
//                weight per slot.  Only labelled pixels are touched (weak labels: ~1e-3 of all)
//
// One order serves both consumers.  Slots are class-major: per class all labelled pixels of the
// batch in global pixel order = prototype_learning's `label == id_c` row order
// (salsanext_proto.py:350-365); a (scan, class) segment -- ContrastMEMLoss' X_ptr unit,
// contrast_pixel_loss.py:82-109 -- is the contiguous run seg_start[c*B + b] .. + seg_cnt[c*B + b].
// The loss enumerates its segments in (scan, class) order through seg_tidx[b*C + c] (index among
// the non-empty segments, -1 if empty).  No atomics decide positions: the order is deterministic.
#pragma once
#include "common.cuh"

namespace c3d {

constexpr int kTile = 2048;        // pixels per count CTA (8 rounds x 256 threads)
constexpr int kTileRounds = kTile / 256;
constexpr int kMaxClasses = 64;
constexpr int kFlagBadLabel = 8;
enum SplitInfo { kInfoT = 0, kInfoPl = 1, kInfoFlags = 2, kInfoDone = 3, kInfoDone2 = 4 };

struct SplitWs {
  int32_t* info;       // [8 + B]: counters / flags + per-scan tickets (zeroed by the caller)
  int32_t* tile_tot;   // [nblk] labelled pixels of the tile
  int32_t* blk_cnt;    // [nblk * C] counts, then exclusive prefix inside (scan, class)
  int32_t* seg_cnt;    // [C * B] class-major
  int32_t* seg_start;  // [C * B] class-major
  int32_t* seg_tidx;   // [B * C] scan-major: index among non-empty segments, or -1
  int32_t* stage;      // [B * HW] tile-local records at the tile's own pixel range: pixel | class << 11 | rank << 17
  int32_t* pix_list;   // [B * HW] b*HW + pixel, sorted by (class, scan, pixel)
  int32_t* cls_list;   // [B * HW]
};

inline size_t split_align(size_t x) { return (x + 255) & ~(size_t)255; }
inline int split_tiles_per_scan(int HW) { return (HW + kTile - 1) / kTile; }

// Carves the split arrays from `base` at `*off` (advanced).  base may be null (size query).
inline SplitWs carve_split(void* base, size_t* off, int B, int C, int HW) {
  SplitWs w;
  const size_t cap = (size_t)B * HW, nblk = (size_t)B * split_tiles_per_scan(HW);
  auto take = [&](size_t n) { size_t o = *off; *off += split_align(n); return (char*)base + o; };
  w.info = (int32_t*)take((size_t)(8 + B) * 4);
  w.tile_tot = (int32_t*)take(nblk * 4);
  w.blk_cnt = (int32_t*)take(nblk * C * 4);
  w.seg_cnt = (int32_t*)take((size_t)B * C * 4);
  w.seg_start = (int32_t*)take((size_t)B * C * 4);
  w.seg_tidx = (int32_t*)take((size_t)B * C * 4);
  w.stage = (int32_t*)take(cap * 4);
  w.pix_list = (int32_t*)take(cap * 4);
  w.cls_list = (int32_t*)take(cap * 4);
  return w;
}

// Grid of the count kernel: one CTA per work item.
inline int split_grid(int nitems) { return nitems; }

__device__ __forceinline__ int masked_class(const long long* __restrict__ labels,
                                            const uint8_t* __restrict__ keep, size_t i,
                                            int ignore_label) {
  long long l = labels[i];
  if (keep && keep[i] == 0) l = ignore_label;  // contrast_pixel_loss.py:36-38
  return (int)l;
}

// Count + stage + scan in one launch.  A CTA handles `tpc` consecutive tiles of ONE scan (tpc
// divides the tiles per scan): the per-tile work is a handful of instructions when nothing is
// labelled, and the expensive part -- fence, ticket, barriers -- is paid once per CTA.
constexpr int kSplitZeroPage = 8192;
template <bool kFill>
static __global__ void __launch_bounds__(256)
split_count_kernel(const long long* __restrict__ labels, const uint8_t* __restrict__ keep, int HW,
                   int nbps, int tpc, int B, int C, int ignore_label, SplitWs w,
                   float* __restrict__ zero_buf, int zero_n, FillShare fill) {
  __shared__ int s_cnt[kTileRounds][8][kMaxClasses];  // [round][warp][class] -> exclusive prefix
  __shared__ int s_cpre[kMaxClasses + 1];             // exclusive prefix over classes (tile)
  __shared__ int s_flag;
  __shared__ __align__(128) float4 s_zero[kFill ? kSplitZeroPage / 16 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1;
  if (kFill) {   // this CTA's slice of the carried fill (common.cuh), issued while the labels stream in
    carrier_init(s_zero, kSplitZeroPage);
    __syncthreads();
    if (threadIdx.x == 0) carrier_issue(fill, s_zero, kSplitZeroPage, blockIdx.x, gridDim.x, 0, 1);
  }
  // a small caller buffer zeroed on the side (the packed prototype sums of the fused step)
  for (int i = blockIdx.x * 256 + threadIdx.x; i < zero_n; i += gridDim.x * 256) zero_buf[i] = 0.f;
  const int cps = nbps / tpc;                         // CTA work items per scan
  for (int item = blockIdx.x; item < B * cps; item += gridDim.x) {
    const int b = item / cps;
    // the labels of tile t + 1 are requested before tile t is processed, so that the memory
    // system stays busy across the barriers of the per-tile work
    int nxt[kTileRounds];
    auto fetch = [&](int tile) {
#pragma unroll
      for (int r = 0; r < kTileRounds; ++r) {
        const int pix = tile * kTile + r * 256 + threadIdx.x;
        nxt[r] = (pix < HW) ? masked_class(labels, keep, (size_t)b * HW + pix, ignore_label) : ignore_label;
      }
    };
    fetch((item % cps) * tpc);
    for (int tt = 0; tt < tpc; ++tt) {
    const int tile = (item % cps) * tpc + tt, blk = b * nbps + tile;
    int cls[kTileRounds], rank[kTileRounds];
    bool bad = false;
    unsigned any_round = 0;
#pragma unroll
    for (int r = 0; r < kTileRounds; ++r) {
      int c = nxt[r];
      if (c == ignore_label) c = -1;
      else if (c < 0 || c >= C) { bad = true; c = -1; }
      cls[r] = c;
    }
    if (tt + 1 < tpc) fetch(tile + 1);
#pragma unroll
    for (int r = 0; r < kTileRounds; ++r)
      if (__ballot_sync(0xffffffffu, cls[r] >= 0)) any_round |= 1u << r;
    const int tile_any = __syncthreads_or(any_round != 0);
    if (bad) atomicOr(&w.info[kInfoFlags], kFlagBadLabel);
    if (!tile_any) {
      // the common tile under weak labels: nothing labelled
      if (threadIdx.x < C) w.blk_cnt[(size_t)blk * C + threadIdx.x] = 0;
      if (threadIdx.x == 0) w.tile_tot[blk] = 0;
    } else {
      for (int i = threadIdx.x; i < kTileRounds * 8 * kMaxClasses; i += 256) (&s_cnt[0][0][0])[i] = 0;
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kTileRounds; ++r) {
        rank[r] = 0;
        if (any_round & (1u << r)) {   // warp-uniform
          const unsigned peers = __match_any_sync(0xffffffffu, cls[r]);
          rank[r] = __popc(peers & lt);
          if (cls[r] >= 0 && rank[r] == 0) s_cnt[r][warp][cls[r]] = __popc(peers);
        }
      }
      __syncthreads();
      if (threadIdx.x < C) {  // exclusive prefix over (round, warp) for class threadIdx.x
        int run = 0;
#pragma unroll 8
        for (int i = 0; i < kTileRounds * 8; ++i) {
          int* p = &s_cnt[i >> 3][i & 7][threadIdx.x];
          const int v = *p; *p = run; run += v;
        }
        w.blk_cnt[(size_t)blk * C + threadIdx.x] = run;
        s_cpre[threadIdx.x + 1] = run;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < C; ++c) { const int v = s_cpre[c + 1]; s_cpre[c] = run; run += v; }
        s_cpre[C] = run;
        w.tile_tot[blk] = run;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kTileRounds; ++r) {
        const int c = cls[r];
        if (c < 0) continue;
        const int rk = s_cnt[r][warp][c] + rank[r];          // rank within (tile, class), pixel order
        w.stage[(size_t)b * HW + (size_t)tile * kTile + s_cpre[c] + rk] = (r * 256 + (int)threadIdx.x) | (c << 11) | (rk << 17);
      }
    }
    }  // tiles of this CTA

    // ---- last CTA of scan b: exclusive prefix of the tile counts, per class
    __syncthreads();   // this CTA's counts are written; one thread publishes them (cumulative fence)
    if (threadIdx.x == 0) { __threadfence(); s_flag = (atomicAdd(&w.info[8 + b], 1) == cps - 1); }
    __syncthreads();
    if (!s_flag) continue;
    __threadfence();
    for (int c = warp; c < C; c += 8) {
      int carry = 0;
      for (int base = 0; base < nbps; base += 128) {   // four chunks' loads in flight
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = base + q * 32 + lane;
          v[q] = (i < nbps) ? __ldcg(w.blk_cnt + ((size_t)(b * nbps + i)) * C + c) : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = base + q * 32 + lane;
          if (base + q * 32 >= nbps) break;
          int incl = v[q];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < nbps) w.blk_cnt[((size_t)(b * nbps + i)) * C + c] = carry + incl - v[q];
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      if (lane == 0) w.seg_cnt[c * B + b] = carry;
    }
    // ---- last scan: segment tables over the B*C totals
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); w.info[8 + b] = 0; s_flag = (atomicAdd(&w.info[kInfoDone], 1) == B - 1); }
    __syncthreads();
    if (!s_flag) continue;
    __threadfence();
    const int n = B * C;
    if (warp == 0) {          // class-major starts: exclusive prefix of seg_cnt[c*B + b]
      int carry = 0;
      for (int base0 = 0; base0 < n; base0 += 256) {
        int vv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = base0 + q * 32 + lane;
          vv[q] = (i < n) ? __ldcg(w.seg_cnt + i) : 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int base = base0 + q * 32;
          if (base >= n) break;
          const int i = base + lane;
          int incl = vv[q];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < n) w.seg_start[i] = carry + incl - vv[q];
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      if (lane == 0) w.info[kInfoPl] = carry;
    } else if (warp == 1) {   // scan-major index among the non-empty segments
      int tcarry = 0;
      for (int base0 = 0; base0 < n; base0 += 256) {
        int ne[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = base0 + q * 32 + lane;   // i = b*C + c
          ne[q] = (i < n) ? (__ldcg(w.seg_cnt + (i % C) * B + i / C) > 0) : 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int base = base0 + q * 32;
          if (base >= n) break;
          const int i = base + lane;
          int incl = ne[q];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          if (i < n) w.seg_tidx[i] = ne[q] ? (tcarry + incl - 1) : -1;
          tcarry += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      if (lane == 0) {
        w.info[kInfoT] = tcarry;
        if (tcarry == 0) atomicOr(&w.info[kInfoFlags], 1);  // no labelled pixel
        w.info[kInfoDone] = 0;
      }
    }
  }  // work-item loop
  if (kFill && threadIdx.x == 0) bulk_wait_read_all();
}

// Staged pixels -> sorted slots (+ entropy weight, contrast_pixel_loss.py:46-49).  One WARP per
// tile (under weak labels a tile holds a couple of labelled pixels and the work is one dependent
// chain of loads); the records beyond the first 128 of a crowded tile are shared by the CTA.
template <bool kEntropy>
__device__ __forceinline__ void place_record(int rec, int b, int tile, int blk, const float* __restrict__ probs,
                                             int HW, int B, int C, const SplitWs& w,
                                             float* __restrict__ w_list, int32_t* __restrict__ cnt_list) {
  const int pl = rec & 2047, c = (rec >> 11) & 63, rk = rec >> 17;
  const int pix = tile * kTile + pl;
  const int slot = w.seg_start[c * B + b] + w.blk_cnt[(size_t)blk * C + c] + rk;
  w.pix_list[slot] = b * HW + pix;
  w.cls_list[slot] = c;
  if (kEntropy) {
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

template <bool kEntropy>
__global__ void __launch_bounds__(256)
split_place_kernel(const float* __restrict__ probs, int HW, int nbps, int nblk, int B, int C,
                   SplitWs w, float* __restrict__ w_list, int32_t* __restrict__ cnt_list) {
  constexpr int kWarpShare = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int s_n[8];
  for (int blk0 = blockIdx.x * 8; blk0 < nblk; blk0 += gridDim.x * 8) {
    const int blk = blk0 + warp;
    int n = 0;
    if (blk < nblk) n = w.tile_tot[blk];
    if (lane == 0) s_n[warp] = n;
    if (n > 0) {
      const int b = blk / nbps, tile = blk % nbps;
      const size_t base = (size_t)b * HW + (size_t)tile * kTile;
      for (int i = lane; i < min(n, kWarpShare); i += 32)
        place_record<kEntropy>(w.stage[base + i], b, tile, blk, probs, HW, B, C, w, w_list, cnt_list);
    }
    __syncthreads();
    for (int t = 0; t < 8; ++t) {           // crowded tiles (dense labels): the CTA shares the rest
      const int nt = s_n[t];
      if (nt <= kWarpShare) continue;
      const int bk = blk0 + t, b = bk / nbps, tile = bk % nbps;
      const size_t base = (size_t)b * HW + (size_t)tile * kTile;
      for (int i = kWarpShare + threadIdx.x; i < nt; i += 256)
        place_record<kEntropy>(w.stage[base + i], b, tile, bk, probs, HW, B, C, w, w_list, cnt_list);
    }
    __syncthreads();
  }
}

// Launches both kernels on `stream` (info must have been zeroed).  Returns a c3d status.
inline int launch_split(const long long* labels, const uint8_t* keep, const float* probs, int B, int C,
                        int HW, int ignore_label, const SplitWs& w, float* w_list, int32_t* cnt_list,
                        float* zero_buf, int zero_n, cudaStream_t stream, FillShare fill = FillShare{nullptr, 0}) {
  const int nbps = split_tiles_per_scan(HW), nblk = B * nbps;
  // tiles per count CTA: enough CTAs for ~4 per SM, a power of two that divides the tiles per scan
  int tpc = 1;
  while (tpc < 8 && nbps % (tpc * 2) == 0 && nblk / (tpc * 2) >= kNumSMs * 4) tpc *= 2;
  int rc;
  { KernelTimer kt__("split_count_kernel", stream);
    if (fill.bytes)
      split_count_kernel<true><<<split_grid(nblk / tpc), 256, 0, stream>>>(labels, keep, HW, nbps, tpc, B, C,
                                                                           ignore_label, w, zero_buf, zero_n, fill);
    else
      split_count_kernel<false><<<split_grid(nblk / tpc), 256, 0, stream>>>(labels, keep, HW, nbps, tpc, B, C,
                                                                            ignore_label, w, zero_buf, zero_n, fill); }
  if ((rc = check_launch("split_count_kernel"))) return rc;
  const int place_ctas = (nblk + 7) / 8;
  { KernelTimer kt__("split_place_kernel", stream);
    if (probs)
      split_place_kernel<true><<<place_ctas, 256, 0, stream>>>(probs, HW, nbps, nblk, B, C, w, w_list, cnt_list);
    else
      split_place_kernel<false><<<place_ctas, 256, 0, stream>>>(nullptr, HW, nbps, nblk, B, C, w, nullptr, nullptr); }
  return check_launch("split_place_kernel");
}

}  // namespace c3d
