// a3 -- EMA prototype update (prototype_learning + momentum_update + Sinkhorn).
//
// Replaces reference pc_processor/models/salsanext_proto.py:337-402 with its
// pre-step :497-510 and pc_processor/models/sinkhorn.py:5-33.  The reference
// evaluates LayerNorm, the (n x C*M) similarity and LayerNorm_C on all
// n = B*H*W pixels (838 MB of similarities at B=4) and then loops over classes
// in Python; only labelled pixels influence the update, so everything here runs
// on the labelled rows only:
//
//   E1 split_count / split_place (labelsplit.cuh): labelled pixels ->
//      rows sorted by (class, global pixel) = the reference's `label == id_c` order
//   E2 ema_rows      one warp per row: strided NCHW gather, LayerNorm_D, L2 norm,
//                    similarity against all C*M normalised prototypes (bank in
//                    shared memory), amax over M, LayerNorm_C, argmax -> mask;
//                    keeps the row's feature and its M similarities to its own class
//   E3 ema_sinkhorn  one CTA per class: 3 Sinkhorn iterations in the reference's
//                    operation order, argmax / Gumbel-hard assignment
//   E4 ema_segsum    one CTA per class: per-(class, sub-prototype) feature sums and
//                    counts, rows accumulated in order (no atomics) -> packed
//                    [K*D sums | K counts] buffer, the all-reduce payload
//   E5 ema_apply     normalise sums, EMA where count != 0, renormalise (:379-394)
#include <math_constants.h>
#include <string.h>

#include "common.cuh"
#include "labelsplit.cuh"
#include "proto_internal.cuh"
#include "rowgemm.cuh"

namespace c3d {

constexpr int kEmaWarps = 8;
constexpr int kMaxSub = 32;  // sub-prototypes per class handled in registers
enum EmaFlag { kEmaNoRows = 1, kEmaBadLabel = 8, kEmaOverflow = 16 };

struct EmaWs {
  SplitWs s;           // labelled-pixel slots, class-major (labelsplit.cuh)
  float* bank_n;       // [C*M*D]
  float* feat;         // [max_rows * D]
  float* simq;         // [max_rows * M]
  int32_t* maskv;      // [max_rows]
  int32_t* sub;        // [max_rows] assigned sub-prototype
  size_t bytes;
};

static EmaWs carve_ema(void* base, int B, int C, int HW, int D, int M, long long max_rows,
                       const SplitWs* shared = nullptr) {
  EmaWs w;
  size_t off = 0;
  if (shared) w.s = *shared; else w.s = carve_split(base, &off, B, C, HW);
  auto take = [&](size_t n) { size_t o = off; off += split_align(n); return (char*)base + o; };
  w.bank_n = (float*)take((size_t)C * M * D * 4);
  w.feat = (float*)take((size_t)max_rows * D * 4);
  w.simq = (float*)take((size_t)max_rows * M * 4);
  w.maskv = (int32_t*)take((size_t)max_rows * 4);
  w.sub = (int32_t*)take((size_t)max_rows * 4);
  w.bytes = off;
  return w;
}

// bank rows: l2_normalize (salsanext_proto.py:502), one warp per row
__global__ void __launch_bounds__(256)
ema_bank_normalise_kernel(const float* __restrict__ src_rows, int rows, int D, float* __restrict__ dst_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  if (k >= rows) return;
  const float* src = src_rows + (size_t)k * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) { float v = src[d]; s += v * v; }
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int d = lane; d < D; d += 32) dst_rows[(size_t)k * D + d] = src[d] * inv;
}

// ---------------------------------------------------------------- E2 -------
struct EmaRowsParams {
  const float* emb;      // (B, D, HW)
  const float* bank_n;   // (K, D)
  const float* ln_d_w; const float* ln_d_b; const float* ln_c_w; const float* ln_c_b;
  const int32_t* pix_list; const int32_t* cls_list;
  int32_t* info;
  float* feat; float* simq; int32_t* maskv;
  int HW, D, M, C, K, tile_rows, n_tiles, max_rows;
  float eps;
};

__global__ void __launch_bounds__(kEmaWarps * 32, 1)
ema_rows_kernel(EmaRowsParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, K = p.K, ld = D + 4, KPad = (K + 31) & ~31;
  float* s_bank = smem;
  float* s_a = s_bank + (size_t)p.tile_rows * ld;
  float* s_l = s_a + kEmaWarps * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n_rows = p.info[kInfoPl];
  if (n_rows > p.max_rows) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&p.info[kInfoFlags], kEmaOverflow);
    return;
  }
  const int n_groups = (n_rows + kEmaWarps - 1) / kEmaWarps;
  float* my_a = s_a + warp * D;
  float* my_l = s_l + warp * KPad;

  auto load_tile = [&](int tile) {
    const int r0 = tile * p.tile_rows;
    const int rows = min(p.tile_rows, K - r0);
    const int d4 = D >> 2;
    for (int i = threadIdx.x; i < rows * d4; i += blockDim.x) {
      const int r = i / d4, c = i - r * d4;
      cp_async16(s_bank + (size_t)r * ld + c * 4, p.bank_n + (size_t)(r0 + r) * D + c * 4);
    }
    cp_async_wait_all();
  };
  if (p.n_tiles == 1 && (int)blockIdx.x < n_groups) load_tile(0);
  __syncthreads();

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int slot = grp * kEmaWarps + warp;
    const bool active = slot < n_rows;
    int cls = 0;
    if (active) {
      const int gpix = p.pix_list[slot];
      cls = p.cls_list[slot];
      const int b = gpix / p.HW, pix = gpix - b * p.HW;
      const float* src = p.emb + (size_t)b * D * p.HW + pix;
      // LayerNorm over D (salsanext_proto.py:498), biased variance
      float s = 0.f;
      for (int d = lane; d < D; d += 32) { const float v = __ldg(src + (size_t)d * p.HW); my_a[d] = v; s += v; }
      const float mean = warp_sum(s) / (float)D;
      float v2 = 0.f;
      for (int d = lane; d < D; d += 32) { const float t = my_a[d] - mean; v2 += t * t; }
      const float rstd = 1.0f / sqrtf(warp_sum(v2) / (float)D + p.eps);
      float n2 = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float y = (my_a[d] - mean) * rstd * p.ln_d_w[d] + p.ln_d_b[d];
        my_a[d] = y; n2 += y * y;
      }
      const float inv = 1.0f / fmaxf(sqrtf(warp_sum(n2)), 1e-12f);  // l2_normalize (:501)
      for (int d = lane; d < D; d += 32) {
        const float y = my_a[d] * inv;
        my_a[d] = y;
        p.feat[(size_t)slot * D + d] = y;
      }
    }
    __syncwarp();
    // similarities against every (class, sub-prototype) (:504)
    for (int tile = 0; tile < p.n_tiles; ++tile) {
      if (p.n_tiles > 1) { __syncthreads(); load_tile(tile); __syncthreads(); }
      if (active) {
        const int r0 = tile * p.tile_rows;
        const int rows = min(p.tile_rows, K - r0);
        for (int kk = lane; kk < rows; kk += 32) {
          const float4* c4 = reinterpret_cast<const float4*>(s_bank + (size_t)kk * ld);
          const float4* a4 = reinterpret_cast<const float4*>(my_a);
          float acc = 0.f;
#pragma unroll 8
          for (int j = 0; j < (D >> 2); ++j) {
            const float4 c = c4[j], a = a4[j];
            acc += a.x * c.x; acc += a.y * c.y; acc += a.z * c.z; acc += a.w * c.w;
          }
          my_l[r0 + kk] = acc;
        }
      }
    }
    __syncwarp();
    if (active) {
      // nearest = amax over M (:506), LayerNorm over C (:507), argmax (:340)
      float nv[2]; float s = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        float m = -CUDART_INF_F;
        if (c < p.C) for (int j = 0; j < p.M; ++j) m = fmaxf(m, my_l[c * p.M + j]);
        nv[h] = m;
        if (c < p.C) s += m;
      }
      const float mean = warp_sum(s) / (float)p.C;
      float v2 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) if (lane + 32 * h < p.C) { const float t = nv[h] - mean; v2 += t * t; }
      const float rstd = 1.0f / sqrtf(warp_sum(v2) / (float)p.C + p.eps);
      float best = -CUDART_INF_F; int best_c = 0x7fffffff;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        if (c < p.C) {
          const float y = (nv[h] - mean) * rstd * p.ln_c_w[c] + p.ln_c_b[c];
          if (y > best) { best = y; best_c = c; }  // h ascending keeps the first maximum
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
      }
      if (lane == 0) p.maskv[slot] = (best_c == cls);  // mask = label == pred (:341)
      if (lane < p.M) p.simq[(size_t)slot * p.M + lane] = my_l[cls * p.M + lane];  // sim[..., id_c]
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- E2t ------
// The same rows computation, register-tiled (rowgemm.cuh): a CTA takes a batch of up to kEmaGroups
// groups of 16 rows, gathers + normalises them once, and then walks the bank tile by tile (tiles
// end on class boundaries), multiplying every staged tile with all groups of the batch: each
// bank element read from shared memory is reused for 4 rows, each staged tile for up to 64 rows.
// Per (row, class) only the maximum over the M sub-prototypes is kept, plus the row's M
// similarities to its own class -- the logits themselves live in a [16][tile] scratch.  Dot
// products accumulate in the same order as ema_rows_kernel (bitwise equal similarities).
constexpr int kEmaGroups = 4;
constexpr int kEmaCPT = 4;            // tile_logits columns per thread -> tiles of <= 256 bank rows
constexpr int kNvPad = kMaxClasses;   // per-row stride of the class-maximum scratch

constexpr int kEmaZeroPage = 8192;
struct EmaTiledPlan { int tile_classes, n_tiles, ldl; size_t smem; };
static int plan_ema_tiled(int D, int C, int M, EmaTiledPlan* out) {
  const size_t budget = 227 * 1024 - 2048;
  const BankLayout L = BankLayout::make(D);
  const size_t fixed = (size_t)kEmaGroups * kGroupRows * ((size_t)D + kNvPad + kMaxSub) * 4 + kEmaZeroPage;
  for (int n_tiles = 1; n_tiles <= C; ++n_tiles) {
    const int tc = (C + n_tiles - 1) / n_tiles;          // classes per tile
    const int rows = tc * M;
    if (rows > 64 * kEmaCPT) continue;
    const int ldl = (rows + 3) & ~3;
    const size_t need = fixed + (size_t)rows * L.ld * 4 + (size_t)kGroupRows * ldl * 4;
    if (need <= budget) { out->tile_classes = tc; out->n_tiles = (C + tc - 1) / tc; out->ldl = ldl; out->smem = need; return 0; }
  }
  return -1;
}

template <int kDJ>
__global__ void __launch_bounds__(kRowsThreads + 32, 1)
ema_rows_tiled_kernel(EmaRowsParams p, int tile_classes, int ldl, float* __restrict__ raw_rows, FillShare fill) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, M = p.M, C = p.C;
  const BankLayout BL = BankLayout::make(D);
  const int tile_cap = tile_classes * M;
  float* s_bank = smem;                                        // [tile_cap] rows, layout BL
  float* s_A = s_bank + (size_t)tile_cap * BL.ld;              // [G*16][D]
  float* s_L = s_A + (size_t)kEmaGroups * kGroupRows * D;      // [16][ldl]
  float* s_nv = s_L + (size_t)kGroupRows * ldl;                // [G*16][kNvPad] max over M per class
  float* s_own = s_nv + (size_t)kEmaGroups * kGroupRows * kNvPad;   // [G*16][kMaxSub]
  float* s_zero = s_own + (size_t)kEmaGroups * kGroupRows * kMaxSub; // zero page of the carried fill
  __shared__ int s_cls[kEmaGroups * kGroupRows];
  // carried fill: a ninth warp (launched only with a share) sends this CTA's slice on its own
  if (threadIdx.x >= kRowsThreads) {
    carrier_warp_run(fill, s_zero, kEmaZeroPage, blockIdx.x, gridDim.x);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = p.info[kInfoPl];
  if (n_rows > p.max_rows) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&p.info[kInfoFlags], kEmaOverflow);
    return;
  }
  const int n_groups = (n_rows + kGroupRows - 1) / kGroupRows;
  // groups per batch: as few as keeps every CTA busy once, at most kEmaGroups
  int G = (n_groups + (int)gridDim.x - 1) / (int)gridDim.x;
  G = G < 1 ? 1 : (G > kEmaGroups ? kEmaGroups : G);
  const int n_batches = (n_groups + G - 1) / G;
  const int n_tiles = (C + tile_classes - 1) / tile_classes;

  for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int row0 = batch * G * kGroupRows;
    // first bank tile on its way while the rows are gathered
    {
      const int rows0 = min(tile_classes, C) * M;
      stage_bank_tile_issue(s_bank, p.bank_n, 0, rows0, BL);
    }
    // ---- P0: gather, LayerNorm over D (salsanext_proto.py:498), L2 normalise (:501); two rows
    //      per warp at a time, their load chains interleaved
    for (int it = 0; it < G; ++it) {
      float areg[2][kDJ];
      int cls2[2]; bool act[2]; int rl2[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        rl2[rr] = it * kGroupRows + warp * 2 + rr;                // row within the batch
        const int slot = row0 + rl2[rr];
        act[rr] = slot < n_rows;
        int gpix = 0; cls2[rr] = 0;
        if (act[rr]) { gpix = p.pix_list[slot]; cls2[rr] = p.cls_list[slot]; }
        const int b = gpix / p.HW, pix = gpix - b * p.HW;
        const float* src = p.emb + (size_t)b * D * p.HW + pix;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) {
          const int d = lane + 32 * j;
          areg[rr][j] = (act[rr] && d < D) ? __ldg(src + (size_t)d * p.HW) : 0.f;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int slot = row0 + rl2[rr];
        if (raw_rows && act[rr]) {
#pragma unroll
          for (int j = 0; j < kDJ; ++j) { const int d = lane + 32 * j; if (d < D) raw_rows[(size_t)slot * D + d] = areg[rr][j]; }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) sum += areg[rr][j];          // d >= D lanes hold 0
        const float mean = warp_sum(sum) / (float)D;
        float v2 = 0.f;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) { const int d = lane + 32 * j; if (d < D) { const float t = areg[rr][j] - mean; v2 += t * t; } }
        const float rstd = 1.0f / sqrtf(warp_sum(v2) / (float)D + p.eps);
        float n2 = 0.f;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) {
          const int d = lane + 32 * j;
          if (d < D) { const float y = (areg[rr][j] - mean) * rstd * p.ln_d_w[d] + p.ln_d_b[d]; areg[rr][j] = y; n2 += y * y; }
        }
        const float inv = 1.0f / fmaxf(sqrtf(warp_sum(n2)), 1e-12f);
#pragma unroll
        for (int j = 0; j < kDJ; ++j) {
          const int d = lane + 32 * j;
          if (d < D) {
            const float y = act[rr] ? areg[rr][j] * inv : 0.f;
            s_A[(size_t)rl2[rr] * D + d] = y;
            if (act[rr]) p.feat[(size_t)slot * D + d] = y;
          }
        }
        if (lane == 0) s_cls[rl2[rr]] = act[rr] ? cls2[rr] : -1;
      }
    }
    // ---- tiles of the bank x groups of the batch
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int c0 = tile * tile_classes, tc = min(tile_classes, C - c0), rows = tc * M;
      if (tile > 0) { rows_sync(); stage_bank_tile_issue(s_bank, p.bank_n, c0 * M, rows, BL); }
      cp_async_wait_all();
      rows_sync();
      for (int g = 0; g < G; ++g) {
        float acc[4][kEmaCPT];
        tile_logits<kEmaCPT>(s_A + (size_t)g * kGroupRows * D, s_bank, rows, BL, acc);
        const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
        for (int i = 0; i < kEmaCPT; ++i) {
          const int c = cg + 64 * i;
          if (c < rows) {
#pragma unroll
            for (int r = 0; r < 4; ++r) s_L[(rg * 4 + r) * ldl + c] = acc[r][i];
          }
        }
        rows_sync();
        // (row, class of the tile): maximum over the M sub-prototypes (:506)
        for (int q = threadIdx.x; q < kGroupRows * tc; q += 256) {
          const int r = q / tc, cc = q - r * tc;
          const float* l = s_L + r * ldl + cc * M;
          float m = -CUDART_INF_F;
          for (int j = 0; j < M; ++j) m = fmaxf(m, l[j]);
          s_nv[(size_t)(g * kGroupRows + r) * kNvPad + c0 + cc] = m;
        }
        // the row's similarities to its own class (sim[..., id_c], :352)
        for (int q = threadIdx.x; q < kGroupRows * M; q += 256) {
          const int r = q / M, j = q - r * M;
          const int cls = s_cls[g * kGroupRows + r];
          if (cls >= c0 && cls < c0 + tc) s_own[(size_t)(g * kGroupRows + r) * kMaxSub + j] = s_L[r * ldl + (cls - c0) * M + j];
        }
        rows_sync();
      }
    }
    // ---- per row: LayerNorm over C (:507), argmax (:340), mask (:341)
    for (int rl = warp; rl < G * kGroupRows; rl += 8) {
      const int slot = row0 + rl;
      if (slot >= n_rows) continue;      // warp-uniform
      const int cls = s_cls[rl];
      float nv[2]; float sum = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        nv[h] = (c < C) ? s_nv[(size_t)rl * kNvPad + c] : 0.f;
        if (c < C) sum += nv[h];
      }
      const float mean = warp_sum(sum) / (float)C;
      float v2 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) if (lane + 32 * h < C) { const float t = nv[h] - mean; v2 += t * t; }
      const float rstd = 1.0f / sqrtf(warp_sum(v2) / (float)C + p.eps);
      float best = -CUDART_INF_F; int best_c = 0x7fffffff;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        if (c < C) {
          const float y = (nv[h] - mean) * rstd * p.ln_c_w[c] + p.ln_c_b[c];
          if (y > best) { best = y; best_c = c; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
      }
      if (lane == 0) p.maskv[slot] = (best_c == cls);
      if (lane < M) p.simq[(size_t)slot * M + lane] = s_own[(size_t)rl * kMaxSub + lane];
    }
    rows_sync();   // the batch's scratch is reused
  }
}

template <int kDJ>
static int launch_ema_tiled(const EmaRowsParams& p, const EmaTiledPlan& plan, float* raw_rows, FillShare fill,
                            cudaStream_t stream) {
  C3D_CUDA(cudaFuncSetAttribute(ema_rows_tiled_kernel<kDJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)plan.smem));
  KernelTimer kt__("ema_rows_kernel", stream);
  ema_rows_tiled_kernel<kDJ><<<kNumSMs, kRowsThreads + (fill.bytes ? 32 : 0), plan.smem, stream>>>(
      p, plan.tile_classes, plan.ldl, raw_rows, fill);
  return check_launch("ema_rows_tiled_kernel");
}

// ---------------------------------------------------------------- E2d ------
// Rows from the DENSE tensors the reference's forward has already built
// (salsanext_proto.py:497-510), for the drop-in `prototype_learning` (:337-402): out_feat
// (n, D) LayerNorm+L2-normalised rows, nearest (B, C, H, W), sim (n, M, C).  Only the
// labelled rows are read: feature row copy, pred = argmax_C nearest (:340), mask (:341) and the
// row's M similarities to its own class (:352-353).
__global__ void __launch_bounds__(256)
ema_rows_dense_kernel(const float* __restrict__ out_feat, const float* __restrict__ nearest,
                      const float* __restrict__ sim, const int32_t* __restrict__ pix_list,
                      const int32_t* __restrict__ cls_list, int32_t* __restrict__ info,
                      float* __restrict__ feat, float* __restrict__ simq, int32_t* __restrict__ maskv,
                      int HW, int D, int M, int C, int max_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = info[kInfoPl];
  if (n_rows > max_rows) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&info[kInfoFlags], kEmaOverflow);
    return;
  }
  for (int slot = blockIdx.x * 8 + warp; slot < n_rows; slot += gridDim.x * 8) {
    const int gpix = pix_list[slot], cls = cls_list[slot];
    const int b = gpix / HW, pix = gpix - b * HW;
    for (int d = lane; d < D; d += 32) feat[(size_t)slot * D + d] = __ldg(out_feat + (size_t)gpix * D + d);
    float best = -CUDART_INF_F; int best_c = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {   // ascending c keeps the first maximum per lane
      const float y = __ldg(nearest + ((size_t)b * C + c) * HW + pix);
      if (y > best) { best = y; best_c = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
    }
    if (lane == 0) maskv[slot] = (best_c == cls);
    if (lane < M) simq[(size_t)slot * M + lane] = __ldg(sim + ((size_t)gpix * M + lane) * C + cls);
  }
}

// ---------------------------------------------------------------- E3 -------
__device__ __forceinline__ uint4 philox_e(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

// One CTA per class.  Q (n_c x M) lives in shared memory when it fits (the
// weak-label regime: tens of rows per class) and in the global scratch otherwise;
// every element-wise step is element-parallel, row sums are (sub-prototype x 8
// partial) trees, column sums are one thread per row.  All sums have a fixed
// order, so the assignment is bitwise reproducible.
// mode: 0 = one_hot(argmax) (sinkhorn.py:30), 1 = injected Gumbel noise, 2 = device noise
constexpr int kSinkSmemFloats = 24 * 1024;  // 96 KB of dynamic shared memory

constexpr int kSinkWarps = 32;  // 1024 threads per class: the kernel is a chain of short
                                // element-parallel passes separated by barriers (latency bound)
__device__ __forceinline__ void sink_row_sums(const float* Q, int n, int M, float* s_part, float* s_R) {
  // s_R[m] = sum_i Q[i*M + m]; thread (p = warp, m = lane) sums rows p, p+kSinkWarps, ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a = 0.f;
  if (lane < M) for (int i = warp; i < n; i += kSinkWarps) a += Q[(size_t)i * M + lane];
  s_part[warp * 32 + lane] = a;
  __syncthreads();
  if ((int)threadIdx.x < M) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kSinkWarps; ++w) t += s_part[w * 32 + threadIdx.x];
    s_R[threadIdx.x] = t;
  }
  __syncthreads();
}

// General path of the per-class assignment (any n): called by all kSinkWarps*32 threads of the CTA.
__device__ void sinkhorn_general(float* s_dyn, float* s_part, float* s_R, float* s_tot_p, int c, int start, int n,
                                 const int32_t* __restrict__ pix_list, int M, float* __restrict__ simq,
                                 int32_t* __restrict__ sub, const float* __restrict__ gumbel, int mode,
                                 unsigned long long seed, float* __restrict__ proto_target) {
  float& s_tot = *s_tot_p;
  const int ne = n * M;
  const bool fits = ne + n <= kSinkSmemFloats;
  float* G = simq + (size_t)start * M;
  float* Q = fits ? s_dyn : G;
  float* csum = fits ? s_dyn + ne : reinterpret_cast<float*>(sub + start);  // n floats of scratch
  const float fM = (float)M, fn = (float)n;

  // Q = exp(out / eps) (sinkhorn.py:8)
  for (int e = threadIdx.x; e < ne; e += blockDim.x) Q[e] = expf(G[e] / 0.05f);
  __syncthreads();
  // sum_Q (:13): per-sub-prototype sums, then their sum
  sink_row_sums(Q, n, M, s_part, s_R);
  if (threadIdx.x == 0) { float t = 0.f; for (int m = 0; m < M; ++m) t += s_R[m]; s_tot = t; }
  __syncthreads();
  const float sum_q = s_tot;
  for (int e = threadIdx.x; e < ne; e += blockDim.x) Q[e] = Q[e] / sum_q;  // (:14)
  __syncthreads();
  for (int it = 0; it < 3; ++it) {  // sinkhorn.py:16-24
    sink_row_sums(Q, n, M, s_part, s_R);                                    // sum over rows (:18)
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
      float q = Q[e] / s_R[e % M];                                          // (:19)
      Q[e] = q / fM;                                                        // (:20)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {                     // column sums (:23)
      float t = 0.f;
      for (int m = 0; m < M; ++m) t += Q[(size_t)i * M + m];
      csum[i] = t;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
      float q = Q[e] / csum[e / M];                                         // (:23)
      Q[e] = q / fn;                                                        // (:24)
    }
    __syncthreads();
  }

  // Q *= B; argmax; assignment (sinkhorn.py:26-31)
  if (fits && 2 * ne + n <= kSinkSmemFloats) {
    // element-parallel noise (Philox + two logs per element is the expensive part; one thread
    // per ROW left all but n threads idle), then one thread per row takes the two arg-maxes
    float* Y = s_dyn + ne + n;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
      const float q = Q[e] * fn;
      Q[e] = q;
      float g = 0.f;
      const unsigned long long ctr = (unsigned long long)start * M + e;   // slot * M + m
      if (mode == 1) g = gumbel[ctr];
      else if (mode == 2) {
        const uint4 r = philox_e(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u),
                                 make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
        g = -logf(-logf(u));
      }
      Y[e] = (q + g) / 0.5f;  // F.gumbel_softmax(tau=0.5); softmax is monotone
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int slot = start + i;
      float best = -CUDART_INF_F, bestg = -CUDART_INF_F; int idx = 0, hard = 0;
      for (int m = 0; m < M; ++m) {
        const float q = Q[(size_t)i * M + m], y = Y[(size_t)i * M + m];
        if (q > best) { best = q; idx = m; }
        if (y > bestg) { bestg = y; hard = m; }
      }
      sub[slot] = (mode == 0) ? idx : hard;
      if (proto_target) proto_target[pix_list[slot]] = (float)idx + (float)(M * c);  // :390-392
    }
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int slot = start + i;
    float best = -CUDART_INF_F, bestg = -CUDART_INF_F; int idx = 0, hard = 0;
    for (int m = 0; m < M; ++m) {
      const float q = Q[(size_t)i * M + m] * fn;
      if (q > best) { best = q; idx = m; }
      float g = 0.f;
      if (mode == 1) g = gumbel[(size_t)slot * M + m];
      else if (mode == 2) {
        const unsigned long long ctr = (unsigned long long)slot * M + m;
        const uint4 r = philox_e(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u),
                                 make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
        g = -logf(-logf(u));
      }
      const float y = (q + g) / 0.5f;  // F.gumbel_softmax(tau=0.5); softmax is monotone
      if (y > bestg) { bestg = y; hard = m; }
    }
    sub[slot] = (mode == 0) ? idx : hard;
    if (proto_target) proto_target[pix_list[slot]] = (float)idx + (float)(M * c);  // :390-392
  }
}

// ---------------------------------------------------------------- E4 -------
// packed = [K*D sums | K counts] of one class.  The class's rows are dealt round-robin to
// `nsplit` warps, each accumulating its rows in order into a private shared-memory copy (lane
// owns feature columns); the copies are then added in warp order.  Fixed assignment + fixed
// order => atomics-free and bitwise reproducible.  Called by all threads of the CTA.
__device__ void segsum_class(float* s_sum, int c, int start, int n, int M, int D, int K, int nsplit,
                             const float* __restrict__ feat, const int32_t* __restrict__ maskv,
                             const int32_t* __restrict__ sub, float* __restrict__ packed) {
  const int stride = M * D + M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nsplit * stride; i += blockDim.x) s_sum[i] = 0.f;
  __syncthreads();
  if (warp < nsplit) {
    float* acc = s_sum + (size_t)warp * stride;
    // rows warp, warp + nsplit, ... in order; 32 rows at a time: lane j fetches the (mask, sub)
    // of the chunk's j-th row, so the feature loads below have no dependent address chain
    for (int i0 = warp; i0 < n; i0 += nsplit * 32) {
      const int mine = i0 + lane * nsplit;
      int code = -1;                                   // -1: masked out / beyond n
      if (mine < n && maskv[start + mine]) code = sub[start + mine];
      for (int j = 0; j < 32; ++j) {
        const int m = __shfl_sync(0xffffffffu, code, j);
        if (m < 0) continue;                           // m_q = q * mask, c_q = feat * mask (:363-375)
        const int slot = start + i0 + j * nsplit;
        for (int d = lane; d < D; d += 32) acc[m * D + d] += feat[(size_t)slot * D + d];
        if (lane == 0) acc[M * D + m] += 1.0f;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < stride; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nsplit; ++w) t += s_sum[(size_t)w * stride + i];
    if (i < M * D) packed[(size_t)c * M * D + i] = t;
    else packed[(size_t)K * D + c * M + (i - M * D)] = t;
  }
}

// One CTA per class: assignment (E3) and segmented sums (E4) in one launch.
__global__ void __launch_bounds__(kSinkWarps * 32)
ema_assign_sum_kernel(const int32_t* __restrict__ seg_cnt, const int32_t* __restrict__ seg_start,
                      const int32_t* __restrict__ pix_list, int32_t* __restrict__ info, int B, int M,
                      int D, int K, int ignore_label, int max_rows, int nsplit, float* __restrict__ simq,
                      int32_t* __restrict__ sub, const float* __restrict__ gumbel, int mode,
                      unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                      const float* __restrict__ feat,
                      const int32_t* __restrict__ maskv, float* __restrict__ packed,
                      float* __restrict__ proto_target) {
  extern __shared__ float s_dyn[];
  __shared__ float s_part[kSinkWarps * 32];
  __shared__ float s_R[32];
  __shared__ float s_tot;
  const int c = blockIdx.x;
  if (seed_dev) seed += seed_dev[1];   // device-side step counter: fresh noise in every graph replay
  int n = 0, start = 0;
  if (c != ignore_label && info[kInfoPl] <= max_rows) {
    start = seg_start[c * B];
    for (int b = 0; b < B; ++b) n += seg_cnt[c * B + b];
  }
  if (n > 0)  // else: no such class (:356-357), its sums and counts are zero
    sinkhorn_general(s_dyn, s_part, s_R, &s_tot, c, start, n, pix_list, M, simq, sub, gumbel, mode, seed,
                     proto_target);
  __syncthreads();   // sub[] of this class written (same CTA reads it below)
  segsum_class(s_dyn, c, start, n, M, D, K, nsplit, feat, maskv, sub, packed);
}

// ---------------------------------------------------------------- E5 -------
// One bank row (warp): normalise the summed features, EMA where the sub-prototype received rows,
// renormalise; `f` holds the row's summed features, lane-strided (f[j] = element lane + 32 j).
constexpr int kApplyDJ = 16;   // D <= 512
__device__ __forceinline__ void ema_apply_row(const float* protos_in, const float (&f)[kApplyDJ], bool update,
                                              int k, int D, float mom, float one_minus_mom, float* protos_out,
                                              float* __restrict__ normalised_out) {
  const int lane = threadIdx.x & 31;
  const float* old = protos_in + (size_t)k * D;
  float o2 = 0.f, f2 = 0.f;
#pragma unroll
  for (int j = 0; j < kApplyDJ; ++j) {
    const int d = lane + 32 * j;
    if (d < D) { o2 += old[d] * old[d]; f2 += f[j] * f[j]; }
  }
  const float oinv = 1.0f / fmaxf(sqrtf(warp_sum(o2)), 1e-12f);  // prototypes <- l2norm (:502)
  const float finv = 1.0f / fmaxf(sqrtf(warp_sum(f2)), 1e-12f);  // f = normalize(f) (:380)
  float v[kApplyDJ];
  float n2 = 0.f;
#pragma unroll
  for (int j = 0; j < kApplyDJ; ++j) {
    const int d = lane + 32 * j;
    v[j] = 0.f;
    if (d < D) {
      float t = old[d] * oinv;
      if (update) t = mom * t + one_minus_mom * (f[j] * finv);     // momentum_update (:19-31)
      v[j] = t;
      n2 += t * t;
    }
  }
  const float ninv = 1.0f / fmaxf(sqrtf(warp_sum(n2)), 1e-12f);    // l2_normalize(protos) (:394)
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < kApplyDJ; ++j) {
    const int d = lane + 32 * j;
    if (d < D) {
      v[j] = v[j] * ninv;
      protos_out[(size_t)k * D + d] = v[j];
      m2 += v[j] * v[j];
    }
  }
  if (normalised_out) {
    // F.normalize / l2_normalize of the stored bank, as its readers apply it: the loss on this
    // bank (contrast_pixel_loss.py:167) and the next step's similarity pre-step
    // (salsanext_proto.py:502) -- the same arithmetic as the stand-alone normalise kernels
    const float inv2 = 1.0f / fmaxf(sqrtf(warp_sum(m2)), 1e-12f);
#pragma unroll
    for (int j = 0; j < kApplyDJ; ++j) {
      const int d = lane + 32 * j;
      if (d < D) normalised_out[(size_t)k * D + d] = v[j] * inv2;
    }
  }
}

__global__ void __launch_bounds__(256)
ema_apply_kernel(const float* protos_in, const float* __restrict__ packed, int C, int M,
                 int D, int ignore_label, float mom, float one_minus_mom,
                 float* protos_out,   // may alias protos_in (row-local read-then-write)
                 float* __restrict__ normalised_out, unsigned long long* __restrict__ seed_counter) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = C * M;
  if (seed_counter && blockIdx.x == 0 && threadIdx.x == 0) seed_counter[1] += 1;   // next step: new noise
  const int k = blockIdx.x * 8 + warp;
  if (k >= K) return;
  const int c = k / M;
  const float* cnt = packed + (size_t)K * D;
  float csum = 0.f;
  for (int m = lane; m < M; m += 32) csum += cnt[c * M + m];
  csum = warp_sum(csum);
  const bool update = (c != ignore_label) && (csum > 0.f) && (cnt[k] != 0.f);  // :379,383
  float f[kApplyDJ];
#pragma unroll
  for (int j = 0; j < kApplyDJ; ++j) {
    const int d = lane + 32 * j;
    f[j] = d < D ? packed[(size_t)k * D + d] : 0.f;
  }
  ema_apply_row(protos_in, f, update, k, D, mom, one_minus_mom, protos_out, normalised_out);
}

// ------------------------------------------------------- E5 over peer memory -------
// The multi-GPU form of E5 WITHOUT a collective call: the all-reduce of the 206 kB payload and
// the EMA are one kernel, in the style of a low-latency ("LL") collective protocol.  Every rank
// owns a peer-visible exchange buffer of 2 (parities) x world slots x payload x {value, flag}
// pairs (c3d_peer_alloc, opened by the other ranks through CUDA IPC).  Per step s (a device
// counter, so captured graphs replay correctly), parity p = s & 1, flag value s + 1:
//   A  PUSH: all CTAs store their slice of the local payload as 8-byte {value, s + 1} pairs into
//      slot [p][rank] of EVERY rank's buffer (posted remote stores over NVLink: no round trip,
//      no fence, no separate flag -- an aligned 8-byte store arrives whole);
//   D  each warp polls, in its OWN buffer, the pairs of its bank row in all `world` slots until
//      their flags read s + 1 (bounded: a missing peer sets an error word instead of hanging the
//      GPU), sums the values in rank order 0..W-1 -- the same order on every rank, so the banks
//      stay bit-identical -- writes the sum back to the local payload and applies the EMA;
//   E  the last CTA to finish advances s.
// Slot [p][r] is rewritten by rank r at step s + 2, which r reaches only after its step s + 1
// consumed THIS rank's step s + 1 push, issued after this rank finished reading step s.
constexpr int kMaxPeers = 8;
struct PeerPtrs { float2* buf[kMaxPeers]; };

__device__ __forceinline__ float2 ld_pair(const float2* p) {
  float2 v;
  asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_pair(float2* p, float v, unsigned flag) {
  asm volatile("st.volatile.global.v2.f32 [%0], {%1, %2};\n" :: "l"(p), "f"(v), "f"(__uint_as_float(flag)) : "memory");
}

// values of payload element `i` from all ranks, summed in rank order (polls until every flag is there)
__device__ __forceinline__ float peer_sum(const float2* mine_par, size_t slot_pairs, size_t i, int world,
                                          unsigned flagv, long long t0, long long spin_limit, int32_t* err) {
  float2 v[kMaxPeers];
  unsigned pending = 0;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (r < world) { v[r] = ld_pair(mine_par + r * slot_pairs + i); if (__float_as_uint(v[r].y) != flagv) pending |= 1u << r; }
  while (pending) {
    if (clock64() - t0 > spin_limit) { atomicOr(err, (int)pending); break; }
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (pending & (1u << r)) {
        v[r] = ld_pair(mine_par + r * slot_pairs + i);
        if (__float_as_uint(v[r].y) == flagv) pending &= ~(1u << r);
      }
  }
  float t = 0.f;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) if (r < world) t += v[r].x;
  return t;
}

__global__ void __launch_bounds__(256)
ema_apply_peers_kernel(const float* protos_in, float* __restrict__ packed, PeerPtrs peers, int rank, int world,
                       size_t slot_pairs /* pairs per (parity, rank) slot, padded */, int32_t* __restrict__ state,
                       int C, int M, int D, int ignore_label, float mom, float one_minus_mom,
                       float* protos_out, float* __restrict__ normalised_out,
                       unsigned long long* __restrict__ seed_counter, long long spin_limit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = C * M;
  const size_t n_payload = (size_t)K * D + K;
  const unsigned step = *reinterpret_cast<volatile unsigned*>(&state[0]);
  const unsigned par = step & 1u, flagv = step + 1u;
  if (seed_counter && blockIdx.x == 0 && threadIdx.x == 0) seed_counter[1] += 1;
  const long long t0 = clock64();
  // A: push this CTA's rows (the ones it will reduce itself first, so that every rank's early
  //    CTAs feed every rank's early CTAs) and the counts
  const size_t slot_off = ((size_t)par * world + rank) * slot_pairs;
  {
    const size_t lo = (size_t)blockIdx.x * 8 * D, hi = min((size_t)(blockIdx.x + 1) * 8 * D, (size_t)K * D);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float v = packed[i];
      for (int r = 0; r < world; ++r) st_pair(peers.buf[r] + slot_off + i, v, flagv);
    }
    if (blockIdx.x == 0)
      for (size_t i = (size_t)K * D + threadIdx.x; i < n_payload; i += blockDim.x) {
        const float v = packed[i];
        for (int r = 0; r < world; ++r) st_pair(peers.buf[r] + slot_off + i, v, flagv);
      }
  }
  // the local payload is overwritten with the sums below: every thread of this CTA must have
  // read its part first (rows are CTA-private; the counts are read by CTA 0 only, written by all)
  __syncthreads();
  // D: sum over the ranks in rank order, write the sum back, apply the EMA
  const float2* mine_par = peers.buf[rank] + (size_t)par * world * slot_pairs;
  const int k = blockIdx.x * 8 + warp;
  if (k < K) {
    const int c = k / M;
    float csum = 0.f, cnt_k = 0.f;
    for (int m = lane; m < M; m += 32) {
      const float t = peer_sum(mine_par, slot_pairs, (size_t)K * D + c * M + m, world, flagv, t0, spin_limit, &state[3]);
      csum += t;
      if (c * M + m == k) cnt_k = t;
    }
    csum = warp_sum(csum);
    cnt_k = warp_sum(cnt_k);     // exactly one lane holds it, the others add zeros
    float f[kApplyDJ];
#pragma unroll
    for (int j = 0; j < kApplyDJ; ++j) {
      const int d = lane + 32 * j;
      f[j] = 0.f;
      if (d < D) f[j] = peer_sum(mine_par, slot_pairs, (size_t)k * D + d, world, flagv, t0, spin_limit, &state[3]);
    }
    const bool update = (c != ignore_label) && (csum > 0.f) && (cnt_k != 0.f);  // :379,383
    ema_apply_row(protos_in, f, update, k, D, mom, one_minus_mom, protos_out, normalised_out);
    // the summed payload back into `packed` (what an all-reduce would have left there); the counts
    // are written after EVERY CTA has pushed them (ticket below), the rows are CTA-private
#pragma unroll
    for (int j = 0; j < kApplyDJ; ++j) {
      const int d = lane + 32 * j;
      if (d < D) packed[(size_t)k * D + d] = f[j];
    }
    if (lane == 0) state[4 + k] = __float_as_int(cnt_k);
  }
  // E: the last CTA to finish (every CTA has pushed and reduced by then) writes the summed counts,
  //    re-arms the ticket and advances the step
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&state[1], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
      packed[(size_t)K * D + i] = __int_as_float(*reinterpret_cast<volatile int*>(&state[4 + i]));
    __syncthreads();
    if (threadIdx.x == 0) {
      state[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(&state[0]) = step + 1u;
    }
  }
}

static int ema_rows_config(int D, int K, int* tile_rows, int* n_tiles, size_t* smem) {
  const size_t budget = 227 * 1024;
  const size_t fixed = ((size_t)kEmaWarps * D + (size_t)kEmaWarps * ((K + 31) & ~31)) * 4;
  const size_t row = (size_t)(D + 4) * 4;
  if (fixed + 32 * row > budget) return -1;
  int tr = (int)((budget - fixed) / row);
  if (tr >= K) tr = K; else tr &= ~31;
  *tile_rows = tr;
  *n_tiles = (K + tr - 1) / tr;
  *smem = fixed + (size_t)tr * row;
  return 0;
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_proto_ema_workspace_bytes(int batch, int n_classes, int hw, int dim,
                                                int sub_protos, int64_t max_rows) {
  if (batch <= 0 || n_classes < 2 || hw <= 0 || dim <= 0 || sub_protos <= 0 || max_rows <= 0) return 0;
  return carve_ema(nullptr, batch, n_classes, hw, dim, sub_protos, max_rows).bytes;
}

size_t c3d::ema_extra_bytes(int C, int D, int M, long long max_rows) {
  SplitWs none{};
  return carve_ema(nullptr, 1, C, 1, D, M, max_rows, &none).bytes;
}

int c3d::proto_ema_accumulate_impl(
    const float* embedding, const DenseRows* dense, const int64_t* label, const float* prototypes,
    const float* ln_d_w, const float* ln_d_b, const float* ln_c_w, const float* ln_c_b, float ln_eps,
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos, int ignore_label,
    int64_t max_rows, const float* gumbel, int assign_mode, uint64_t seed, void* workspace,
    const SplitWs* shared_split, float* packed, float* proto_target, void* stream_, float* raw_rows,
    int rows_v1, const float* bank_n_in, const uint64_t* seed_dev, FillShare fill) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = batch, D = dim, C = n_classes, M = sub_protos;
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(B > 0 && B <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(C >= 2 && C <= kMaxClasses, "n_classes must be in [2, %d]", kMaxClasses);
  C3D_REQUIRE(M > 0 && M <= kMaxSub, "sub_protos must be in [1, %d]", kMaxSub);
  C3D_REQUIRE(D > 0 && D % 4 == 0 && D <= 1024, "feature dim must be a multiple of 4, <= 1024");
  C3D_REQUIRE(HWll > 0 && B * HWll < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(max_rows > 0 && max_rows <= B * HWll, "max_rows must be in [1, B*H*W]");
  C3D_REQUIRE(assign_mode >= 0 && assign_mode <= 2, "assign_mode must be 0, 1 or 2");
  C3D_REQUIRE(assign_mode != 1 || gumbel, "assign_mode 1 needs the gumbel noise");
  if (dense) {
    C3D_REQUIRE(dense->out_feat && dense->nearest && dense->sim && label && workspace && packed,
                "null pointer argument");
  } else {
    C3D_REQUIRE(embedding && (label || shared_split) && prototypes && ln_d_w && ln_d_b && ln_c_w &&
                ln_c_b && workspace && packed, "null pointer argument");
  }
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256 B aligned");
  const int HW = (int)HWll, K = C * M;
  EmaWs w = carve_ema(workspace, B, C, HW, D, M, max_rows, shared_split);
  int tile_rows = 0, n_tiles = 0; size_t smem = 0;
  C3D_REQUIRE(dense || ema_rows_config(D, K, &tile_rows, &n_tiles, &smem) == 0,
              "bank does not fit shared memory tiling (D=%d, K=%d)", D, K);
  int seg_split = 8;
  while (seg_split > 1 && (size_t)seg_split * ((size_t)M * D + M) * sizeof(float) > 200 * 1024) seg_split >>= 1;
  const size_t seg_smem = (size_t)seg_split * ((size_t)M * D + M) * sizeof(float);
  C3D_REQUIRE(seg_smem <= 227 * 1024, "M*D too large for the segmented-sum kernel");

  if (proto_target) C3D_CUDA(cudaMemsetAsync(proto_target, 0, (size_t)B * HW * 4, stream));
  int rc;
  if (!shared_split) {
    C3D_CUDA(cudaMemsetAsync(w.s.info, 0, (size_t)(8 + B) * 4, stream));
    if ((rc = launch_split((const long long*)label, nullptr, nullptr, B, C, HW, ignore_label, w.s, nullptr,
                           nullptr, nullptr, 0, stream))) return rc;
  }
  if (!dense && !bank_n_in) {
    KernelTimer kt__("bank_normalise_kernel", stream);
    ema_bank_normalise_kernel<<<(K + 7) / 8, 256, 0, stream>>>(prototypes, K, D, w.bank_n);
    if ((rc = check_launch("bank_normalise_kernel"))) return rc;
  }

  if (dense) {
    if (fill.bytes && (rc = launch_fill(fill.ptr, fill.bytes, stream))) return rc;
    KernelTimer kt__("ema_rows_dense_kernel", stream);
    ema_rows_dense_kernel<<<kNumSMs * 2, 256, 0, stream>>>(dense->out_feat, dense->nearest, dense->sim,
                                                           w.s.pix_list, w.s.cls_list, w.s.info, w.feat, w.simq,
                                                           w.maskv, HW, D, M, C, (int)max_rows);
    if ((rc = check_launch("ema_rows_dense_kernel"))) return rc;
  } else {
    EmaRowsParams p{};
    p.emb = embedding; p.bank_n = bank_n_in ? bank_n_in : w.bank_n; p.ln_d_w = ln_d_w; p.ln_d_b = ln_d_b;
    p.ln_c_w = ln_c_w; p.ln_c_b = ln_c_b; p.pix_list = w.s.pix_list; p.cls_list = w.s.cls_list;
    p.info = w.s.info; p.feat = w.feat; p.simq = w.simq; p.maskv = w.maskv;
    p.HW = HW; p.D = D; p.M = M; p.C = C; p.K = K; p.tile_rows = tile_rows; p.n_tiles = n_tiles;
    p.max_rows = (int)max_rows; p.eps = ln_eps;
    EmaTiledPlan plan;
    if (!rows_v1 && D % 32 == 0 && D <= 256 && plan_ema_tiled(D, C, M, &plan) == 0) {
      if (D <= 32) rc = launch_ema_tiled<1>(p, plan, raw_rows, fill, stream);
      else if (D <= 64) rc = launch_ema_tiled<2>(p, plan, raw_rows, fill, stream);
      else if (D <= 128) rc = launch_ema_tiled<4>(p, plan, raw_rows, fill, stream);
      else rc = launch_ema_tiled<8>(p, plan, raw_rows, fill, stream);
      if (rc) return rc;
    } else {
      C3D_REQUIRE(raw_rows == nullptr, "raw rows need the tiled EMA rows kernel (D %% 32 == 0, D <= 256)");
      if (fill.bytes && (rc = launch_fill(fill.ptr, fill.bytes, stream))) return rc;   // this form does not carry
      C3D_CUDA(cudaFuncSetAttribute(ema_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      { KernelTimer kt__("ema_rows_kernel", stream); ema_rows_kernel<<<kNumSMs, kEmaWarps * 32, smem, stream>>>(p); }
      if ((rc = check_launch("ema_rows_kernel"))) return rc;
    }
  }
  size_t dyn = (size_t)kSinkSmemFloats * 4;
  if (seg_smem > dyn) dyn = seg_smem;
  C3D_CUDA(cudaFuncSetAttribute(ema_assign_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  { KernelTimer kt__("ema_assign_sum_kernel", stream);
    ema_assign_sum_kernel<<<C, kSinkWarps * 32, dyn, stream>>>(
        w.s.seg_cnt, w.s.seg_start, w.s.pix_list, w.s.info, B, M, D, K, ignore_label, (int)max_rows, seg_split,
        w.simq, w.sub, gumbel, assign_mode, seed, reinterpret_cast<const unsigned long long*>(seed_dev),
        w.feat, w.maskv, packed, proto_target); }
  return check_launch("ema_assign_sum_kernel");
}

extern "C" int c3d_proto_ema_accumulate(
    const float* embedding, const int64_t* label, const float* prototypes, const float* ln_d_w,
    const float* ln_d_b, const float* ln_c_w, const float* ln_c_b, float ln_eps, int batch, int dim,
    int proj_h, int proj_w, int n_classes, int sub_protos, int ignore_label, int64_t max_rows,
    const float* gumbel, int assign_mode, uint64_t seed, void* workspace, float* packed,
    float* proto_target, void* stream) {
  return proto_ema_accumulate_impl(embedding, nullptr, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                                   ln_eps, batch, dim, proj_h, proj_w, n_classes, sub_protos,
                                   ignore_label, max_rows, gumbel, assign_mode, seed, workspace, nullptr,
                                   packed, proto_target, stream, nullptr, 0, nullptr, nullptr, FillShare{nullptr, 0});
}

extern "C" int c3d_proto_ema_accumulate_dense(
    const float* out_feat, const float* nearest, const int64_t* label, const float* feat_proto_sim,
    int batch, int dim, int proj_h, int proj_w, int n_classes, int sub_protos, int ignore_label,
    int64_t max_rows, const float* gumbel, int assign_mode, uint64_t seed, void* workspace,
    float* packed, float* proto_target, void* stream) {
  DenseRows d{out_feat, nearest, feat_proto_sim};
  return proto_ema_accumulate_impl(nullptr, &d, label, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f,
                                   batch, dim, proj_h, proj_w, n_classes, sub_protos, ignore_label,
                                   max_rows, gumbel, assign_mode, seed, workspace, nullptr, packed,
                                   proto_target, stream, nullptr, 0, nullptr, nullptr, FillShare{nullptr, 0});
}

extern "C" int c3d_proto_ema_apply(const float* prototypes_in, const float* packed, int n_classes,
                                   int sub_protos, int dim, int ignore_label, double momentum,
                                   float* prototypes_out, float* normalised_out, uint64_t* seed_counter,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(prototypes_in && packed && prototypes_out, "null pointer argument");
  C3D_REQUIRE(n_classes >= 2 && sub_protos > 0 && dim > 0 && dim <= 32 * kApplyDJ, "bad prototype shape (dim <= %d)", 32 * kApplyDJ);
  const int K = n_classes * sub_protos;
  // weak-scalar rounding of the reference: momentum and (1 - momentum) are Python
  // floats multiplied into float32 tensors (salsanext_proto.py:20)
  { KernelTimer kt__("ema_apply_kernel", stream); ema_apply_kernel<<<(K + 7) / 8, 256, 0, stream>>>(prototypes_in, packed, n_classes, sub_protos,
                                                    dim, ignore_label, (float)momentum,
                                                    (float)(1.0 - momentum), prototypes_out, normalised_out,
                                                    reinterpret_cast<unsigned long long*>(seed_counter)); }
  return check_launch("ema_apply_kernel");
}

// ---- peer exchange: buffer management (explicit, outside the hot path) and the fused kernel
static size_t peer_slot_pairs(int K, int D) { return (((size_t)K * D + K) + 63) & ~(size_t)63; }

extern "C" size_t c3d_peer_exchange_bytes(int n_classes, int sub_protos, int dim, int world) {
  if (n_classes < 2 || sub_protos <= 0 || dim <= 0 || world < 1 || world > kMaxPeers) return 0;
  return 2 * (size_t)world * peer_slot_pairs(n_classes * sub_protos, dim) * sizeof(float2);
}

extern "C" size_t c3d_peer_state_bytes(int n_classes, int sub_protos) {
  if (n_classes < 2 || sub_protos <= 0) return 0;
  return (size_t)(4 + n_classes * sub_protos) * sizeof(int32_t);
}

extern "C" int c3d_peer_alloc(size_t bytes, void** ptr) {
  C3D_REQUIRE(ptr && bytes > 0, "bad argument");
  C3D_CUDA(cudaMalloc(ptr, bytes));           // a dedicated allocation: IPC handles export whole allocations
  C3D_CUDA(cudaMemset(*ptr, 0, bytes));
  C3D_CUDA(cudaDeviceSynchronize());
  return C3D_OK;
}
extern "C" int c3d_peer_free(void* ptr) { C3D_CUDA(cudaFree(ptr)); return C3D_OK; }
extern "C" int c3d_peer_export(void* ptr, unsigned char* handle64) {
  C3D_REQUIRE(ptr && handle64, "null pointer argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  C3D_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return C3D_OK;
}
extern "C" int c3d_peer_import(const unsigned char* handle64, void** ptr) {
  C3D_REQUIRE(ptr && handle64, "null pointer argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  C3D_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return C3D_OK;
}
extern "C" int c3d_peer_close(void* ptr) { C3D_CUDA(cudaIpcCloseMemHandle(ptr)); return C3D_OK; }

extern "C" int c3d_proto_ema_apply_peers(const float* prototypes_in, float* packed, void* const* peer_bufs,
                                         int rank, int world, int32_t* state, int n_classes, int sub_protos,
                                         int dim, int ignore_label, double momentum, float* prototypes_out,
                                         float* normalised_out, uint64_t* seed_counter, double timeout_s,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(prototypes_in && packed && prototypes_out && peer_bufs && state, "null pointer argument");
  C3D_REQUIRE(n_classes >= 2 && sub_protos > 0 && dim > 0 && dim <= 32 * kApplyDJ, "bad prototype shape (dim <= %d)", 32 * kApplyDJ);
  C3D_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank / world out of range (<= %d ranks)", kMaxPeers);
  const int K = n_classes * sub_protos;
  PeerPtrs pp{};
  for (int r = 0; r < world; ++r) {
    C3D_REQUIRE(peer_bufs[r] != nullptr, "peer buffer %d is null", r);
    pp.buf[r] = reinterpret_cast<float2*>(peer_bufs[r]);
  }
  const long long spin = (long long)((timeout_s > 0 ? timeout_s : 2.0) * 1.9e9);   // SM clocks
  { KernelTimer kt__("ema_apply_peers_kernel", stream);
    ema_apply_peers_kernel<<<(K + 7) / 8, 256, 0, stream>>>(
        prototypes_in, packed, pp, rank, world, peer_slot_pairs(K, dim), state, n_classes, sub_protos, dim,
        ignore_label, (float)momentum, (float)(1.0 - momentum), prototypes_out, normalised_out,
        reinterpret_cast<unsigned long long*>(seed_counter), spin); }
  return check_launch("ema_apply_peers_kernel");
}

extern "C" int c3d_proto_bank_normalise(const float* prototypes, int rows, int dim, float* normalised_out,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(prototypes && normalised_out && rows > 0 && dim > 0, "bad argument");
  { KernelTimer kt__("bank_normalise_kernel", stream);
    ema_bank_normalise_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(prototypes, rows, dim, normalised_out); }
  return check_launch("bank_normalise_kernel");
}

extern "C" int c3d_proto_ema_info(const void* workspace, int32_t* host_info4, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(workspace && host_info4, "null pointer argument");
  C3D_CUDA(cudaMemcpyAsync(host_info4, workspace, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  C3D_CUDA(cudaStreamSynchronize(stream));
  return C3D_OK;
}
