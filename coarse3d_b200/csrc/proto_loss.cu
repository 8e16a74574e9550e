// a2 -- class-prototype contrastive loss (ContrastMEMLoss), forward + backward.
//
// Replaces reference pc_processor/loss/contrast_pixel_loss.py:27-195 and its
// autograd.  What the reference does with a full NCHW->NHWC copy, a B x C Python
// loop of unique / multinomial / gather launches, a 380 x D GEMM over A*T
// duplicated rows and an index_put backward is restated as:
//
//   K1-K2 split_count / split_place  (labelsplit.cuh) labelled pixels
//         -> slots sorted by (class, scan, pixel) + entropy weight exp(-H^2) per slot (:46-49)
//   K4 loss_sample   one CTA per (scan, class) segment: CDF, A draws with replacement
//         (Philox) or the injected `keep` indices -> multiplicity per slot; the
//         distinct sampled slots are compacted into ROWS (<= A per segment)
//   K5 loss_rows     one warp per distinct row: strided NCHW gather, L2 normalise,
//         similarity against the bank staged in shared memory (cp.async),
//         temperature, max-shift, masked sums (:166-193); rows are weighted by their
//         multiplicity instead of being duplicated.  With need_grad the same kernel
//         also evaluates the closed-form gradient row d loss / d feats[:, pixel]
//         (per unit upstream gradient) into a compact (rows x D) buffer, so the
//         backward pass has no arithmetic left on its critical path.
//   K6 fill_zero     dense (B,D,H,W) gradient zero fill, 128-bit streaming stores --
//         98 % of the path's compulsory bytes; deliberately low-occupancy so that it
//         can share the SMs with the latency-bound kernels of the other chains.
//   K7 loss_grad_scatter  D strided stores per distinct row, scaled by grad_out.
//
// All reductions are order-deterministic (no float atomics).
#include <math_constants.h>

#include "common.cuh"
#include "labelsplit.cuh"
#include "rowgemm.cuh"
#include "proto_internal.cuh"

namespace c3d {

constexpr int kRowWarps = 8;       // rows processed concurrently per loss_rows CTA

enum LossFlag { kFlagNoAnchor = 1, kFlagBadKeep = 2, kFlagKeepRows = 4, kFlagNoGradRows = 32 };
// info[] slots beyond SplitInfo
enum LossInfo { kInfoU = 5, kInfoDone3 = 6, kInfoHasGrad = 7 };

struct LossWs {
  SplitWs s;           // labelled-pixel slots, class-major (labelsplit.cuh)
  int32_t* seg_nd;     // [C * B] distinct sampled slots of the segment (class-major index)
  int32_t* row_base;   // [B * C + 1] exclusive prefix of seg_nd over non-empty segments, t order
  int32_t* seg_of_t;   // [B * C] class-major segment index of the t-th non-empty segment
  float* w_list;       // [cap] weights -> in-place CDF -> (int) distinct-slot list
  int32_t* cnt_list;   // [cap] sampling multiplicity
  float* loss_part;    // [max_rows]
  int32_t* row_pix;    // [max_rows] b*HW + pixel of each distinct row
  float* grad_rows;    // [max_rows * D]
  float* bank_n;       // [(C-1)*M*D] normalised prototypes, classes 1..C-1
  size_t max_rows;
  size_t bytes;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static LossWs carve(void* base, int B, int C, int HW, int D, int M, int A) {
  LossWs w;
  const size_t cap = (size_t)B * HW;
  size_t max_rows = (size_t)A * B * (C - 1);  // <= A distinct rows per segment
  if (max_rows > cap) max_rows = cap;
  size_t off = 0;
  w.s = carve_split(base, &off, B, C, HW);
  auto take = [&](size_t n) { size_t o = off; off += align_up(n); return (char*)base + o; };
  w.seg_nd = (int32_t*)take((size_t)B * C * 4);
  w.row_base = (int32_t*)take(((size_t)B * C + 1) * 4);
  w.seg_of_t = (int32_t*)take((size_t)B * C * 4);
  w.w_list = (float*)take(cap * 4);
  w.cnt_list = (int32_t*)take(cap * 4);
  w.loss_part = (float*)take(max_rows * 4);
  w.row_pix = (int32_t*)take(max_rows * 4);
  w.grad_rows = (float*)take(max_rows * D * 4);
  w.bank_n = (float*)take((size_t)(C - 1) * M * D * 4);
  w.max_rows = max_rows;
  w.bytes = off;
  return w;
}

// bank rows: F.normalize(x, p=2, dim=-1), eps 1e-12; one warp per row
__global__ void __launch_bounds__(256)
bank_normalise_kernel(const float* __restrict__ src_rows, int rows, int D, float* __restrict__ dst_rows) {
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

// ---------------------------------------------------------------- K4 -------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void __launch_bounds__(256)
loss_sample_kernel(const int32_t* __restrict__ seg_cnt, const int32_t* __restrict__ seg_start,
                   const int32_t* __restrict__ seg_tidx, const int32_t* __restrict__ pix_list,
                   float* __restrict__ w_list, int32_t* __restrict__ cnt_list, int HW, int B, int C, int A,
                   const long long* __restrict__ keep, int keep_rows, unsigned long long seed,
                   unsigned long long* __restrict__ seed_dev, int32_t* __restrict__ seg_nd, int32_t* __restrict__ row_base,
                   int32_t* __restrict__ seg_of_t, int32_t* __restrict__ info) {
  // one CTA per (class, scan) segment; seg = c*B + b (class-major storage), its rank among the
  // non-empty segments in the reference's (scan, class) order is seg_tidx[b*C + c]
  const int seg = blockIdx.x, nseg = gridDim.x;
  if (seed_dev) seed += seed_dev[0];   // device-side step counter (advanced by the last CTA below)
  const int n = seg_cnt[seg];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kCdfSmem = 2048;
  __shared__ float s_cdf[kCdfSmem];
  __shared__ int s_hit[kCdfSmem];   // multiplicities of a small segment (512 draws land on a
                                    // handful of slots: shared, not same-address global atomics)
  __shared__ float s_warp[8];
  __shared__ float s_carry;
  __shared__ int s_iw[8];
  __shared__ int s_last;
  if (keep && seg == 0 && threadIdx.x == 0 && info[kInfoT] != keep_rows)
    atomicOr(&info[kInfoFlags], kFlagKeepRows);
  if (n > 0) {
    const int b = seg % B, start = seg_start[seg], t = seg_tidx[b * C + seg / B];
    if (keep) {
      bool bad = false;
      for (int a = threadIdx.x; a < A && t < keep_rows; a += 256) {
        const long long pix = keep[(size_t)t * A + a];
        const long long key = (long long)b * HW + pix;
        int lo = 0, hi = n;  // first slot with pix_list >= key
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (pix_list[start + mid] < key) lo = mid + 1; else hi = mid;
        }
        if (pix < 0 || pix >= HW || lo >= n || pix_list[start + lo] != key) bad = true;
        else atomicAdd(&cnt_list[start + lo], 1);
      }
      if (bad) atomicOr(&info[kInfoFlags], kFlagBadKeep);
    } else {
      // inclusive scan of the weights -> CDF, 256 elements per step; small segments
      // (the weak-label regime) keep the CDF in shared memory for the binary searches
      float* cdf = w_list + start;
      const bool small = n <= kCdfSmem;
      if (small) {
        for (int i = threadIdx.x; i < n; i += 256) { s_cdf[i] = w_list[start + i]; s_hit[i] = 0; }
        cdf = s_cdf;
      }
      if (threadIdx.x == 0) s_carry = 0.f;
      __syncthreads();
      for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        float v = (i < n) ? cdf[i] : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float u = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += u;
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        float pre = s_carry;
        for (int w = 0; w < warp; ++w) pre += s_warp[w];
        if (i < n) cdf[i] = pre + v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = pre + v;
        __syncthreads();
      }
      const float total = s_carry;
      for (int a = threadIdx.x; a < A; a += 256) {
        const unsigned long long ctr = (unsigned long long)t * A + a;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float x = (float)(r.x >> 8) * (1.0f / 16777216.0f) * total;
        int lo = 0, hi = n - 1;  // first slot with cdf > x (clamped)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cdf[mid] > x) hi = mid; else lo = mid + 1;
        }
        if (small) atomicAdd(&s_hit[lo], 1); else atomicAdd(&cnt_list[start + lo], 1);
      }
      if (small) {   // cnt_list was zeroed by split_scatter
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 256) cnt_list[start + i] = s_hit[i];
      }
    }
    // ---- compact the distinct sampled slots of this segment, in slot order,
    //      into the (now dead) w_list storage of the segment
    __syncthreads();
    int32_t* dist = reinterpret_cast<int32_t*>(w_list) + start;
    int carry = 0;
    for (int base = 0; base < n; base += 256) {
      const int i = base + threadIdx.x;
      const bool f = (i < n) && (__ldcg(cnt_list + start + i) > 0);
      const unsigned bal = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_iw[warp] = __popc(bal);
      __syncthreads();
      int pre = carry, tot = 0;
      for (int w = 0; w < 8; ++w) { if (w < warp) pre += s_iw[w]; tot += s_iw[w]; }
      __syncthreads();
      if (f) dist[pre + __popc(bal & ((1u << lane) - 1))] = start + i;
      carry += tot;
    }
    if (threadIdx.x == 0) seg_nd[seg] = carry;
  } else if (threadIdx.x == 0) {
    seg_nd[seg] = 0;
  }
  // ---- last CTA: exclusive prefix of the row counts over non-empty segments
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&info[kInfoDone3], 1) == nseg - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < nseg; base += 32) {
      const int i = base + lane;                           // i = b*C + c, the reference's order
      const int t = (i < nseg) ? seg_tidx[i] : -1;
      const int sg = (i % C) * B + i / C;                  // class-major segment index
      const int v = (t >= 0) ? __ldcg(seg_nd + sg) : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (t >= 0) { row_base[t] = carry + incl - v; seg_of_t[t] = sg; }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      row_base[info[kInfoT]] = carry;
      info[kInfoU] = carry;
      info[kInfoDone3] = 0;
      if (seed_dev) seed_dev[0] += 1;   // every CTA has read it (they all passed the ticket)
    }
  }
}

// ---------------------------------------------------------------- K5 -------
struct RowsParams {
  const float* feats;      // (B, D, HW)
  const float* raw_rows;   // [slots, D] un-normalised feature rows gathered by the EMA kernel, or null
  int raw_cap;             // rows the EMA kernel had room for (more labelled pixels: it wrote none)
  FillShare fill;          // share of the carried zero fill (common.cuh), or {null, 0}
  const float* bank_n;     // (Kc, D)
  const int32_t* pix_list;
  const int32_t* cls_list;
  const int32_t* cnt_list;
  const int32_t* dist_list;  // distinct slots, per segment at seg_start
  const int32_t* seg_start;
  const int32_t* row_base;
  const int32_t* seg_of_t;
  int32_t* info;
  float* loss_part;        // [rows]
  int32_t* row_pix;        // [rows]
  float* grad_rows;        // [rows * D]
  float* loss_out;         // [1]
  int HW, D, M, Kc, A, tile_rows, n_tiles, ldl;
  float temperature, base_temperature;
};

// row index -> labelled-pixel slot, through the per-segment distinct lists
__device__ __forceinline__ int slot_of_row(int row, int T, const int32_t* __restrict__ row_base,
                                           const int32_t* __restrict__ seg_of_t,
                                           const int32_t* __restrict__ seg_start,
                                           const int32_t* __restrict__ dist_list,
                                           const int32_t* s_row_base = nullptr) {
  int lo = 0, hi = T;  // row_base[lo] <= row < row_base[hi]
  if (s_row_base) {    // copy in shared memory: 8 steps of ~30 cycles instead of L2 round trips
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_row_base[mid] <= row) lo = mid; else hi = mid;
    }
  } else {
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(row_base + mid) <= row) lo = mid; else hi = mid;
    }
  }
  const int seg = __ldg(seg_of_t + lo);
  return __ldcg(dist_list + __ldg(seg_start + seg) + (row - __ldg(row_base + lo)));
}

template <bool kWithGrad, int kChunks>
__global__ void __launch_bounds__(kRowWarps * 32, 1)
loss_rows_kernel(RowsParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, Kc = p.Kc, ld = D + 4;
  const int KcPad = (Kc + 31) & ~31;
  float* s_bank = smem;                                    // [tile_rows][D+4]
  float* s_a = s_bank + (size_t)p.tile_rows * ld;          // [warps][D]
  float* s_l = s_a + kRowWarps * D;                        // [warps][KcPad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = p.info[kInfoU];
  const int T = p.info[kInfoT];
  const int n_groups = (n_rows + kRowWarps - 1) / kRowWarps;
  float* my_a = s_a + warp * D;
  float* my_l = s_l + warp * KcPad;
  const float scale_row = p.temperature / p.base_temperature;
  const float inv_R = 1.0f / ((float)p.A * (float)T);  // mean over R = A*T rows (:193)

  // bank tile -> shared memory with cp.async (LDGSTS): every 16 B chunk is in
  // flight at once instead of one L2 round trip per loop iteration
  auto load_tile = [&](int tile) {
    const int r0 = tile * p.tile_rows;
    const int rows = min(p.tile_rows, Kc - r0);
    const int d4 = D >> 2;
    for (int i = threadIdx.x; i < rows * d4; i += blockDim.x) {
      const int r = i / d4, c = i - r * d4;
      cp_async16(s_bank + (size_t)r * ld + c * 4, p.bank_n + (size_t)(r0 + r) * D + c * 4);
    }
    cp_async_wait_all();
  };

  if (p.n_tiles == 1 && (int)blockIdx.x < n_groups) { load_tile(0); }
  __syncthreads();

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int row = grp * kRowWarps + warp;
    const bool active = row < n_rows;
    int cls = 0, cnt = 0;
    float inv_norm = 0.f;
    if (active) {
      const int slot = slot_of_row(row, T, p.row_base, p.seg_of_t, p.seg_start, p.dist_list);
      cnt = __ldcg(p.cnt_list + slot);
      const int gpix = p.pix_list[slot];
      cls = p.cls_list[slot];
      if (lane == 0) p.row_pix[row] = gpix;
      const int b = gpix / p.HW, pix = gpix - b * p.HW;
      const float* src = p.feats + (size_t)b * D * p.HW + pix;
      float n2 = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float v = __ldg(src + (size_t)d * p.HW);
        my_a[d] = v;
        n2 += v * v;
      }
      n2 = warp_sum(n2);
      inv_norm = 1.0f / fmaxf(sqrtf(n2), 1e-12f);  // F.normalize eps (:166)
      for (int d = lane; d < D; d += 32) my_a[d] *= inv_norm;
    }
    __syncwarp();

    // ---- logits z_k = (a_hat . c_hat_k) / temperature  (:168-172)
    for (int tile = 0; tile < p.n_tiles; ++tile) {
      if (p.n_tiles > 1) { __syncthreads(); load_tile(tile); __syncthreads(); }
      if (active) {
        const int r0 = tile * p.tile_rows;
        const int rows = min(p.tile_rows, Kc - r0);
        for (int kk = lane; kk < rows; kk += 32) {
          const float4* c4 = reinterpret_cast<const float4*>(s_bank + (size_t)kk * ld);
          const float4* a4 = reinterpret_cast<const float4*>(my_a);
          float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
          for (int j = 0; j < (D >> 2); ++j) {
            const float4 c = c4[j], a = a4[j];
            acc0 += a.x * c.x; acc1 += a.y * c.y; acc2 += a.z * c.z; acc3 += a.w * c.w;
          }
          my_l[r0 + kk] = ((acc0 + acc1) + (acc2 + acc3)) / p.temperature;
        }
      }
    }
    __syncwarp();

    if (active) {
      // ---- softmax statistics (:175-188)
      float mx = -CUDART_INF_F;
      for (int k = lane; k < Kc; k += 32) mx = fmaxf(mx, my_l[k]);
      mx = warp_max(mx);
      const int pos_lo = (cls - 1) * p.M, pos_hi = cls * p.M;  // bank rows of class `cls`
      float neg = 0.f;
      for (int k = lane; k < Kc; k += 32) {
        const float e = expf(my_l[k] - mx);
        if (k < pos_lo || k >= pos_hi) neg += e;
      }
      neg = warp_sum(neg);
      float s = 0.f, inv_den = 0.f; int npos = 0;
      for (int k = pos_lo + lane; k < pos_hi; k += 32) {
        if (k >= 0 && k < Kc) {
          const float l = my_l[k] - mx;
          const float den = expf(l) + neg + 1e-6f;
          s += l - logf(den);
          inv_den += 1.0f / den;
          ++npos;
        }
      }
      s = warp_sum(s);
      inv_den = warp_sum(inv_den);
      npos = __reduce_add_sync(0xffffffffu, npos);
      if (lane == 0)  // (:191-192), weighted by the row's multiplicity
        p.loss_part[row] = (float)cnt * (-scale_row * (s / (float)npos));
      if (kWithGrad) {
        // dL/dz_k: positives -s (1 - e_k/den_k); negatives s e_k sum_{j in pos} 1/den_j
        const float sg = scale_row / (float)npos;
        for (int k = lane; k < Kc; k += 32) {
          const float e = expf(my_l[k] - mx);
          float g;
          if (k >= pos_lo && k < pos_hi) g = -sg * (1.0f - e / (e + neg + 1e-6f));
          else g = sg * e * inv_den;
          my_l[k] = g / p.temperature;  // dL/d(a_hat . c_hat_k)
        }
      }
    }
    __syncwarp();

    if (kWithGrad) {
      // ---- d a_hat = sum_k g_k c_hat_k ; lane owns 4-wide chunks of D
      float4 acc[kChunks];
#pragma unroll
      for (int i = 0; i < kChunks; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int d4 = D >> 2;
      for (int tile = 0; tile < p.n_tiles; ++tile) {
        if (p.n_tiles > 1) { __syncthreads(); load_tile(tile); __syncthreads(); }
        if (active) {
          const int r0 = tile * p.tile_rows;
          const int rows = min(p.tile_rows, Kc - r0);
#pragma unroll 4
          for (int kk = 0; kk < rows; ++kk) {
            const float g = my_l[r0 + kk];
            const float4* c4 = reinterpret_cast<const float4*>(s_bank + (size_t)kk * ld);
#pragma unroll
            for (int i = 0; i < kChunks; ++i) {
              const int ch = lane + 32 * i;
              if (ch < d4) {
                const float4 c = c4[ch];
                acc[i].x += g * c.x; acc[i].y += g * c.y; acc[i].z += g * c.z; acc[i].w += g * c.w;
              }
            }
          }
        }
      }
      if (active) {
        // normalize backward: da = (dhat - a_hat (a_hat . dhat)) / max(|a|, eps)
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
          const int ch = lane + 32 * i;
          if (ch < d4) {
            const float4 a = reinterpret_cast<const float4*>(my_a)[ch];
            dot += a.x * acc[i].x + a.y * acc[i].y + a.z * acc[i].z + a.w * acc[i].w;
          }
        }
        dot = warp_sum(dot);
        const float w = (float)cnt * inv_R * inv_norm;
        const float sub = (inv_norm >= 1e12f) ? 0.f : dot;  // |a| < eps: y = x / eps
        float4* dst = reinterpret_cast<float4*>(p.grad_rows + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
          const int ch = lane + 32 * i;
          if (ch < d4) {
            const float4 a = reinterpret_cast<const float4*>(my_a)[ch];
            dst[ch] = make_float4((acc[i].x - a.x * sub) * w, (acc[i].y - a.y * sub) * w,
                                  (acc[i].z - a.z * sub) * w, (acc[i].w - a.w * sub) * w);
          }
        }
      }
    }
    __syncwarp();
  }

  // ---- deterministic final reduction by the last CTA: loss = sum / (A*T)
  __shared__ int s_last;
  __shared__ float s_red[kRowWarps];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.info[kInfoDone2], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float s = 0.f;
  for (int i = threadIdx.x; i < n_rows; i += blockDim.x) s += __ldcg(p.loss_part + i);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kRowWarps; ++w) tot += s_red[w];
    p.loss_out[0] = tot * inv_R;  // T == 0 -> NaN (the reference crashes)
    p.info[kInfoDone2] = 0;
    p.info[kInfoHasGrad] = kWithGrad ? 1 : 0;
  }
}

// Developer aid (compile with -DC3D_DEBUG_STAMPS): SM-clock stamps of CTA 0 at the phase
// boundaries of loss_rows16, read back by tools/rows_stamps.py.
#ifdef C3D_DEBUG_STAMPS
__device__ long long g_rows_dbg[16];
#define DBG_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_rows_dbg[i] = clock64(); } while (0)
#else
#define DBG_STAMP(i) do { } while (0)
#endif

// ---------------------------------------------------------------- K5b ------
// loss_rows16: the same math as loss_rows_kernel, organised as register-tiled products
// over groups of 16 rows (rowgemm.cuh).  Default; C3D_LOSS_ROWS_V1=1 selects the
// warp-per-row kernel above (kept for A/B measurements).
// kMma: the two products on the tensor cores (mma.sync m16n8k8, 3xTF32; rowgemm.cuh) instead of
// the register-tiled FFMA form.
template <bool kWithGrad, int kKS, int kRP, int kDch, int kDJ, bool kMma>
__global__ void __launch_bounds__(kRowsThreads + 32, 1)
loss_rows16_kernel(RowsParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, Kc = p.Kc, ldl = p.ldl;
  const BankLayout BL = BankLayout::make(D);
  float* s_bank = smem;                                   // [tile_rows] rows, layout BL
  const int lda = D + (kMma ? kMmaPadA : 0);
  float* s_A = s_bank + (size_t)p.tile_rows * BL.ld;      // [16][lda] a_hat, later d a_hat
  float* s_L = s_A + kGroupRows * lda;                    // [16][ldl] logits, later dL/dlogit
  __shared__ int s_cnt[kGroupRows], s_cls[kGroupRows];
  __shared__ float s_inv[kGroupRows];
  __shared__ int s_last;
  __shared__ float s_red[8];
  constexpr int kRbCap = 192;
  __shared__ int32_t s_rb[kRbCap];
  constexpr int kLossZeroPage = 2048;        // all the shared memory the bank tile leaves free
  __shared__ __align__(128) float4 s_zero[kLossZeroPage / 16];
  // carried fill: a ninth warp (launched only with a share) sends this CTA's slice on its own
  if (threadIdx.x >= kRowsThreads) {
    carrier_warp_run(p.fill, s_zero, kLossZeroPage, blockIdx.x, gridDim.x);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int n_rows = p.info[kInfoU];
  const int T = p.info[kInfoT];
  // row -> segment search: every `rb_stride`-th entry of row_base[0..T] lives in shared memory
  // (binary search there), the entries in between come from ONE round of independent loads
  const int rb_stride = (T + kRbCap) / kRbCap;          // 1 when the whole table fits
  const int rb_n = T / rb_stride + 1;                   // coarse entries 0, stride, 2 stride, ... <= T
  if ((int)threadIdx.x < rb_n) s_rb[threadIdx.x] = __ldg(p.row_base + threadIdx.x * rb_stride);
  const int n_groups = (n_rows + kGroupRows - 1) / kGroupRows;
  const float scale_row = p.temperature / p.base_temperature;
  const float inv_R = 1.0f / ((float)p.A * (float)T);  // mean over R = A*T rows (:193)
  const bool use_raw = p.raw_rows != nullptr && p.info[kInfoPl] <= p.raw_cap;

  DBG_STAMP(0);
  // single-tile case: start the bank copies now, complete them after the first gather
  bool staged = !(p.n_tiles == 1 && (int)blockIdx.x < n_groups);
  if (!staged) stage_bank_tile_issue(s_bank, p.bank_n, 0, Kc, BL);
  rows_sync();  // s_rb visible

  // The chain row -> segment -> distinct-slot list -> slot -> (count, pixel, class) is four
  // dependent global round trips before the features can be fetched.  It is walked ONE GROUP
  // AHEAD, one link per phase of the current group (each link's load is issued at a phase
  // boundary and first used a phase later), so that a group's P0 only waits for its features.
  bool n_act[2];
  int n_row[2], n_seg[2], n_base[2], n_slot[2], n_cnt[2], n_pix[2], n_cls[2];
  auto link_a = [&](int g2) {        // row -> segment index, segment id + first row of the segment
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = g2 * kGroupRows + warp * 2 + rr;
      n_row[rr] = row;
      n_act[rr] = g2 < n_groups && row < n_rows;
      n_seg[rr] = 0; n_base[rr] = 0;
      if (n_act[rr]) {
        int lo = 0, hi = rb_n;       // s_rb[lo] <= row < s_rb[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_rb[mid] <= row) lo = mid; else hi = mid; }
        lo *= rb_stride;
        if (rb_stride > 1) {         // row_base is non-decreasing: count the entries <= row
          int more = 0;
          for (int j = 1; j < rb_stride; ++j) {
            const int idx = lo + j;
            more += (idx < T && __ldg(p.row_base + idx) <= row) ? 1 : 0;
          }
          lo += more;
        }
        n_seg[rr] = __ldg(p.seg_of_t + lo);
        n_base[rr] = __ldg(p.row_base + lo);
      }
    }
  };
  auto link_b = [&]() {              // segment id -> start of its distinct-slot list
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) if (n_act[rr]) n_seg[rr] = __ldg(p.seg_start + n_seg[rr]);
  };
  auto link_c = [&]() {              // -> slot
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
      n_slot[rr] = n_act[rr] ? __ldcg(p.dist_list + n_seg[rr] + (n_row[rr] - n_base[rr])) : 0;
  };
  auto link_d = [&]() {              // -> multiplicity, pixel, class
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      n_cnt[rr] = n_act[rr] ? __ldcg(p.cnt_list + n_slot[rr]) : 0;
      n_pix[rr] = n_act[rr] ? p.pix_list[n_slot[rr]] : 0;
      n_cls[rr] = n_act[rr] ? p.cls_list[n_slot[rr]] : 0;
    }
  };
  link_a(blockIdx.x); link_b(); link_c(); link_d();     // the CTA's first group: un-overlapped

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    // ---- P0: gather + L2-normalise two rows per warp (:166)
    float areg[2][kDJ];
    int slot[2], cnt2[2], cls2[2], gpix2[2];
    bool act[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      act[rr] = n_act[rr]; slot[rr] = n_slot[rr]; cnt2[rr] = n_cnt[rr]; gpix2[rr] = n_pix[rr]; cls2[rr] = n_cls[rr];
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (use_raw) {      // fused step: the EMA kernel has left the row contiguous (512 B, coalesced)
        const float* src = p.raw_rows + (size_t)slot[rr] * D;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) {
          const int d = lane + 32 * j;
          areg[rr][j] = (act[rr] && d < D) ? __ldcg(src + d) : 0.f;
        }
      } else {
        const int b = gpix2[rr] / p.HW, pix = gpix2[rr] - b * p.HW;
        const float* src = p.feats + (size_t)b * D * p.HW + pix;
#pragma unroll
        for (int j = 0; j < kDJ; ++j) {
          const int d = lane + 32 * j;
          areg[rr][j] = (act[rr] && d < D) ? __ldg(src + (size_t)d * p.HW) : 0.f;
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int rl = warp * 2 + rr;
      float n2 = 0.f;
#pragma unroll
      for (int j = 0; j < kDJ; ++j) n2 += areg[rr][j] * areg[rr][j];
      n2 = warp_sum(n2);
      const float inv_norm = act[rr] ? 1.0f / fmaxf(sqrtf(n2), 1e-12f) : 0.f;
#pragma unroll
      for (int j = 0; j < kDJ; ++j) {
        areg[rr][j] *= inv_norm;
        const int d = lane + 32 * j;
        if (d < D) s_A[rl * lda + d] = areg[rr][j];
      }
      if (lane == 0) {
        s_cnt[rl] = cnt2[rr]; s_cls[rl] = cls2[rr]; s_inv[rl] = inv_norm;
        if (act[rr]) p.row_pix[grp * kGroupRows + rl] = gpix2[rr];
      }
    }
    DBG_STAMP(1);
    if (!staged) { cp_async_wait_all(); staged = true; }
    rows_sync();
    DBG_STAMP(2);
    link_a(grp + gridDim.x);

    // ---- P1: logits z = (a_hat . c_hat) / temperature (:168-172)
    for (int tile = 0; tile < p.n_tiles; ++tile) {
      const int r0 = tile * p.tile_rows, rows = min(p.tile_rows, Kc - r0);
      if (p.n_tiles > 1) { rows_sync(); stage_bank_tile(s_bank, p.bank_n, r0, rows, BL); rows_sync(); }
      if (kMma) {
        mma_tile_logits<kColsPerThread>(s_A, lda, s_bank, rows, BL, s_L, ldl, r0, p.temperature);
      } else {
        float acc[4][kColsPerThread];
        tile_logits(s_A, s_bank, rows, BL, acc);
#pragma unroll
        for (int i = 0; i < kColsPerThread; ++i) {
          const int c = cg + 64 * i;
          if (c < rows) {
#pragma unroll
            for (int r = 0; r < 4; ++r) s_L[(rg * 4 + r) * ldl + r0 + c] = acc[r][i] / p.temperature;
          }
        }
      }
    }
    rows_sync();

    DBG_STAMP(3);
    link_b();
    // ---- P2: softmax statistics, loss term, dL/dlogit (:175-193); one half-warp per
    //      row, so the two rows of a warp advance together
    {
      const int l16 = lane & 15, rl = warp * 2 + (lane >> 4), row = grp * kGroupRows + rl;
      float* my_l = s_L + rl * ldl;
      const int cnt = s_cnt[rl], cls = s_cls[rl];
      const bool live = row < n_rows;
      auto sum16 = [](float v) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
      };
      float mx = -CUDART_INF_F;
      if (live) {
#pragma unroll 4
        for (int k = l16; k < Kc; k += 16) mx = fmaxf(mx, my_l[k]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int pos_lo = (cls - 1) * p.M, pos_hi = cls * p.M;
      // shifted logits of this lane's (<= 2) positives stay in registers; every logit is
      // then replaced by exp(l), so exp is evaluated once per element
      const bool one_pass = p.M <= 32;
      float l_pos[2]; bool is_pos[2]; int kp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        kp[u] = pos_lo + l16 + 16 * u;
        is_pos[u] = live && one_pass && (l16 + 16 * u) < p.M && kp[u] >= 0 && kp[u] < Kc;
        l_pos[u] = is_pos[u] ? my_l[kp[u]] - mx : 0.f;
      }
      __syncwarp();
      float neg = 0.f;
      if (live) {
#pragma unroll 4
        for (int k = l16; k < Kc; k += 16) {
          const float e = expf(my_l[k] - mx);
          if (one_pass) my_l[k] = e;
          if (k < pos_lo || k >= pos_hi) neg += e;
        }
      }
      neg = sum16(neg);
      __syncwarp();
      float s = 0.f, inv_den = 0.f; float npos = 0.f;
      if (live) {
        if (one_pass) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (is_pos[u]) {
              const float den = my_l[kp[u]] + neg + 1e-6f;
              s += l_pos[u] - logf(den);
              inv_den += 1.0f / den;
              npos += 1.f;
            }
          }
        } else {
          for (int k = pos_lo + l16; k < pos_hi; k += 16) {
            if (k >= 0 && k < Kc) {
              const float l = my_l[k] - mx;
              const float den = expf(l) + neg + 1e-6f;
              s += l - logf(den);
              inv_den += 1.0f / den;
              npos += 1.f;
            }
          }
        }
      }
      s = sum16(s); inv_den = sum16(inv_den); npos = sum16(npos);
      if (live && l16 == 0) p.loss_part[row] = (float)cnt * (-scale_row * (s / npos));
      if (kWithGrad) {
        if (live) {
          const float sg = scale_row / npos;
#pragma unroll 4
          for (int k = l16; k < Kc; k += 16) {
            const float e = one_pass ? my_l[k] : expf(my_l[k] - mx);
            float g;
            if (k >= pos_lo && k < pos_hi) g = -sg * (1.0f - e / (e + neg + 1e-6f));
            else g = sg * e * inv_den;
            my_l[k] = g / p.temperature;
          }
        } else {
          for (int k = l16; k < Kc; k += 16) my_l[k] = 0.f;
        }
      }
    }
    rows_sync();

    DBG_STAMP(4);
    link_c();
    if (kWithGrad && kMma) {
      // ---- P3 (tensor cores): d a_hat = G . bank; warp w owns the feature slices 8 (w + 8 j)
      constexpr int kNTG = (kDJ + 1) / 2;
      float accm[kNTG][4];
#pragma unroll
      for (int j = 0; j < kNTG; ++j) accm[j][0] = accm[j][1] = accm[j][2] = accm[j][3] = 0.f;
      for (int tile = 0; tile < p.n_tiles; ++tile) {
        const int r0 = tile * p.tile_rows, rows = min(p.tile_rows, Kc - r0);
        if (p.n_tiles > 1) { rows_sync(); stage_bank_tile(s_bank, p.bank_n, r0, rows, BL); rows_sync(); }
        mma_tile_gradT<kNTG>(s_L, ldl, r0, s_bank, rows, BL, accm);
      }
      {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int j = 0; j < kNTG; ++j) {
          const int d0 = 8 * (warp + 8 * j) + 2 * t;
          if (d0 < D) {
            s_A[g * lda + d0] = accm[j][0]; s_A[g * lda + d0 + 1] = accm[j][1];
            s_A[(g + 8) * lda + d0] = accm[j][2]; s_A[(g + 8) * lda + d0 + 1] = accm[j][3];
          }
        }
      }
      rows_sync();
    }
    if (kWithGrad && !kMma) {
      // ---- P3: d a_hat = G . bank
      float4 acc4[kRP][kDch];
#pragma unroll
      for (int r = 0; r < kRP; ++r)
#pragma unroll
        for (int q = 0; q < kDch; ++q) acc4[r][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int tile = 0; tile < p.n_tiles; ++tile) {
        const int r0 = tile * p.tile_rows, rows = min(p.tile_rows, Kc - r0);
        if (p.n_tiles > 1) { rows_sync(); stage_bank_tile(s_bank, p.bank_n, r0, rows, BL); rows_sync(); }
        tile_gradT<kKS, kRP, kDch>(s_L, ldl, r0, s_bank, rows, BL, acc4);
      }
      {
        constexpr int kRowGroups = kGroupRows / kRP;
        constexpr int CT = 256 / (kKS * kRowGroups);
        const int ct = threadIdx.x % CT, rgp = (threadIdx.x / CT) % kRowGroups;
        const int kh = threadIdx.x / (CT * kRowGroups), d4 = D >> 2;
        if (kKS > 1) {
          // add the k-split partials through shared memory (G in s_L is dead now)
          rows_sync();
          float4* s_part = reinterpret_cast<float4*>(s_L);  // [kKS-1][16 rows][d4]
          if (kh > 0) {
#pragma unroll
            for (int q = 0; q < kDch; ++q) {
              const int ch = ct + CT * q;
              if (ch < d4) {
#pragma unroll
                for (int r = 0; r < kRP; ++r)
                  s_part[((kh - 1) * kGroupRows + rgp * kRP + r) * d4 + ch] = acc4[r][q];
              }
            }
          }
          rows_sync();
          if (kh == 0) {
#pragma unroll
            for (int q = 0; q < kDch; ++q) {
              const int ch = ct + CT * q;
              if (ch < d4) {
#pragma unroll
                for (int r = 0; r < kRP; ++r) {
                  for (int h = 1; h < kKS; ++h) {
                    const float4 v = s_part[((h - 1) * kGroupRows + rgp * kRP + r) * d4 + ch];
                    acc4[r][q].x += v.x; acc4[r][q].y += v.y; acc4[r][q].z += v.z; acc4[r][q].w += v.w;
                  }
                }
              }
            }
          }
        }
        if (kh == 0) {
#pragma unroll
          for (int q = 0; q < kDch; ++q) {
            const int ch = ct + CT * q;
            if (ch < d4) {
#pragma unroll
              for (int r = 0; r < kRP; ++r)
                *reinterpret_cast<float4*>(s_A + (rgp * kRP + r) * D + ch * 4) = acc4[r][q];
            }
          }
        }
      }
      rows_sync();
    }
    link_d();
    if (kWithGrad) {
      // ---- P4: normalize backward, weighted gradient row
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int rl = warp * 2 + rr, row = grp * kGroupRows + rl;
        if (row < n_rows) {
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < kDJ; ++j) {
            const int d = lane + 32 * j;
            if (d < D) dot += areg[rr][j] * s_A[rl * lda + d];
          }
          dot = warp_sum(dot);
          const float inv_norm = s_inv[rl];
          const float w = (float)s_cnt[rl] * inv_R * inv_norm;
          const float sub = (inv_norm >= 1e12f) ? 0.f : dot;  // |a| < eps: y = x / eps
#pragma unroll
          for (int j = 0; j < kDJ; ++j) {
            const int d = lane + 32 * j;
            if (d < D) p.grad_rows[(size_t)row * D + d] = (s_A[rl * lda + d] - areg[rr][j] * sub) * w;
          }
        }
      }
    }
    rows_sync();
  }

  DBG_STAMP(6);
  // ---- deterministic final reduction by the last CTA: loss = sum / (A*T)
  __threadfence();
  rows_sync();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.info[kInfoDone2], 1) == (int)gridDim.x - 1);
  rows_sync();
  DBG_STAMP(7);
  if (!s_last) return;
  __threadfence();
  float s = 0.f;
  for (int i = threadIdx.x; i < n_rows; i += kRowsThreads) s += __ldcg(p.loss_part + i);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  rows_sync();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    p.loss_out[0] = tot * inv_R;  // T == 0 -> NaN (the reference crashes)
    p.info[kInfoDone2] = 0;
    p.info[kInfoHasGrad] = kWithGrad ? 1 : 0;
  }
}

template <bool kWithGrad, int kKS, int kRP, int kDch, int kDJ, bool kMma = false>
static int launch_rows16(const RowsParams& p, size_t smem, cudaStream_t stream) {
  C3D_CUDA(cudaFuncSetAttribute(loss_rows16_kernel<kWithGrad, kKS, kRP, kDch, kDJ, kMma>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelTimer kt__("loss_rows_kernel", stream);
  loss_rows16_kernel<kWithGrad, kKS, kRP, kDch, kDJ, kMma>
      <<<kNumSMs, kRowsThreads + (p.fill.bytes ? 32 : 0), smem, stream>>>(p);
  return check_launch("loss_rows16_kernel");
}

// ---------------------------------------------------------------- K6 -------
// Streaming zero fill, two launch shapes (C3D_FILL_PERSISTENT=1 selects the second):
//  * short CTAs: each CTA writes one contiguous 32 KB block (256 threads x 8
//    independent 128-bit stores) and retires, so SM slots turn over within a
//    microsecond when higher-priority kernels are waiting;
//  * persistent: one 256-thread CTA per SM looping over 32 KB blocks -- a fixed small
//    footprint (4 k registers, no shared memory) that can co-reside with other kernels.
constexpr int kFillPerThread = 8;
__global__ void __launch_bounds__(256)
fill_zero_kernel(float4* __restrict__ dst, size_t n4, float* __restrict__ tail, int ntail) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t chunk = 256 * kFillPerThread;
  for (size_t blk = blockIdx.x; blk * chunk < n4; blk += gridDim.x) {
    const size_t base = blk * chunk + threadIdx.x;
    if (blk * chunk + chunk <= n4) {
#pragma unroll
      for (int j = 0; j < kFillPerThread; ++j) __stcs(dst + base + j * 256, z);
    } else {
      for (int j = 0; j < kFillPerThread; ++j)
        if (base + j * 256 < n4) __stcs(dst + base + j * 256, z);
    }
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

int launch_fill(void* dst, size_t nbytes, cudaStream_t stream) {
  const size_t n = nbytes / 4, n4 = n / 4;
  const size_t per_cta = 256 * kFillPerThread;
  size_t grid = (n4 + per_cta - 1) / per_cta;
  if (grid == 0) grid = 1;
  if (grid > 0x7fffffffull) { set_error("fill too large"); return C3D_INVALID_ARGUMENT; }
  KernelTimer kt__("fill_zero_kernel", stream);
  fill_zero_kernel<<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<float4*>(dst), n4,
                                                       reinterpret_cast<float*>(dst) + n4 * 4,
                                                       (int)(n - n4 * 4));
  return check_launch("fill_zero_kernel");
}

// ---------------------------------------------------------------- K7 -------
__global__ void __launch_bounds__(256)
loss_grad_scatter_kernel(const float* __restrict__ grad_rows, const int32_t* __restrict__ row_pix,
                         int32_t* __restrict__ info, const float* __restrict__ grad_out, int HW, int D,
                         float* __restrict__ grad_feats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = info[kInfoU];
  if (!info[kInfoHasGrad]) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&info[kInfoFlags], kFlagNoGradRows);
    return;
  }
  const float go = __ldg(grad_out);
  for (int row = blockIdx.x * 8 + warp; row < n_rows; row += gridDim.x * 8) {
    const int gpix = __ldcg(row_pix + row);
    const int b = gpix / HW, pix = gpix - b * HW;
    float* dst = grad_feats + (size_t)b * D * HW + pix;
    const float* src = grad_rows + (size_t)row * D;
    for (int d = lane; d < D; d += 32) dst[(size_t)d * HW] = __ldcg(src + d) * go;
  }
}

static int rows_config(int D, int Kc, int* tile_rows, int* n_tiles, size_t* smem) {
  const size_t budget = 227 * 1024;
  const size_t fixed = ((size_t)kRowWarps * D + (size_t)kRowWarps * ((Kc + 31) & ~31)) * 4;
  const size_t row = (size_t)(D + 4) * 4;
  if (fixed + 32 * row > budget) return -1;
  int tr = (int)((budget - fixed) / row);
  if (tr >= Kc) tr = Kc; else tr &= ~31;
  *tile_rows = tr;
  *n_tiles = (Kc + tr - 1) / tr;
  *smem = fixed + (size_t)tr * row;
  return 0;
}

template <bool kWithGrad, int kChunks>
static int launch_rows(const RowsParams& p, size_t smem, cudaStream_t stream) {
  C3D_CUDA(cudaFuncSetAttribute(loss_rows_kernel<kWithGrad, kChunks>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelTimer kt__("loss_rows_kernel", stream);
  loss_rows_kernel<kWithGrad, kChunks><<<kNumSMs, kRowWarps * 32, smem, stream>>>(p);
  return check_launch("loss_rows_kernel");
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_proto_loss_workspace_bytes(int batch, int n_classes, int hw, int dim,
                                                 int sub_protos, int num_anchor) {
  if (batch <= 0 || n_classes < 2 || hw <= 0 || dim <= 0 || sub_protos <= 0 || num_anchor <= 0)
    return 0;
  return carve(nullptr, batch, n_classes, hw, dim, sub_protos, num_anchor).bytes;
}

namespace c3d {
size_t loss_ws_bytes(int B, int C, int HW, int D, int M, int A) { return carve(nullptr, B, C, HW, D, M, A).bytes; }
SplitWs loss_ws_split(void* base, int B, int C, int HW, int D, int M, int A) { return carve(base, B, C, HW, D, M, A).s; }
}  // namespace c3d
using namespace c3d;

// phases (internal bit mask): kPhaseSplit = label split, kPhaseSample = anchor sampling,
// kPhaseRows = loss and gradient rows.  zero_buf / zero_n: a small buffer the split zeroes
// on the side (the fused step's packed prototype sums).
int c3d::proto_loss_forward_impl(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* proto_queue, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, uint64_t seed, int need_grad, int phases, void* workspace,
    float* loss_out, float* zero_buf, int zero_n, void* stream_, const float* raw_rows, int raw_cap,
    const float* bank_n_in, uint64_t* seed_dev, int rows_mode, FillShare fill) {
  cudaStream_t stream = (cudaStream_t)stream_;
  need_grad &= 1;   // bit 1 of the public argument selects rows_mode (passed separately here)
  const int B = batch, D = dim, C = n_classes, M = sub_protos;
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(B > 0 && B <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(C >= 2 && C <= kMaxClasses, "n_classes must be in [2, %d]", kMaxClasses);
  C3D_REQUIRE(D > 0 && D % 4 == 0 && D <= 1024, "feature dim must be a multiple of 4, <= 1024");
  C3D_REQUIRE(M > 0 && num_anchor > 0, "sub_protos and num_anchor must be positive");
  C3D_REQUIRE(HWll > 0 && B * HWll < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(workspace, "null workspace");
  C3D_REQUIRE(!(phases & kPhaseSplit) || (probs && labels), "null pointer argument (probs / labels)");
  C3D_REQUIRE(!(phases & kPhaseRows) || (feats && (proto_queue || bank_n_in) && loss_out),
              "null pointer argument (feats / proto_queue / loss_out)");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256 B aligned");
  C3D_REQUIRE(temperature > 0 && base_temperature > 0, "temperatures must be positive");
  const int HW = (int)HWll;
  LossWs w = carve(workspace, B, C, HW, D, M, num_anchor);
  const int Kc = (C - 1) * M;
  int tile_rows, n_tiles; size_t smem;
  C3D_REQUIRE(rows_config(D, Kc, &tile_rows, &n_tiles, &smem) == 0,
              "bank does not fit shared memory tiling (D=%d, Kc=%d)", D, Kc);

  int rc;
  if (phases & kPhaseSplit) {
  C3D_CUDA(cudaMemsetAsync(w.s.info, 0, (size_t)(8 + B) * 4, stream));
  if ((rc = launch_split((const long long*)labels, keep_mask, probs, B, C, HW, ignore_label, w.s, w.w_list,
                         w.cnt_list, zero_buf, zero_n, stream,
                         (phases & kPhaseRows) ? FillShare{nullptr, 0} : fill))) return rc;
  }
  if (phases & kPhaseSample) {
  { KernelTimer kt__("loss_sample_kernel", stream);
    loss_sample_kernel<<<B * C, 256, 0, stream>>>(w.s.seg_cnt, w.s.seg_start, w.s.seg_tidx, w.s.pix_list,
                                                  w.w_list, w.cnt_list, HW, B, C, num_anchor,
                                                  (const long long*)keep, keep_rows, seed,
                                                  reinterpret_cast<unsigned long long*>(seed_dev), w.seg_nd,
                                                  w.row_base, w.seg_of_t, w.s.info); }
  if ((rc = check_launch("loss_sample_kernel"))) return rc;
  }
  if (!(phases & kPhaseRows)) return C3D_OK;

  // F.normalize of the bank rows of classes 1..C-1 (:167).  Part of phase 2: the bank may be
  // written between the phases (the EMA update precedes the loss in a training step,
  // salsanext_proto.py:520-527 -> trainer.py:675-686).
  if (!bank_n_in) {
    { KernelTimer kt__("bank_normalise_kernel", stream);
      bank_normalise_kernel<<<(Kc + 7) / 8, 256, 0, stream>>>(proto_queue + (size_t)M * D, Kc, D, w.bank_n); }
    if ((rc = check_launch("bank_normalise_kernel"))) return rc;
  }

  RowsParams p{};
  p.feats = feats; p.raw_rows = raw_rows; p.raw_cap = raw_cap; p.fill = fill; p.bank_n = bank_n_in ? bank_n_in + (size_t)M * D : w.bank_n; p.pix_list = w.s.pix_list; p.cls_list = w.s.cls_list;
  p.cnt_list = w.cnt_list; p.dist_list = reinterpret_cast<const int32_t*>(w.w_list);
  p.seg_start = w.s.seg_start; p.row_base = w.row_base; p.seg_of_t = w.seg_of_t; p.info = w.s.info;
  p.loss_part = w.loss_part; p.row_pix = w.row_pix; p.grad_rows = w.grad_rows; p.loss_out = loss_out;
  p.HW = HW; p.D = D; p.M = M; p.Kc = Kc; p.A = num_anchor; p.tile_rows = tile_rows;
  p.n_tiles = n_tiles; p.temperature = temperature; p.base_temperature = base_temperature;
  RowsPlan plan;
  // rows_mode 1: tensor cores (mma.sync 3xTF32) for the two products, D a multiple of 32
  if (rows_mode == 1 && D % 32 == 0 && D <= 256 && plan_rows16(D, Kc, &plan, kMmaPadA) == 0) {
    p.tile_rows = plan.tile_rows; p.n_tiles = plan.n_tiles; p.ldl = plan.ldl;
    if (need_grad) {
      if (D <= 32) return launch_rows16<true, 1, 4, 1, 1, true>(p, plan.smem, stream);
      if (D <= 64) return launch_rows16<true, 1, 4, 1, 2, true>(p, plan.smem, stream);
      if (D <= 128) return launch_rows16<true, 1, 4, 1, 4, true>(p, plan.smem, stream);
      return launch_rows16<true, 1, 4, 1, 8, true>(p, plan.smem, stream);
    }
    if (D <= 32) return launch_rows16<false, 1, 4, 1, 1, true>(p, plan.smem, stream);
    if (D <= 64) return launch_rows16<false, 1, 4, 1, 2, true>(p, plan.smem, stream);
    if (D <= 128) return launch_rows16<false, 1, 4, 1, 4, true>(p, plan.smem, stream);
    return launch_rows16<false, 1, 4, 1, 8, true>(p, plan.smem, stream);
  }
  if (plan_rows16(D, Kc, &plan) == 0) {
    p.tile_rows = plan.tile_rows; p.n_tiles = plan.n_tiles; p.ldl = plan.ldl;
    // <grad, k-splits, rows per thread in P3, chunks per thread, D/32>; chunk-threads
    // CT = 256 / (kKS * 16 / kRP) must cover D/4 chunks (times kDch)
    if (need_grad) {
      if (D <= 64) return launch_rows16<true, 4, 4, 1, 2>(p, plan.smem, stream);     // CT = 16
      if (D <= 128) return launch_rows16<true, 2, 4, 1, 4>(p, plan.smem, stream);    // CT = 32
      if (D <= 256) return launch_rows16<true, 1, 4, 1, 8>(p, plan.smem, stream);    // CT = 64
      return launch_rows16<true, 1, 4, 4, 32>(p, plan.smem, stream);
    }
    if (D <= 64) return launch_rows16<false, 4, 4, 1, 2>(p, plan.smem, stream);
    if (D <= 128) return launch_rows16<false, 2, 4, 1, 4>(p, plan.smem, stream);
    if (D <= 256) return launch_rows16<false, 1, 4, 1, 8>(p, plan.smem, stream);
    return launch_rows16<false, 1, 4, 4, 32>(p, plan.smem, stream);
  }
  C3D_REQUIRE(raw_rows == nullptr && fill.bytes == 0, "raw rows / carried fill need the tiled loss rows kernel");
  if (!need_grad) return launch_rows<false, 1>(p, smem, stream);
  if (D <= 128) return launch_rows<true, 1>(p, smem, stream);
  if (D <= 256) return launch_rows<true, 2>(p, smem, stream);
  if (D <= 512) return launch_rows<true, 4>(p, smem, stream);
  return launch_rows<true, 8>(p, smem, stream);
}

extern "C" int c3d_proto_loss_forward(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* proto_queue, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, uint64_t seed, int need_grad, void* workspace,
    float* loss_out, void* stream) {
  return proto_loss_forward_impl(feats, probs, labels, keep_mask, proto_queue, batch, dim, proj_h, proj_w,
                                 n_classes, sub_protos, ignore_label, temperature, base_temperature,
                                 num_anchor, keep, keep_rows, seed, need_grad,
                                 kPhaseSplit | kPhaseSample | kPhaseRows, workspace, loss_out, nullptr, 0,
                                 stream, nullptr, 0, nullptr, nullptr, need_grad >> 1, FillShare{nullptr, 0});
}

extern "C" int c3d_proto_loss_forward_phase(
    const float* feats, const float* probs, const int64_t* labels, const uint8_t* keep_mask,
    const float* proto_queue, int batch, int dim, int proj_h, int proj_w, int n_classes,
    int sub_protos, int ignore_label, float temperature, float base_temperature, int num_anchor,
    const int64_t* keep, int keep_rows, uint64_t seed, int need_grad, int phases, void* workspace,
    float* loss_out, void* stream) {
  C3D_REQUIRE(phases >= 1 && phases <= 3, "phases must be 1 (select), 2 (rows) or 3 (both)");
  const int internal = ((phases & 1) ? (kPhaseSplit | kPhaseSample) : 0) | ((phases & 2) ? kPhaseRows : 0);
  return proto_loss_forward_impl(feats, probs, labels, keep_mask, proto_queue, batch, dim, proj_h, proj_w,
                                 n_classes, sub_protos, ignore_label, temperature, base_temperature,
                                 num_anchor, keep, keep_rows, seed, need_grad, internal, workspace, loss_out,
                                 nullptr, 0, stream, nullptr, 0, nullptr, nullptr, need_grad >> 1, FillShare{nullptr, 0});
}

extern "C" int c3d_proto_loss_backward(int batch, int dim, int proj_h, int proj_w, int n_classes,
                                       int sub_protos, int num_anchor, void* workspace,
                                       const float* grad_out, float* grad_feats, int grad_is_zeroed,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = batch, D = dim, C = n_classes, M = sub_protos;
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(B > 0 && B <= kMaxBatch && C >= 2 && C <= kMaxClasses, "bad batch / n_classes");
  C3D_REQUIRE(D > 0 && D % 4 == 0 && D <= 1024, "feature dim must be a multiple of 4, <= 1024");
  C3D_REQUIRE(HWll > 0 && B * HWll < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(workspace && grad_out && grad_feats, "null pointer argument");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(grad_feats) & 15) == 0, "grad_feats must be 16 B aligned");
  const int HW = (int)HWll;
  LossWs w = carve(workspace, B, C, HW, D, M, num_anchor);
  int rc;
  if (!grad_is_zeroed && (rc = launch_fill(grad_feats, (size_t)B * D * HW * 4, stream))) return rc;
  { KernelTimer kt__("loss_grad_scatter_kernel", stream);
    loss_grad_scatter_kernel<<<kNumSMs * 2, 256, 0, stream>>>(
        w.grad_rows, w.row_pix, w.s.info, grad_out, HW, D, grad_feats); }
  return check_launch("loss_grad_scatter_kernel");
}

extern "C" int c3d_zero_fill(void* dst, size_t nbytes, void* stream_) {
  // The dense-gradient zero fill as its own entry point, so a caller can run it on
  // a side stream concurrently with the forward pass (then pass grad_is_zeroed=1).
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(dst && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16 B aligned");
  C3D_REQUIRE(nbytes % 4 == 0, "nbytes must be a multiple of 4");
  return launch_fill(dst, nbytes, stream);
}

extern "C" int c3d_proto_loss_info(const void* workspace, int32_t* host_info4, void* stream_) {
  // Synchronous debug/strict helper: copies {T, labelled slots, flags, 0} to the host.
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(workspace && host_info4, "null pointer argument");
  C3D_CUDA(cudaMemcpyAsync(host_info4, workspace, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  C3D_CUDA(cudaStreamSynchronize(stream));
  return C3D_OK;
}

extern "C" int c3d_proto_loss_rows(const void* workspace, int batch, int dim, int hw, int n_classes,
                                   int sub_protos, int num_anchor, int64_t capacity, int32_t* pix,
                                   int32_t* cls, int32_t* cnt, void* stream_) {
  // Exports the labelled-pixel slots of the last forward (device to device):
  // pix = scan*HW + pixel, cls = class, cnt = how many anchors hit the slot.
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(workspace && pix && cls && cnt, "null pointer argument");
  C3D_REQUIRE(capacity > 0 && capacity <= (int64_t)batch * hw, "bad capacity");
  LossWs w = carve(const_cast<void*>(workspace), batch, n_classes, hw, dim, sub_protos, num_anchor);
  const size_t n = (size_t)capacity * 4;
  C3D_CUDA(cudaMemcpyAsync(pix, w.s.pix_list, n, cudaMemcpyDeviceToDevice, stream));
  C3D_CUDA(cudaMemcpyAsync(cls, w.s.cls_list, n, cudaMemcpyDeviceToDevice, stream));
  C3D_CUDA(cudaMemcpyAsync(cnt, w.cnt_list, n, cudaMemcpyDeviceToDevice, stream));
  return C3D_OK;
}

#ifdef C3D_DEBUG_STAMPS
extern "C" int c3d_debug_rows_stamps(long long* host16) {
  C3D_CUDA(cudaDeviceSynchronize());
  C3D_CUDA(cudaMemcpyFromSymbol(host16, g_rows_dbg, sizeof(long long) * 16));
  return C3D_OK;
}
#endif
