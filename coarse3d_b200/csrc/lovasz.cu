// f4 -- Lovasz-softmax loss, forward and backward, for the weak-label regime.
//
// Replaces lovasz_softmax / lovasz_softmax_flat / lovasz_grad / flatten_probas (reference
// pc_processor/loss/lovasz_softmax.py:51-157) as the trainer calls them
// (tasks/weak_segmentation/trainer.py:362-364,650: probabilities (B,C,H,W), weak labels,
// ignore = ignore_cls, classes = "present", per_image = False).
//
// Per class the reference sorts the P valid pixels' errors (torch.sort, one launch chain per
// class) and evaluates  loss_c = sum_r e_sorted[r] * (J_r - J_{r-1}),  J_r = 1 - I_r / U_r
// (:51-64); J_r depends only on r and F_r = #foreground among the r+1 largest errors.
//   L1 lovasz_compact  valid pixels (label != ignore) -> (pixel, label) list, class histogram
//   L2 lovasz_sort     one CTA per class: 64-bit keys (error bits | pixel | sign | foreground)
//                      built from the probabilities, bitonic sort in shared memory (P <= 16384:
//                      128 KB of keys), block scan of the foreground bits -> F_r, closed-form
//                      gradient entry and e * g term per rank, fixed-order class sum
//   L3 lovasz_finalize mean over the averaged classes (:33-48)
//   backward           dense (B,C,H,W) zero fill + P x C scattered entries, scaled by grad_out
// For 16384 < P <= 32768 the keys do not fit shared memory: lovasz_keys + lovasz_rank count,
// per element, the larger keys of its class by all pairs (tiles broadcast from shared memory;
// quadratic, 0.75 ms at P = 8.4k, so only the fallback) and lovasz_reduce sums the classes.
// Keys are unique (they embed the pixel), so both paths give identical ranks, no atomics are
// needed, and the result does not depend on the order of the compacted list.
// Dense / pseudo-label regime (max_valid > 32768, up to 2^24 pixels in the batch): the keys of
// ALL classes go through ONE device-wide radix sort -- class id in the top 6 key bits, so each
// class comes out as a contiguous descending run (cub::DeviceRadixSort, the CUDA toolkit's
// library sort: the one library call of this repository, used where the reference calls
// torch.sort) -- and the runs are walked in 8192-rank chunks by the whole GPU: foreground count
// per chunk, then per chunk the blocked prefix count of the foreground bits, the closed-form
// gradient entry and the e * g term per rank.
// Tie rule (torch.sort is unstable): equal errors rank by pixel index.
#include <math_constants.h>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace c3d {

constexpr int kLovMaxClasses = 64;
constexpr long long kLovMaxPairs = 32768;          // all-pairs rank path (quadratic) up to here
constexpr long long kLovMaxValid = 1ll << 24;      // radix path: 24 pixel bits in the key
constexpr int kLovSortMax = 16384;        // keys of one class that fit shared memory (128 KB)
enum LovInfo { kLovP = 0, kLovPresent = 1, kLovFlags = 2 };
enum LovFlag { kLovOverflow = 1, kLovEmpty = 2 };
// classes = 'present' (mode 0: classes with a foreground pixel), 'all' (1), or a list (2: the
// classes of `mask`, absent ones included -- lovasz_softmax.py:117-122)
__device__ __forceinline__ bool lov_included(int mode, unsigned long long mask, int hist_c, int c) {
  return mode == 1 || (mode == 0 && hist_c > 0) || (mode == 2 && ((mask >> c) & 1ull));
}

struct LovWs {
  int32_t* info;      // [4] P, n_present, flags
  int32_t* hist;      // [C] foreground count per class (gts)
  int32_t* pix;       // [cap] b*HW + pixel
  int32_t* lab;       // [cap]
  float* cls_loss;    // [C] loss_c
  unsigned long long* keys;  // [C * cap] fallback / radix paths only
  unsigned long long* keys2; // [C * cap] radix path: sorted keys
  void* sort_tmp;     // radix path: cub temporary storage
  size_t sort_tmp_bytes;
  int32_t* chunk_fg;  // [C * n_chunks] radix path: foreground keys per 8192-rank chunk
  float* chunk_loss;  // [C * n_chunks] radix path: partial losses
  float* term;        // [C * cap] e * g at slot rank (fallback path)
  float* gval;        // [C * cap] d loss_c / d p of the element at slot rank
  int32_t* gpix;      // [C * cap] its pixel
  size_t bytes;
};

static LovWs carve_lov(void* base, int C, long long cap) {
  LovWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~(size_t)255; return (char*)base + o; };
  w.info = (int32_t*)take(16);
  w.hist = (int32_t*)take((size_t)C * 4);
  w.pix = (int32_t*)take((size_t)cap * 4);
  w.lab = (int32_t*)take((size_t)cap * 4);
  w.cls_loss = (float*)take((size_t)C * 4);
  w.gval = (float*)take((size_t)C * cap * 4);
  w.gpix = (int32_t*)take((size_t)C * cap * 4);
  const bool fallback = cap > kLovSortMax, radix = cap > kLovMaxPairs;
  w.keys = (unsigned long long*)take(fallback ? (size_t)C * cap * 8 : 0);
  w.term = (float*)take(fallback && !radix ? (size_t)C * cap * 4 : 0);
  w.keys2 = (unsigned long long*)take(radix ? (size_t)C * cap * 8 : 0);
  w.sort_tmp_bytes = 0;
  if (radix) {   // size query only: no launch, no allocation
    const unsigned long long* kin = nullptr; unsigned long long* kout = nullptr;
    cub::DeviceRadixSort::SortKeysDescending(nullptr, w.sort_tmp_bytes, kin, kout, C * cap, 0, 64);
  }
  w.sort_tmp = take(w.sort_tmp_bytes);
  const size_t n_chunks = radix ? (size_t)((cap + 8191) / 8192) : 0;
  w.chunk_fg = (int32_t*)take((size_t)C * n_chunks * 4);
  w.chunk_loss = (float*)take((size_t)C * n_chunks * 4);
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------- L1 -------
__global__ void __launch_bounds__(256)
lovasz_compact_kernel(const long long* __restrict__ labels, long long total, int C, int ignore,
                      int cap, int32_t* __restrict__ pix, int32_t* __restrict__ lab,
                      int32_t* __restrict__ hist, int32_t* __restrict__ info) {
  // class histogram per CTA in shared memory, flushed once (dense labels: one global atomic per
  // valid pixel on ~20 addresses was most of this kernel's time)
  __shared__ int s_hist[kLovMaxClasses];
  if (threadIdx.x < kLovMaxClasses) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x; q0 < total; q0 += stride) {
    const long long q = q0 + threadIdx.x;
    long long l = ignore;
    if (q < total) l = labels[q];
    const bool valid = (q < total) && (l != ignore) && (l >= 0) && (l < C);
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&info[kLovP], __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (valid) {
        const int pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < cap) { pix[pos] = (int)q; lab[pos] = (int)l; atomicAdd(&s_hist[(int)l], 1); }
        else atomicOr(&info[kLovFlags], kLovOverflow);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < C && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
}

// Key of (class c, element): descending error, then ascending pixel -- a larger key sorts
// first.  Bit 1: d|fg - p|/dp is negative (p < fg); bit 0: foreground.  Needs B*H*W <= 2^30.
__device__ __forceinline__ unsigned long long lovasz_key(float p, bool fg, int gpix) {
  const float fgf = fg ? 1.0f : 0.0f;
  const float e = fabsf(fgf - p);                                  // lovasz_softmax.py:129
  return ((unsigned long long)__float_as_uint(e) << 32) |
         ((unsigned long long)(0x3FFFFFFFu - (unsigned)gpix) << 2) | (p < fgf ? 2ull : 0ull) |
         (fg ? 1ull : 0ull);
}
__device__ __forceinline__ int key_pixel(unsigned long long k) {
  return (int)(0x3FFFFFFFu - (unsigned)((k >> 2) & 0x3FFFFFFFu));
}
// torch's abs backward: sign(fg - p) * -1, 0 at 0
__device__ __forceinline__ float key_sign(unsigned long long k) {
  return ((unsigned)(k >> 32) == 0u) ? 0.0f : ((k & 2ull) ? -1.0f : 1.0f);
}

// J_r of lovasz_grad (:51-64): gts foreground pixels in total, F of them among the first r+1
__device__ __forceinline__ float jaccard_at(float gts, int r, int F) {
  const float inter = gts - (float)F;
  const float uni = gts + (float)(r + 1 - F);
  return 1.0f - inter / uni;
}
__device__ __forceinline__ float lovasz_grad_at(float gts, int r, int F, int fg) {
  float g = jaccard_at(gts, r, F);
  if (r > 0) g = g - jaccard_at(gts, r - 1, F - fg);               // :62-63
  return g;
}

// ---------------------------------------------------------------- L2 -------
// One CTA (1024 threads) per class.  n = next power of two >= P keys in dynamic shared memory.
__global__ void __launch_bounds__(1024)
lovasz_sort_kernel(const float* __restrict__ probs, int HW, int C, int cap, int classes_all, unsigned long long cls_mask,
                   const int32_t* __restrict__ pix, const int32_t* __restrict__ lab,
                   const int32_t* __restrict__ hist, const int32_t* __restrict__ info,
                   float* __restrict__ gval, int32_t* __restrict__ gpix,
                   float* __restrict__ cls_loss) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ int s_warp[32];
  __shared__ float s_sum[32];
  const int c = blockIdx.x;
  const int P = min(info[kLovP], cap);
  if (P > kLovSortMax) return;                                     // fallback path runs instead
  if (threadIdx.x == 0) cls_loss[c] = 0.0f;
  if (!lov_included(classes_all, cls_mask, hist[c], c) || P == 0) return;   // :121-122
  int n = 1;
  while (n < P) n <<= 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long k = 0ull;                                   // padding sorts last
    if (i < P) {
      const int g = pix[i];
      const int b = g / HW, hw = g - b * HW;
      k = lovasz_key(__ldg(probs + ((size_t)b * C + c) * HW + hw), lab[i] == c, g);  // :140-149, :120
    }
    s_keys[i] = k;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long a = s_keys[i], b2 = s_keys[l];
        const bool desc = (i & k) == 0;
        if ((a < b2) == desc) { s_keys[i] = b2; s_keys[l] = a; }
      }
      __syncthreads();
    }
  }
  // thread t owns ranks [t*per, (t+1)*per): foreground prefix count by a block scan
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int r0 = threadIdx.x * per;
  int local = 0;
  for (int q = 0; q < per; ++q) { const int r = r0 + q; if (r < P) local += (int)(s_keys[r] & 1ull); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = incl - local;
  for (int wv = 0; wv < warp; ++wv) before += s_warp[wv];
  const float gts = (float)hist[c];
  float part = 0.0f;
  int F = before;
  for (int q = 0; q < per; ++q) {
    const int r = r0 + q;
    if (r >= P) break;
    const unsigned long long k = s_keys[r];
    const int fg = (int)(k & 1ull);
    F += fg;
    const float g = lovasz_grad_at(gts, r, F, fg);
    part += __uint_as_float((unsigned)(k >> 32)) * g;              // dot(errors_sorted, grad) :133
    gval[(size_t)c * cap + r] = g * key_sign(k);
    gpix[(size_t)c * cap + r] = key_pixel(k);
  }
  part = warp_sum(part);
  if (lane == 0) s_sum[warp] = part;
  __syncthreads();
  if (warp == 0) {
    float t = s_sum[lane];
    t = warp_sum(t);
    if (lane == 0) cls_loss[c] = t;
  }
}

// ---------------------------------------------------------------- L3 -------
__global__ void lovasz_finalize_kernel(int C, int cap, int classes_all, unsigned long long cls_mask, const int32_t* __restrict__ hist,
                                       int32_t* __restrict__ info, const float* __restrict__ cls_loss,
                                       float* __restrict__ loss_out) {
  if (threadIdx.x != 0) return;
  const int P = min(info[kLovP], cap);
  int n = 0;
  float acc = 0.0f;
  for (int c = 0; c < C; ++c) {
    if (P == 0 || !lov_included(classes_all, cls_mask, hist[c], c)) continue;
    acc += cls_loss[c];                                            // mean(): acc = acc + v (:44-45)
    ++n;
  }
  if (P == 0) atomicOr(&info[kLovFlags], kLovEmpty);
  info[kLovPresent] = n;
  // More valid pixels than the workspace holds: an arbitrary subset was kept, so the value
  // would be wrong and non-deterministic.  Fail loudly without a host round trip: the loss is
  // NaN (the module's own NaN assertion, lovasz_softmax.py:178, then fires).
  const bool overflow = (info[kLovFlags] & kLovOverflow) != 0;
  *loss_out = overflow ? __int_as_float(0x7fc00000) : ((n > 1) ? acc / (float)n : acc);  // :46-48
}

// ------------------------------------------------------ fallback: keys ----
__global__ void __launch_bounds__(256)
lovasz_keys_kernel(const float* __restrict__ probs, int HW, int C, int cap, int classes_all, unsigned long long cls_mask,
                   const int32_t* __restrict__ pix, const int32_t* __restrict__ lab,
                   const int32_t* __restrict__ hist, const int32_t* __restrict__ info,
                   unsigned long long* __restrict__ keys) {
  const int c = blockIdx.y;
  const int P = min(info[kLovP], cap);
  if (P <= kLovSortMax) return;                                   // the sort path handled it
  if (!lov_included(classes_all, cls_mask, hist[c], c)) return;   // lovasz_softmax.py:121-122
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int g = pix[i];
  const int b = g / HW, hw = g - b * HW;
  keys[(size_t)c * cap + i] = lovasz_key(__ldg(probs + ((size_t)b * C + c) * HW + hw), lab[i] == c, g);
}

// ------------------------------------------------------ fallback: rank ----
constexpr int kLovTile = 1024;
__global__ void __launch_bounds__(256)
lovasz_rank_kernel(int C, int cap, int classes_all, unsigned long long cls_mask, const int32_t* __restrict__ hist,
                   const int32_t* __restrict__ info, const unsigned long long* __restrict__ keys,
                   float* __restrict__ term, float* __restrict__ gval, int32_t* __restrict__ gpix) {
  __shared__ unsigned long long s_key[kLovTile];
  const int c = blockIdx.y;
  const int P = min(info[kLovP], cap);
  if (P <= kLovSortMax) return;
  if (!lov_included(classes_all, cls_mask, hist[c], c)) return;
  if (blockIdx.x * blockDim.x >= P) return;
  const unsigned long long* kc = keys + (size_t)c * cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < P;
  const unsigned long long mine = act ? kc[i] : ~0ull;
  int rank = 0, before_fg = 0;
  for (int j0 = 0; j0 < P; j0 += kLovTile) {
    const int n = min(kLovTile, P - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) s_key[j] = kc[j0 + j];
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < n; ++j) {
      const unsigned long long k = s_key[j];
      const bool gt = k > mine;
      rank += gt ? 1 : 0;
      before_fg += (gt && (k & 1ull)) ? 1 : 0;
    }
  }
  if (!act) return;
  const int fg = (int)(mine & 1ull);
  const float g = lovasz_grad_at((float)hist[c], rank, before_fg + fg, fg);
  term[(size_t)c * cap + rank] = __uint_as_float((unsigned)(mine >> 32)) * g;
  gval[(size_t)c * cap + rank] = g * key_sign(mine);
  gpix[(size_t)c * cap + rank] = key_pixel(mine);
}

// ---------------------------------------------------- fallback: reduce ----
__global__ void __launch_bounds__(1024)
lovasz_reduce_kernel(int C, int cap, int classes_all, unsigned long long cls_mask, const int32_t* __restrict__ hist,
                     const int32_t* __restrict__ info, const float* __restrict__ term,
                     float* __restrict__ cls_loss) {
  __shared__ float s_part[32];
  const int c = blockIdx.x;
  const int P = min(info[kLovP], cap);
  if (P <= kLovSortMax) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v = 0.0f;
  if (lov_included(classes_all, cls_mask, hist[c], c))
    for (int r = threadIdx.x; r < P; r += blockDim.x) v += term[(size_t)c * cap + r];
  v = warp_sum(v);
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = s_part[lane];
    t = warp_sum(t);
    if (lane == 0) cls_loss[c] = t;
  }
}

// ------------------------------------------------------ radix path: keys ----
// Key of (class c, element): [63:58] class, [57:26] error bits, [25:2] 0xFFFFFF - pixel,
// [1] sign of d|fg - p|/dp, [0] foreground.  One descending sort of all C * cap keys leaves every
// included class as a contiguous run (classes descending), padding and excluded classes (key 0)
// at the very end.
__device__ __forceinline__ unsigned long long lovasz_key_big(int c, float p, bool fg, int gpix) {
  const float fgf = fg ? 1.0f : 0.0f;
  const float e = fabsf(fgf - p);
  return ((unsigned long long)c << 58) | ((unsigned long long)__float_as_uint(e) << 26) |
         ((unsigned long long)(0xFFFFFFu - (unsigned)gpix) << 2) | (p < fgf ? 2ull : 0ull) | (fg ? 1ull : 0ull);
}

__global__ void __launch_bounds__(256)
lovasz_keys_big_kernel(const float* __restrict__ probs, int HW, int C, int cap, int classes_all,
                       unsigned long long cls_mask, const int32_t* __restrict__ pix,
                       const int32_t* __restrict__ lab, const int32_t* __restrict__ hist,
                       const int32_t* __restrict__ info, unsigned long long* __restrict__ keys) {
  const int c = blockIdx.y;
  const int P = min(info[kLovP], cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  unsigned long long k = 0ull;
  if (i < P && lov_included(classes_all, cls_mask, hist[c], c)) {
    const int g = pix[i];
    const int b = g / HW, hw = g - b * HW;
    // c + 1 in the class field: a real key is never 0 (error 0, pixel 0xFFFFFF, class 0 would be)
    k = lovasz_key_big(c + 1, __ldg(probs + ((size_t)b * C + c) * HW + hw), lab[i] == c, g);
  }
  keys[(size_t)c * cap + i] = k;
}

// ------------------------------------------------------ radix path: runs ----
// The class's sorted run is cut into chunks of 8192 ranks, one CTA (1024 threads, 8 consecutive
// ranks per thread) per (chunk, class), so the whole GPU works on it:
//   lovasz_chunk_count  foreground keys per chunk;
//   lovasz_chunk_apply  F_r = (foreground in the earlier chunks) + a block-wide exclusive scan of
//                       the per-thread counts; closed-form gradient entry and e * g term per
//                       rank; the chunk's partial loss;
//   lovasz_chunk_reduce the class's partial losses summed in chunk order (fixed order).
constexpr int kLovItems = 8;
constexpr int kLovChunk = 1024 * kLovItems;

// first key of class c's run and its length P (0 if the class is not averaged)
__device__ __forceinline__ const unsigned long long* lov_run(const unsigned long long* sorted, int C, int c, int P,
                                                              int classes_all, unsigned long long cls_mask,
                                                              const int32_t* __restrict__ hist) {
  int later = 0;   // the runs of the included classes with a larger id come first
  for (int cc = c + 1; cc < C; ++cc) later += lov_included(classes_all, cls_mask, hist[cc], cc) ? 1 : 0;
  return sorted + (size_t)later * P;
}

__global__ void __launch_bounds__(1024)
lovasz_chunk_count_kernel(int C, int cap, int n_chunks, int classes_all, unsigned long long cls_mask,
                          const int32_t* __restrict__ hist, const int32_t* __restrict__ info,
                          const unsigned long long* __restrict__ sorted, int32_t* __restrict__ chunk_fg) {
  __shared__ int s_warp[32];
  const int c = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = min(info[kLovP], cap);
  int cnt = 0;
  if (lov_included(classes_all, cls_mask, hist[c], c) && chunk * kLovChunk < P) {
    const unsigned long long* run = lov_run(sorted, C, c, P, classes_all, cls_mask, hist);
    const int base = chunk * kLovChunk + threadIdx.x * kLovItems;
#pragma unroll
    for (int j = 0; j < kLovItems; ++j) cnt += (base + j < P) ? (int)(run[base + j] & 1ull) : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    const int t = __reduce_add_sync(0xffffffffu, s_warp[lane]);
    if (lane == 0) chunk_fg[c * n_chunks + chunk] = t;
  }
}

__global__ void __launch_bounds__(1024)
lovasz_chunk_apply_kernel(int C, int cap, int n_chunks, int classes_all, unsigned long long cls_mask,
                          const int32_t* __restrict__ hist, const int32_t* __restrict__ info,
                          const unsigned long long* __restrict__ sorted, const int32_t* __restrict__ chunk_fg,
                          float* __restrict__ gval, int32_t* __restrict__ gpix, float* __restrict__ chunk_loss) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ float s_part[32];
  const int c = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = min(info[kLovP], cap);
  if (!lov_included(classes_all, cls_mask, hist[c], c) || chunk * kLovChunk >= P) {
    if (threadIdx.x == 0) chunk_loss[c * n_chunks + chunk] = 0.0f;
    return;
  }
  const unsigned long long* run = lov_run(sorted, C, c, P, classes_all, cls_mask, hist);
  const float gts = (float)hist[c];
  if (warp == 0) {            // foreground keys in the earlier chunks of this class
    int t = 0;
    for (int j = lane; j < chunk; j += 32) t += chunk_fg[c * n_chunks + j];
    t = __reduce_add_sync(0xffffffffu, t);
    if (lane == 0) s_carry = t;
  }
  const int base = chunk * kLovChunk + threadIdx.x * kLovItems;
  unsigned long long k[kLovItems];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < kLovItems; ++j) {
    k[j] = (base + j < P) ? run[base + j] : 0ull;
    cnt += (int)(k[j] & 1ull);
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
    s_warp[lane] = wi - w;                    // exclusive prefix over warps
  }
  __syncthreads();
  int F = s_carry + s_warp[warp] + incl - cnt;   // foreground among the ranks before this thread's
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < kLovItems; ++j) {
    const int r = base + j;
    if (r < P) {
      const int fg = (int)(k[j] & 1ull);
      F += fg;
      const float g = lovasz_grad_at(gts, r, F, fg);
      const unsigned ebits = (unsigned)((k[j] >> 26) & 0xFFFFFFFFull);
      acc += __uint_as_float(ebits) * g;
      const float sign = (ebits == 0u) ? 0.0f : ((k[j] & 2ull) ? -1.0f : 1.0f);
      gval[(size_t)c * cap + r] = g * sign;
      gpix[(size_t)c * cap + r] = (int)(0xFFFFFFu - (unsigned)((k[j] >> 2) & 0xFFFFFFull));
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float t = s_part[lane];
    t = warp_sum(t);
    if (lane == 0) chunk_loss[c * n_chunks + chunk] = t;
  }
}

__global__ void __launch_bounds__(32)
lovasz_chunk_reduce_kernel(int n_chunks, const float* __restrict__ chunk_loss, float* __restrict__ cls_loss) {
  const int c = blockIdx.x, lane = threadIdx.x;
  float t = 0.0f;
  for (int j = lane; j < n_chunks; j += 32) t += chunk_loss[c * n_chunks + j];   // lane-strided, then a tree:
  t = warp_sum(t);                                                               // one fixed order
  if (lane == 0) cls_loss[c] = t;
}

// ------------------------------------------------------------- backward ----
__global__ void __launch_bounds__(256)
lovasz_scatter_kernel(int HW, int C, int cap, int classes_all, unsigned long long cls_mask, const int32_t* __restrict__ hist,
                      const int32_t* __restrict__ info, const float* __restrict__ gval,
                      const int32_t* __restrict__ gpix, const float* __restrict__ grad_out,
                      float* __restrict__ grad_probs) {
  const int c = blockIdx.y;
  const int P = min(info[kLovP], cap);
  if (!lov_included(classes_all, cls_mask, hist[c], c)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const int n_present = info[kLovPresent];
  const float scale = __ldg(grad_out) / (float)(n_present > 1 ? n_present : 1);
  const int g = gpix[(size_t)c * cap + r];
  const int b = g / HW, hw = g - b * HW;
  grad_probs[((size_t)b * C + c) * HW + hw] = gval[(size_t)c * cap + r] * scale;
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_lovasz_workspace_bytes(int n_classes, int64_t max_valid) {
  if (n_classes < 1 || n_classes > kLovMaxClasses || max_valid <= 0 || max_valid > kLovMaxValid) return 0;
  return carve_lov(nullptr, n_classes, max_valid).bytes;
}

extern "C" int c3d_lovasz_forward(const float* probs, const int64_t* labels, int batch, int n_classes,
                                  int proj_h, int proj_w, int ignore, int classes_all, uint64_t class_mask,
                                  int64_t max_valid, void* workspace, float* loss_out,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned long long cls_mask = class_mask;
  C3D_REQUIRE(classes_all >= 0 && classes_all <= 2, "classes: 0 present, 1 all, 2 the classes of class_mask");
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(n_classes >= 1 && n_classes <= kLovMaxClasses, "n_classes must be in [1, %d]", kLovMaxClasses);
  C3D_REQUIRE(HWll > 0 && batch * HWll <= (1ll << 30), "batch*H*W must be <= 2^30");
  C3D_REQUIRE(probs && labels && workspace && loss_out, "null pointer argument");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256 B aligned");
  if (max_valid <= 0 || max_valid > kLovMaxValid ||
      (max_valid > kLovMaxPairs && (batch * HWll > kLovMaxValid || n_classes > 63))) {
    set_error("Lovasz: max_valid=%lld outside [1, %lld] (the radix path keeps 24 pixel bits in its keys: "
              "batch*H*W <= 2^24, n_classes <= 63)", (long long)max_valid, kLovMaxValid);
    return C3D_UNSUPPORTED;
  }
  const int HW = (int)HWll, C = n_classes, cap = (int)max_valid;
  const long long total = (long long)batch * HW;
  LovWs w = carve_lov(workspace, C, cap);
  C3D_CUDA(cudaMemsetAsync(w.info, 0, (size_t)((char*)w.pix - (char*)w.info), stream));  // info + hist
  int rc;
  {
    KernelTimer kt__("lovasz_compact_kernel", stream);
    lovasz_compact_kernel<<<wave_grid(total, 256, 8), 256, 0, stream>>>(
        (const long long*)labels, total, C, ignore, cap, w.pix, w.lab, w.hist, w.info);
  }
  if ((rc = check_launch("lovasz_compact_kernel"))) return rc;
  if (cap > kLovMaxPairs) {   // dense / pseudo-label regime: one device-wide radix sort
    const dim3 grid((cap + 255) / 256, C);
    {
      KernelTimer kt__("lovasz_keys_big_kernel", stream);
      lovasz_keys_big_kernel<<<grid, 256, 0, stream>>>(probs, HW, C, cap, classes_all, cls_mask, w.pix, w.lab,
                                                       w.hist, w.info, w.keys);
    }
    if ((rc = check_launch("lovasz_keys_big_kernel"))) return rc;
    {
      KernelTimer kt__("lovasz_radix_sort(cub)", stream);
      size_t tmp = w.sort_tmp_bytes;
      const unsigned long long* kin = w.keys;
      C3D_CUDA(cub::DeviceRadixSort::SortKeysDescending(w.sort_tmp, tmp, kin, w.keys2, C * cap, 0, 64,
                                                        stream));
    }
    {
      const int n_chunks = (cap + kLovChunk - 1) / kLovChunk;
      const dim3 cgrid(n_chunks, C);
      { KernelTimer kt__("lovasz_chunk_count_kernel", stream);
        lovasz_chunk_count_kernel<<<cgrid, 1024, 0, stream>>>(C, cap, n_chunks, classes_all, cls_mask, w.hist, w.info,
                                                              w.keys2, w.chunk_fg); }
      if ((rc = check_launch("lovasz_chunk_count_kernel"))) return rc;
      { KernelTimer kt__("lovasz_chunk_apply_kernel", stream);
        lovasz_chunk_apply_kernel<<<cgrid, 1024, 0, stream>>>(C, cap, n_chunks, classes_all, cls_mask, w.hist, w.info,
                                                              w.keys2, w.chunk_fg, w.gval, w.gpix, w.chunk_loss); }
      if ((rc = check_launch("lovasz_chunk_apply_kernel"))) return rc;
      { KernelTimer kt__("lovasz_chunk_reduce_kernel", stream);
        lovasz_chunk_reduce_kernel<<<C, 32, 0, stream>>>(n_chunks, w.chunk_loss, w.cls_loss); }
      if ((rc = check_launch("lovasz_chunk_reduce_kernel"))) return rc;
    }
    KernelTimer kt__("lovasz_finalize_kernel", stream);
    lovasz_finalize_kernel<<<1, 32, 0, stream>>>(C, cap, classes_all, cls_mask, w.hist, w.info, w.cls_loss, loss_out);
    return check_launch("lovasz_finalize_kernel");
  }
  {
    int n = 1024;
    while (n < cap && n < kLovSortMax) n <<= 1;
    const size_t smem = (size_t)n * sizeof(unsigned long long);
    // per call: the attribute is per device, and a process may drive several devices
    C3D_CUDA(cudaFuncSetAttribute(lovasz_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kLovSortMax * (int)sizeof(unsigned long long)));
    KernelTimer kt__("lovasz_sort_kernel", stream);
    lovasz_sort_kernel<<<C, 1024, smem, stream>>>(probs, HW, C, cap, classes_all, cls_mask, w.pix, w.lab, w.hist,
                                                  w.info, w.gval, w.gpix, w.cls_loss);
  }
  if ((rc = check_launch("lovasz_sort_kernel"))) return rc;
  if (cap > kLovSortMax) {  // more keys than shared memory holds: all-pairs ranks
    const dim3 grid((cap + 255) / 256, C);
    {
      KernelTimer kt__("lovasz_keys_kernel", stream);
      lovasz_keys_kernel<<<grid, 256, 0, stream>>>(probs, HW, C, cap, classes_all, cls_mask, w.pix, w.lab, w.hist,
                                                   w.info, w.keys);
    }
    if ((rc = check_launch("lovasz_keys_kernel"))) return rc;
    {
      KernelTimer kt__("lovasz_rank_kernel", stream);
      lovasz_rank_kernel<<<grid, 256, 0, stream>>>(C, cap, classes_all, cls_mask, w.hist, w.info, w.keys, w.term,
                                                   w.gval, w.gpix);
    }
    if ((rc = check_launch("lovasz_rank_kernel"))) return rc;
    {
      KernelTimer kt__("lovasz_reduce_kernel", stream);
      lovasz_reduce_kernel<<<C, 1024, 0, stream>>>(C, cap, classes_all, cls_mask, w.hist, w.info, w.term, w.cls_loss);
    }
    if ((rc = check_launch("lovasz_reduce_kernel"))) return rc;
  }
  {
    KernelTimer kt__("lovasz_finalize_kernel", stream);
    lovasz_finalize_kernel<<<1, 32, 0, stream>>>(C, cap, classes_all, cls_mask, w.hist, w.info, w.cls_loss, loss_out);
  }
  return check_launch("lovasz_finalize_kernel");
}

extern "C" int c3d_lovasz_backward(int batch, int n_classes, int proj_h, int proj_w, int classes_all,
                                   uint64_t class_mask, int64_t max_valid, void* workspace, const float* grad_out,
                                   float* grad_probs, int grad_is_zeroed, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned long long cls_mask = class_mask;
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch && n_classes >= 1 && n_classes <= kLovMaxClasses,
              "bad batch / n_classes");
  C3D_REQUIRE(HWll > 0 && batch * HWll < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(max_valid > 0 && max_valid <= kLovMaxValid, "bad max_valid");
  C3D_REQUIRE(workspace && grad_out && grad_probs, "null pointer argument");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(grad_probs) & 15) == 0, "grad_probs must be 16 B aligned");
  const int HW = (int)HWll, C = n_classes, cap = (int)max_valid;
  LovWs w = carve_lov(workspace, C, cap);
  int rc;
  if (!grad_is_zeroed && (rc = launch_fill(grad_probs, (size_t)batch * C * HW * 4, stream))) return rc;
  const dim3 grid((cap + 255) / 256, C);
  KernelTimer kt__("lovasz_scatter_kernel", stream);
  lovasz_scatter_kernel<<<grid, 256, 0, stream>>>(HW, C, cap, classes_all, cls_mask, w.hist, w.info, w.gval, w.gpix,
                                                  grad_out, grad_probs);
  return check_launch("lovasz_scatter_kernel");
}

extern "C" int c3d_lovasz_info(const void* workspace, int32_t* host_info4, void* stream_) {
  // Synchronous helper: {valid pixels P, classes averaged, flags (1 = more than max_valid
  // labelled pixels, 2 = none), 0}.
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(workspace && host_info4, "null pointer argument");
  C3D_CUDA(cudaMemcpyAsync(host_info4, workspace, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  C3D_CUDA(cudaStreamSynchronize(stream));
  return C3D_OK;
}
