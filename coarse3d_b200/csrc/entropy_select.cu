// f3 -- entropy-based pseudo-label selection for a batch of scans.
//
// Replaces Trainer.entropy_based_selection (reference
// tasks/weak_segmentation/trainer.py:447-518): per (scan, class present in the weak
// labels) a weighted draw WITHOUT replacement of int(count * select_ratio) pixels among
// the pixels predicted as that class, weights exp(-entropy).  The reference loops over
// B x C in Python, each iteration running unique / mask / multinomial(131072 weights) /
// scatter kernels with host synchronisations.
//
// torch.multinomial(replacement=False) draws q ~ Exp(1) per element and takes
// topk(w / q) (aten/native/Distributions.cpp).  Every pixel belongs to at most one class
// (its arg-max), so it has ONE key w / q and the whole selection is:
//   S1 select_prepare    per pixel: entropy, arg-max, key; per-(scan, class) counts
//   S2 select_threshold  per (scan, class): the class's keys compacted into shared memory,
//                        then the k-th largest by a 3-pass radix select (11 + 11 + 10 bits)
//   S3 select_apply      per pixel: selected = key >= threshold; pseudo label, ground
//                        truth kept where the weak mask is set (:512-516)
// Noise: injected (`noise[b][c][pixel]`, the reference's per-iteration draws) for exact
// parity, else Philox on the device.  Ties at the threshold are all selected (torch.topk
// picks arbitrarily among them); with continuous noise they do not occur.
#include <math_constants.h>

#include "common.cuh"

namespace c3d {

constexpr int kSelMaxClasses = 64;
constexpr unsigned kNoThreshold = 0xFFFFFFFFu;

struct SelWs {
  uint8_t* pseudo;    // [B*HW] arg-max class, 255 = not a candidate
  float* key;         // [B*HW]
  int32_t* count;     // [B*C] candidates per (scan, class)
  int32_t* present;   // [B*C] class occurs in train_label[b]
  uint32_t* thr;      // [B*C] key bits of the k-th largest, kNoThreshold = select nothing
  int32_t* flags;     // [1] bit 0: a guard band overflowed (fast threshold kept for that class)
  size_t bytes;
};

static SelWs carve_sel(void* base, int B, int C, int HW) {
  SelWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~(size_t)255; return (char*)base + o; };
  w.count = (int32_t*)take((size_t)B * C * 4);
  w.present = (int32_t*)take((size_t)B * C * 4);
  w.thr = (uint32_t*)take((size_t)B * C * 4);
  w.flags = (int32_t*)take(4);
  w.key = (float*)take((size_t)B * HW * 4);
  w.pseudo = (uint8_t*)take((size_t)B * HW);
  w.bytes = off;
  return w;
}

__device__ __forceinline__ uint4 philox_s(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

// The Exp(1) draw of pixel `gi` on the device sampler's Philox stream.
__device__ __forceinline__ float philox_exp1(unsigned long long gi, unsigned long long seed) {
  const uint4 r = philox_s(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), 2u, 0u),
                           make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return -logf(((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f));
}

// The key of one pixel under the oracle's rule (oracle/entropy_select.py, "rounded"): every
// float32 operation of trainer.py:459-466 correctly rounded, in the reference's order -- log and
// exp evaluated in float64 and rounded once, products and the class sum as separate float32
// operations (no contraction), IEEE division.  Only the few pixels whose fast key lies in the
// guard band around a (scan, class) threshold are evaluated this way (select_threshold).
__device__ __forceinline__ float rounded_rule_key(const float* __restrict__ probs_b, int pix, int HW, int C,
                                                  float q) {
  float ent = 0.f;
  for (int c = 0; c < C; ++c) {
    const float p = __ldg(probs_b + (size_t)c * HW + pix);
    const float x = __fadd_rn(p, 1e-10f);
    const float l = (float)log((double)x);
    ent = __fadd_rn(ent, __fmul_rn(p, l));
  }
  const float w = (float)exp((double)ent);       // exp(-1 * entropy), entropy = -sum
  return __fdiv_rn(w, q);
}

// ---------------------------------------------------------------- S1 -------
// kPx pixels per thread (adjacent, one 64-bit load per class when kPx == 2): twice the bytes
// in flight per thread and half the address arithmetic of the one-pixel form.
template <int kPx>
__global__ void __launch_bounds__(256)
select_prepare_kernel(const float* __restrict__ probs, const long long* __restrict__ train_label,
                      const uint8_t* __restrict__ eval_mask, int HW, int C, int ignore_cls,
                      const float* __restrict__ noise, unsigned long long seed,
                      uint8_t* __restrict__ pseudo, float* __restrict__ key,
                      int32_t* __restrict__ count, int32_t* __restrict__ present) {
  __shared__ int s_cnt[kSelMaxClasses];
  __shared__ int s_pre[kSelMaxClasses];
  const int b = blockIdx.y;
  if (threadIdx.x < C) { s_cnt[threadIdx.x] = 0; s_pre[threadIdx.x] = 0; }
  __syncthreads();
  const int pix0 = (blockIdx.x * blockDim.x + threadIdx.x) * kPx;
  if (pix0 < HW) {
    const float* p = probs + (size_t)b * C * HW + pix0;
    float ent[kPx], best[kPx];
    int arg[kPx];
#pragma unroll
    for (int e = 0; e < kPx; ++e) { ent[e] = 0.f; best[e] = -CUDART_INF_F; arg[e] = 0; }
    for (int c0 = 0; c0 < C; c0 += 8) {  // 8 coalesced loads in flight
      float v[8][kPx];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < C) {
          if (kPx == 2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(p + (size_t)(c0 + j) * HW));
            v[j][0] = t.x; v[j][kPx - 1] = t.y;
          } else {
            v[j][0] = __ldg(p + (size_t)(c0 + j) * HW);
          }
        } else {
#pragma unroll
          for (int e = 0; e < kPx; ++e) v[j][e] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < C) {
#pragma unroll
          for (int e = 0; e < kPx; ++e) {
            ent[e] += v[j][e] * logf(v[j][e] + 1e-10f);                  // trainer.py:459-461
            if (v[j][e] > best[e]) { best[e] = v[j][e]; arg[e] = c0 + j; }  // torch.max: first maximum (:463)
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
      const int pix = pix0 + e;
      const float w = expf(-1.0f * (-ent[e]));                   // :466
      const size_t gi = (size_t)b * HW + pix;
      const bool ev = eval_mask[gi] != 0;
      const bool cand = ev && arg[e] != ignore_cls;              // :469, :477-480
      float k = 0.f;
      if (cand) {
        float q;
        if (noise) q = noise[((size_t)b * C + arg[e]) * HW + pix];
        else q = philox_exp1(gi, seed);                          // Exp(1)
        k = w / q;                                               // multinomial: topk(w / q)
        atomicAdd(&s_cnt[arg[e]], 1);
      }
      pseudo[gi] = cand ? (uint8_t)arg[e] : (uint8_t)255;
      key[gi] = k;
      const long long tl = train_label[gi];
      if (tl >= 0 && tl < C) s_pre[(int)tl] = 1;                 // unique(train_label[b]) (:474)
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    if (s_cnt[threadIdx.x]) atomicAdd(&count[b * C + threadIdx.x], s_cnt[threadIdx.x]);
    if (s_pre[threadIdx.x]) present[b * C + threadIdx.x] = 1;
  }
}

// ---------------------------------------------------------------- S2 -------
// Suffix search over a histogram in shared memory: finds the highest bin `sel` such that
// the number of elements in bins > sel is < k <= that number + hist[sel]; returns the
// count in bins above `sel` through s_out[1] and sel through s_out[0].
template <int NBINS>
__device__ __forceinline__ void find_bin_from_top(const int* s_hist, int k, int* s_scan, int* s_out) {
  constexpr int PER = NBINS / 512;   // bins per thread (512 threads)
  const int t = threadIdx.x;
  // thread t owns bins [NBINS - (t+1)*PER, NBINS - t*PER): thread 0 has the top bins
  int local = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) local += s_hist[NBINS - 1 - (t * PER + j)];
  // inclusive prefix over threads (from the top)
  int incl = local;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  int pre = 0;
  for (int w = 0; w < warp; ++w) pre += s_scan[w];
  incl += pre;
  const int excl = incl - local;
  if (excl < k && k <= incl) {  // the crossing is inside this thread's bins
    int above = excl;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int bin = NBINS - 1 - (t * PER + j);
      const int h = s_hist[bin];
      if (above < k && k <= above + h) { s_out[0] = bin; s_out[1] = above; }
      above += h;
    }
  }
  __syncthreads();
}

// One CTA per (scan, class).  The class's keys are first compacted into shared memory
// (one vectorised pass over the scan's arg-max bytes; a class holds ~HW/C of the pixels), and
// the three radix passes then run from shared memory.  A class with more candidates than the
// buffer holds (kSelSmemKeys) falls back to re-reading global memory in every pass.
//
// Exactness.  The fast keys of S1 carry the device's logf / expf errors (a few 1e-7 relative), so
// a key next to the k-th one could land on the other side of it than under the oracle's
// correctly rounded rule.  After the radix select, every candidate whose fast key lies within
// kSelGuard (relative) of the fast threshold -- the threshold's own pixel and, rarely, a
// neighbour -- is re-evaluated under that rule (rounded_rule_key), its stored key is replaced,
// and the threshold becomes the r-th largest exact key of the band, r = k - #{keys above the
// band}.  Keys outside the band are more than 100 error bounds away from the threshold, so the
// selection of S3 is the oracle's, bit for bit.
constexpr int kSelSmemKeys = 6144;    // 24 KB keys + 24 KB pixel ids + 8 KB histogram: 4 CTAs per SM
constexpr int kSelBandCap = 256;      // band candidates re-evaluated per (scan, class)
constexpr float kSelGuard = 6.1035156e-05f;   // 2^-14 >> the fast keys' error (<= ~1e-5 worst case)
__global__ void __launch_bounds__(512)
select_threshold_kernel(const uint8_t* __restrict__ pseudo, float* __restrict__ key,
                        const int32_t* __restrict__ count, const int32_t* __restrict__ present,
                        int HW, int C, int ignore_cls, float select_ratio,
                        const float* __restrict__ probs, const float* __restrict__ noise,
                        unsigned long long seed, uint32_t* __restrict__ thr, int32_t* __restrict__ flags) {
  __shared__ int s_hist[2048];
  __shared__ int s_scan[16];
  __shared__ int s_out[2];
  __shared__ int s_n, s_above, s_nb;
  __shared__ int s_bpix[kSelBandCap];
  __shared__ float s_bkey[kSelBandCap];
  extern __shared__ uint32_t s_keys[];
  uint32_t* s_pix = s_keys + kSelSmemKeys;
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const int cnt = count[blockIdx.x];
  // select_num = int(cls_mask.sum() * select_ratio): int64 0-dim tensor times a Python
  // float is computed in float32 (:485)
  const int k0 = (int)((float)cnt * select_ratio);
  if (c == ignore_cls || !present[blockIdx.x] || cnt == 0 || k0 < 1) {   // :477-488
    if (threadIdx.x == 0) thr[blockIdx.x] = kNoThreshold;
    return;
  }
  const uint8_t* ps = pseudo + (size_t)b * HW;
  const uint32_t* kb = reinterpret_cast<const uint32_t*>(key) + (size_t)b * HW;
  const bool in_smem = cnt <= kSelSmemKeys;
  if (in_smem) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // 16 arg-max bytes per load (HW % 16 == 0 and a 16 B aligned scan base, else bytewise)
    const bool vec = (HW % 16 == 0) && ((reinterpret_cast<uintptr_t>(ps) & 15) == 0);
    if (vec) {
      const uint4* ps4 = reinterpret_cast<const uint4*>(ps);
      const uint32_t pat = 0x01010101u * (uint32_t)c;
      for (int i = threadIdx.x; i < HW / 16; i += blockDim.x) {
        const uint4 v = __ldg(ps4 + i);
        const uint32_t w[4] = {v.x ^ pat, v.y ^ pat, v.z ^ pat, v.w ^ pat};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // any zero byte in w[q]?  (exact test)
          if (((w[q] - 0x01010101u) & ~w[q] & 0x80808080u) != 0u) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (((w[q] >> (8 * e)) & 0xFFu) == 0u) {
                const int at = atomicAdd(&s_n, 1), pix = i * 16 + q * 4 + e;
                s_keys[at] = kb[pix]; s_pix[at] = pix;
              }
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x)
        if (ps[i] == (uint8_t)c) { const int at = atomicAdd(&s_n, 1); s_keys[at] = kb[i]; s_pix[at] = i; }
    }
    __syncthreads();
  }
  const int n_keys = in_smem ? s_n : HW;
  int k = k0;
  uint32_t prefix = 0, prefix_mask = 0;
  // keys are positive floats: their bit patterns order like unsigned integers
  const int shifts[3] = {21, 10, 0};
  const int nbins[3] = {2048, 2048, 1024};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int sh = shifts[pass];
    const uint32_t bmask = nbins[pass] - 1;
    if (in_smem) {
      for (int i = threadIdx.x; i < n_keys; i += blockDim.x) {
        const uint32_t v = s_keys[i];
        if ((v & prefix_mask) == prefix) atomicAdd(&s_hist[(v >> sh) & bmask], 1);
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        if (ps[i] == (uint8_t)c) {
          const uint32_t v = kb[i];
          if ((v & prefix_mask) == prefix) atomicAdd(&s_hist[(v >> sh) & bmask], 1);
        }
      }
    }
    __syncthreads();
    if (nbins[pass] == 2048) find_bin_from_top<2048>(s_hist, k, s_scan, s_out);
    else find_bin_from_top<1024>(s_hist, k, s_scan, s_out);
    const int bin = s_out[0];
    k -= s_out[1];
    prefix |= (uint32_t)bin << sh;
    prefix_mask |= bmask << sh;
    __syncthreads();
  }
  // ---- guard band around the fast threshold `prefix` (bits of the k0-th largest fast key)
  const float t_fast = __uint_as_float(prefix);
  const float lo = t_fast * (1.0f - kSelGuard), hi = t_fast * (1.0f + kSelGuard);
  if (threadIdx.x == 0) { s_above = 0; s_nb = 0; }
  __syncthreads();
  int above = 0;
  auto visit = [&](uint32_t bits, int pix) {
    const float v = __uint_as_float(bits);
    if (v > hi) ++above;
    else if (v >= lo) { const int at = atomicAdd(&s_nb, 1); if (at < kSelBandCap) s_bpix[at] = pix; }
  };
  if (in_smem) {
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x) visit(s_keys[i], (int)s_pix[i]);
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) if (ps[i] == (uint8_t)c) visit(kb[i], i);
  }
  above = __reduce_add_sync(0xffffffffu, above);
  if ((threadIdx.x & 31) == 0 && above) atomicAdd(&s_above, above);
  __syncthreads();
  const int nb = s_nb, r = k0 - s_above;      // 1 <= r <= nb: the fast threshold is in the band
  if (nb > kSelBandCap) {                     // a pile-up of (near-)equal keys: keep the fast threshold
    if (threadIdx.x == 0) { thr[blockIdx.x] = prefix; atomicOr(flags, 1); }
    return;
  }
  if ((int)threadIdx.x < nb) {
    const int pix = s_bpix[threadIdx.x];
    const size_t gi = (size_t)b * HW + pix;
    const float q = noise ? noise[((size_t)b * C + c) * HW + pix] : philox_exp1(gi, seed);
    const float kx = rounded_rule_key(probs + (size_t)b * C * HW, pix, HW, C, q);
    s_bkey[threadIdx.x] = kx;
    key[gi] = kx;
  }
  __syncthreads();
  if ((int)threadIdx.x < nb) {
    const float mine = s_bkey[threadIdx.x];
    int gt = 0, ge = 0;
    for (int j = 0; j < nb; ++j) { gt += s_bkey[j] > mine; ge += s_bkey[j] >= mine; }
    if (gt < r && r <= ge) thr[blockIdx.x] = __float_as_uint(mine);   // ties write the same bits
  }
}

// ---------------------------------------------------------------- S3 -------
__global__ void __launch_bounds__(256)
select_apply_kernel(const uint8_t* __restrict__ pseudo, const float* __restrict__ key,
                    const uint32_t* __restrict__ thr, const long long* __restrict__ train_label,
                    const uint8_t* __restrict__ wss_mask, int HW, int C, int ignore_cls,
                    long long total, long long* __restrict__ out_label,
                    uint8_t* __restrict__ out_mask) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= total) return;
  const int b = (int)(gi / HW);
  const int c = pseudo[gi];
  long long lab = 0;                                          // pseudo * mask (:512)
  if (c != 255) {
    const uint32_t t = thr[b * C + c];
    if (t != kNoThreshold && __float_as_uint(key[gi]) >= t) lab = c;
  }
  if (wss_mask[gi]) lab = train_label[gi];                    // :515
  out_label[gi] = lab;
  out_mask[gi] = lab != ignore_cls;                           // :516
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_entropy_select_workspace_bytes(int batch, int n_classes, int hw) {
  if (batch <= 0 || n_classes < 1 || hw <= 0) return 0;
  return carve_sel(nullptr, batch, n_classes, hw).bytes;
}

extern "C" int c3d_entropy_select_batch(
    const float* probs, const int64_t* train_label, const uint8_t* wss_mask,
    const uint8_t* eval_mask, int batch, int n_classes, int proj_h, int proj_w, int ignore_cls,
    float select_ratio, const float* noise, uint64_t seed, void* workspace, int64_t* out_label,
    uint8_t* out_mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long HWll = (long long)proj_h * proj_w;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(n_classes >= 1 && n_classes <= kSelMaxClasses, "n_classes must be in [1, %d]", kSelMaxClasses);
  C3D_REQUIRE(HWll > 0 && batch * HWll < (1ll << 31), "batch*H*W must be < 2^31");
  C3D_REQUIRE(probs && train_label && wss_mask && eval_mask && workspace && out_label && out_mask,
              "null pointer argument");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256 B aligned");
  C3D_REQUIRE(select_ratio >= 0.f, "select_ratio must be >= 0");
  const int HW = (int)HWll, B = batch, C = n_classes;
  SelWs w = carve_sel(workspace, B, C, HW);
  C3D_CUDA(cudaMemsetAsync(w.count, 0, (size_t)((char*)w.thr - (char*)w.count), stream));  // count + present
  C3D_CUDA(cudaMemsetAsync(w.flags, 0, 4, stream));
  int rc;
  {
    KernelTimer kt__("select_prepare_kernel", stream);
    if (HW % 2 == 0 && (reinterpret_cast<uintptr_t>(probs) & 7) == 0) {
      dim3 grid((HW / 2 + 255) / 256, B);
      select_prepare_kernel<2><<<grid, 256, 0, stream>>>(probs, (const long long*)train_label, eval_mask,
                                                         HW, C, ignore_cls, noise, seed, w.pseudo, w.key,
                                                         w.count, w.present);
    } else {
      dim3 grid((HW + 255) / 256, B);
      select_prepare_kernel<1><<<grid, 256, 0, stream>>>(probs, (const long long*)train_label, eval_mask,
                                                         HW, C, ignore_cls, noise, seed, w.pseudo, w.key,
                                                         w.count, w.present);
    }
  }
  if ((rc = check_launch("select_prepare_kernel"))) return rc;
  {
    KernelTimer kt__("select_threshold_kernel", stream);
    const size_t smem = (size_t)kSelSmemKeys * 2 * sizeof(uint32_t);   // 48 KB + 11 KB static: opt-in
    C3D_CUDA(cudaFuncSetAttribute(select_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_threshold_kernel<<<B * C, 512, smem, stream>>>(w.pseudo, w.key, w.count, w.present, HW, C,
                                                          ignore_cls, select_ratio, probs, noise, seed, w.thr,
                                                          w.flags);
  }
  if ((rc = check_launch("select_threshold_kernel"))) return rc;
  {
    const long long total = (long long)B * HW;
    KernelTimer kt__("select_apply_kernel", stream);
    select_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        w.pseudo, w.key, w.thr, (const long long*)train_label, wss_mask, HW, C, ignore_cls, total,
        (long long*)out_label, out_mask);
  }
  return check_launch("select_apply_kernel");
}
