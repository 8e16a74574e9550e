// a4 -- KNN range-image label vote for a CSR batch of scans.
//
// Replaces KNN.forward (reference pc_processor/postproc/knn.py:54-142) without
// materialising the two (1, S*S, H*W) unfolds or the three (1, S*S, P) gathers.
//
// One thread per point.  The S x S window of the range image is read through the
// read-only path (<= 0.5 MB per scan, L1/L2 resident) into registers.  Selection
// of the k nearest slots by (distance, slot) -- the tie rule fixed by the oracle --
// is done without tracking indices:
//   1. the k smallest VALUES are kept by a min/max insertion network (2 ALU
//      instructions per compare-exchange, no index bookkeeping);
//   2. t = k-th smallest value; a slot is selected iff d < t, or d == t and it is
//      among the first (k - #{d < t}) such slots in slot order;
//   3. only the selected slots within the cut-off fetch their class and vote; the
//      vote is an O(k^2) register comparison (no (C+1)-wide one-hot).
// This is ~2.3x fewer instructions than an index-tracking selection; the kernel is
// ALU-pipe bound, not memory bound (profiles/).
//
// Compiled with -fmad=false (|a-b| * w must round like torch's separate ops).
//
// Co-scheduled zero fill (kFill).  The vote is ALU-pipe bound and leaves HBM idle; the
// dense-gradient zero fill of the loss backward (98 % of the step's bytes) is the opposite.
// Run as two kernels they do not overlap (the fill's CTAs occupy every SM slot first), so
// the caller may hand the fill to this kernel.  Two forms (kFill):
//   1  every thread issues 128-bit streaming stores between the window rows.  Measured: no
//      gain (132 us fused vs 62 + 81 us apart at B=8) -- the vote's gathers and the fill's
//      stores share the LSU / L1TEX path, which is what both kernels are really bound by.
//   2  TMA: each CTA keeps an 8 KB zero page in shared memory and ONE thread issues
//      cp.async.bulk shared->global copies of it (UBLKCP) over the CTA's share of the buffer.
//      The copy engine streams the page to L2 without LSU wavefronts or issue slots, so the
//      vote keeps the LSU and the ALU while the fill keeps HBM busy (117 us).
//   3  (default) the same with an L2 evict-first policy on the copies, so the zero lines do
//      not displace the range / class images the vote gathers from (113 us; step -2 %).
#include <math_constants.h>

#include "common.cuh"

namespace c3d {

constexpr int kZeroPage = 8192;  // bytes of the shared-memory zero page (TMA fill source)

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

// KT >= knn is the compile-time capacity of the top-k network (KT == knn for knn <= 8).
template <int S, int KT, int kFill>
__global__ void __launch_bounds__(256, (S <= 5 && KT <= 8) ? 5 : 1)
knn_vote_kernel(const float* __restrict__ proj_range, const void* __restrict__ proj_argmax,
                const float* __restrict__ unproj_range, const void* __restrict__ px_,
                const void* __restrict__ py_, const int32_t* __restrict__ offsets, int batch,
                int total, int H, int W, int knn, float cutoff, int nclasses,
                const float* __restrict__ inv_gauss, void* __restrict__ out, int pxy64, int lab64,
                int vec_ok, float4* __restrict__ cofill, unsigned long long cofill_n4,
                unsigned long long cofill_per_cta, int fill_early) {
  constexpr int S2 = S * S;
  constexpr int PAD = (S - 1) / 2;
  extern __shared__ int32_t s_off[];
  __shared__ float s_w[S2];
  __shared__ int s_b0;
  __shared__ __align__(128) float4 s_zero[kFill >= 2 ? kZeroPage / 16 : 1];
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  for (int i = threadIdx.x; i < S2; i += blockDim.x) s_w[i] = inv_gauss[i];
  if (kFill >= 2) {
    for (int i = threadIdx.x; i < kZeroPage / 16; i += blockDim.x) s_zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // page visible to the copy engine
  }
  __syncthreads();
  if (threadIdx.x == 0) s_b0 = scan_of(s_off, batch, min(blockIdx.x * blockDim.x, total - 1));
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  // this CTA's share of the co-scheduled fill, in 16-byte units
  float4* fbase = nullptr;
  int f_cnt = 0;        // kFill 1: stores left to this thread; kFill 2: pages of this CTA
  unsigned f_last = 0;  // kFill 2: bytes of the CTA's last page
  if (kFill) {
    const unsigned long long lo = blockIdx.x * cofill_per_cta;
    const unsigned long long hi = min(lo + cofill_per_cta, cofill_n4);
    if (kFill == 1) {
      fbase = cofill + lo + threadIdx.x;
      f_cnt = lo + threadIdx.x < hi ? (int)((hi - lo - threadIdx.x + 255) / 256) : 0;
    } else if (lo < hi) {  // kFill 2, 3
      fbase = cofill + lo;
      f_cnt = (int)((hi - lo + kZeroPage / 16 - 1) / (kZeroPage / 16));
      f_last = (unsigned)((hi - lo - (unsigned long long)(f_cnt - 1) * (kZeroPage / 16)) * 16);
    }
  }
  auto fill_part = [&](int part, int nparts) {
    if (kFill == 1) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = part; j < f_cnt; j += nparts) __stcs(fbase + (size_t)j * 256, z);
    } else if (kFill >= 2) {
      if (threadIdx.x == 0) {
        for (int j = part; j < f_cnt; j += nparts) {
          const unsigned nb = j == f_cnt - 1 ? f_last : (unsigned)kZeroPage;
          if (kFill == 3) bulk_store_evict_first(fbase + (size_t)j * (kZeroPage / 16), s_zero, nb);
          else bulk_store(fbase + (size_t)j * (kZeroPage / 16), s_zero, nb);
        }
        bulk_commit();
      }
    }
  };
  if (g >= total) { fill_part(0, 1); return; }
  int b = s_b0;
  while (g >= s_off[b + 1]) ++b;
  const int HW = H * W;
  float r;
  int x0, y0, gout = g;
  if (pxy64 == 2) {
    // binned order (c3d_knn_sort_points): one 16-byte record {range, x, y, original index} per
    // point, points of a 32-pixel row segment adjacent, so a warp's window and class loads share
    // cache lines
    const float4 rec = __ldcs(reinterpret_cast<const float4*>(px_) + g);
    r = rec.x; x0 = __float_as_int(rec.y); y0 = __float_as_int(rec.z); gout = __float_as_int(rec.w);
    b = scan_of(s_off, batch, gout);     // records may come in any order: the scan follows the point
  } else if (pxy64) {
    r = __ldcs(unproj_range + g);
    x0 = (int)__ldcs(reinterpret_cast<const long long*>(px_) + g);
    y0 = (int)__ldcs(reinterpret_cast<const long long*>(py_) + g);
  } else {
    r = __ldcs(unproj_range + g);
    x0 = __ldcs(reinterpret_cast<const int*>(px_) + g);
    y0 = __ldcs(reinterpret_cast<const int*>(py_) + g);
  }
  const float* img = proj_range + (size_t)b * HW;

  float d[S2];
  if (vec_ok) {
    // Row segments as aligned 128-bit loads: NQ quads starting at xa = (x0-PAD) & ~3
    // cover the S wanted columns at offset o = (x0-PAD) & 3.  One warp instruction
    // then touches <= 32 sectors for 4 columns instead of 32 sectors per column
    // (the kernel is L2-sector bound with scalar gathers; W % 4 == 0 makes every
    // quad lie fully inside or fully outside the row).
    constexpr int NQ = (S + 3 + 3) / 4;
    const int xs = x0 - PAD;
    const int xa = xs & ~3, o = xs & 3;
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = y0 + dy - PAD;
      const bool rowok = (y >= 0) && (y < H);
      const float* rowp = img + y * W + xa;
      float w[NQ * 4 + 4];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // F.unfold zero padding (knn.py:79-81)
        const int xq = xa + 4 * q;
        if (rowok && xq >= 0 && xq < W) v = __ldg(reinterpret_cast<const float4*>(rowp + 4 * q));
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      if (fill_early == 1) { if (dy == 0) fill_part(0, 1); } else fill_part(dy, fill_early == 2 ? S + 3 : S);
#pragma unroll
      for (int i = NQ * 4; i < NQ * 4 + 4; ++i) w[i] = 0.f;
      // v[dx] = w[o + dx], o in 0..3: two-level select
      float a[NQ * 4 + 2];
#pragma unroll
      for (int i = 0; i < NQ * 4 + 2; ++i) a[i] = (o & 1) ? w[i + 1] : w[i];
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        float v = (o & 2) ? a[dx + 2] : a[dx];
        if (v < 0.0f) v = CUDART_INF_F;              // knn.py:90
        if (dy == PAD && dx == PAD) v = r;           // knn.py:93-94
        d[dy * S + dx] = fabsf(v - r) * s_w[dy * S + dx];  // knn.py:97,107
      }
    }
  } else {
    bool colok[S];
#pragma unroll
    for (int dx = 0; dx < S; ++dx) { const int x = x0 + dx - PAD; colok[dx] = (x >= 0) && (x < W); }
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int y = y0 + dy - PAD;
      const bool rowok = (y >= 0) && (y < H);
      const float* rowp = img + y * W + (x0 - PAD);
      if (fill_early == 1) { if (dy == 0) fill_part(0, 1); } else fill_part(dy, fill_early == 2 ? S + 3 : S);
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        float v = 0.0f;  // F.unfold zero padding (knn.py:79-81)
        if (rowok && colok[dx]) v = __ldg(rowp + dx);
        if (v < 0.0f) v = CUDART_INF_F;              // knn.py:90
        if (dy == PAD && dx == PAD) v = r;           // knn.py:93-94
        d[dy * S + dx] = fabsf(v - r) * s_w[dy * S + dx];  // knn.py:97,107
      }
    }
  }

  // 1. k smallest values (ascending) -- min/max insertion, values only
  float top[KT];
#pragma unroll
  for (int i = 0; i < KT; ++i) top[i] = CUDART_INF_F;
#pragma unroll
  for (int s = 0; s < S2; ++s) {
    float c = d[s];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const float lo = fminf(top[i], c);
      c = fmaxf(top[i], c);
      top[i] = lo;
    }
  }
  if (fill_early == 2) fill_part(S, S + 3);      // spread: three more parts in the later phases
  float t = top[KT - 1];
  if (KT > 8) {  // generic capacity: pick entry knn-1
#pragma unroll
    for (int i = 0; i < KT; ++i) if (i == knn - 1) t = top[i];
  }

  // 2./3. voters = selected slots within the cut-off.
  // selected = {d < t} plus the first (k - #{d < t}) slots with d == t.  If the cut-off
  // is below t, every voter has d <= cutoff < t and is selected; otherwise, when exactly k
  // slots have d <= t (no tie reaches past the k-th), selected = {d <= t}.  Both cases
  // are a single compare per slot; only a real tie at the threshold takes the slow path.
  const bool has_cut = cutoff > 0.0f;                       // knn.py:124-127
  const float tv = (has_cut && cutoff < t) ? cutoff : t;
  int n_le = 0;
  unsigned voters = 0;
#pragma unroll
  for (int s = 0; s < S2; ++s) {
    const bool le = d[s] <= tv;
    n_le += le ? 1 : 0;
    if (S2 <= 32) voters |= le ? (1u << s) : 0u;
  }
  const bool simple = (has_cut && cutoff < t) || (n_le == knn);
  if (S2 > 32 && simple) {
#pragma unroll
    for (int s = 0; s < S2; ++s) if (d[s] <= tv) d[s] = -1.0f;  // large windows: mark in place
  }
  if (!simple) {
    int n_lt = 0;
#pragma unroll
    for (int s = 0; s < S2; ++s) n_lt += (d[s] < t) ? 1 : 0;
    int need = knn - n_lt;
    voters = 0;
#pragma unroll
    for (int s = 0; s < S2; ++s) {
      const bool lt = d[s] < t;
      const bool eq = d[s] == t;
      const bool take = lt || (eq && need > 0);
      if (eq) --need;
      const bool in_cut = !has_cut || !(d[s] > cutoff);
      if (take && in_cut) voters |= (S2 <= 32) ? (1u << s) : 0u;
      if (S2 > 32 && take && in_cut) d[s] = -1.0f;
    }
  }

  if (fill_early == 2) fill_part(S + 1, S + 3);
  // class of each voter (zero padding => class 0, which can never win)
  int cls[KT];
  const char* cbase = reinterpret_cast<const char*>(proj_argmax);
  if (S2 <= 32) {
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      cls[j] = 0;
      if (voters) {
        const int s = __ffs(voters) - 1;
        voters &= voters - 1;
        const int y = y0 + s / S - PAD, x = x0 + s % S - PAD;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          const size_t o = (size_t)b * HW + y * W + x;
          cls[j] = (lab64 & 1) ? (int)__ldg(reinterpret_cast<const long long*>(cbase) + o)
                         : __ldg(reinterpret_cast<const int*>(cbase) + o);
        }
      }
    }
  } else {
    int j = 0;
#pragma unroll
    for (int i = 0; i < KT; ++i) cls[i] = 0;
#pragma unroll
    for (int s = 0; s < S2; ++s) {
      if (d[s] == -1.0f) {
        const int y = y0 + s / S - PAD, x = x0 + s % S - PAD;
        int c = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          const size_t o = (size_t)b * HW + y * W + x;
          c = (lab64 & 1) ? (int)__ldg(reinterpret_cast<const long long*>(cbase) + o)
                    : __ldg(reinterpret_cast<const int*>(cbase) + o);
        }
#pragma unroll
        for (int i = 0; i < KT; ++i) if (i == j) cls[i] = c;
        ++j;
      }
    }
  }

  if (fill_early == 2) fill_part(S + 2, S + 3);
  // vote over classes 1..C-1, first maximum wins (knn.py:131-137)
  int best_c = 1, best_n = 0;
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const int c = cls[j];
    if (c >= 1 && c < nclasses) {
      int n = 0;
#pragma unroll
      for (int i = 0; i < KT; ++i) n += (cls[i] == c) ? 1 : 0;
      if (n > best_n || (n == best_n && c < best_c)) { best_n = n; best_c = c; }
    }
  }
  if (lab64 & 2) reinterpret_cast<uint8_t*>(out)[gout] = (uint8_t)best_c;   // opt-in compact output
  else if (lab64) reinterpret_cast<long long*>(out)[gout] = best_c;
  else reinterpret_cast<int*>(out)[gout] = best_c;
  // the zero page must outlive the copies that read it (thread 0 is always a valid point)
  if (kFill >= 2 && threadIdx.x == 0) bulk_wait_read_all();
}

template <int S, int KT>
static int launch_knn_sk(const float* proj_range, const void* proj_argmax, const float* unproj_range,
                         const void* px, const void* py, const int32_t* offsets, int batch, int total,
                         int H, int W, int knn, float cutoff, int nclasses, const float* inv_gauss,
                         void* out, int pxy64, int lab64, void* cofill, size_t cofill_bytes,
                         cudaStream_t stream) {
  const int threads = 256;
  const int grid = (total + threads - 1) / threads;  // short CTAs: SM slots free up quickly
  const size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
  const int vec_ok = (W % 4 == 0 && (reinterpret_cast<uintptr_t>(proj_range) & 15) == 0) ? 1 : 0;
  if (cofill && cofill_bytes) {
    const unsigned long long n4 = cofill_bytes / 16;
    unsigned long long per_cta = (n4 + grid - 1) / grid;
    per_cta = (per_cta + 511) / 512 * 512;  // whole 8 KB pages (and 4 KB CTA-wide store rounds)
    // Co-fill form: TMA bulk stores of a zero page with an L2 evict-first policy, spread over
    // S + 3 instalments from the window loads down to the vote (the measured best of the
    // variants described at the top of this file).
    const int fill_early = 2;
    KernelTimer timer("knn_vote_fill_kernel", stream);
    knn_vote_kernel<S, KT, 3><<<grid, threads, smem, stream>>>(
        proj_range, proj_argmax, unproj_range, px, py, offsets, batch, total, H, W, knn, cutoff,
        nclasses, inv_gauss, out, pxy64, lab64, vec_ok, reinterpret_cast<float4*>(cofill), n4, per_cta, fill_early);
    return check_launch("knn_vote_fill_kernel");
  }
  KernelTimer timer("knn_vote_kernel", stream);
  knn_vote_kernel<S, KT, 0><<<grid, threads, smem, stream>>>(
      proj_range, proj_argmax, unproj_range, px, py, offsets, batch, total, H, W, knn, cutoff,
      nclasses, inv_gauss, out, pxy64, lab64, vec_ok, nullptr, 0, 0, 0);
  return check_launch("knn_vote_kernel");
}

template <int S>
static int launch_knn_s(int knn, const float* proj_range, const void* proj_argmax,
                        const float* unproj_range, const void* px, const void* py,
                        const int32_t* offsets, int batch, int total, int H, int W, float cutoff,
                        int nclasses, const float* inv_gauss, void* out, int pxy64, int lab64,
                        void* cofill, size_t cofill_bytes, cudaStream_t stream) {
#define KNN_CALL(KT_)                                                                          \
  return launch_knn_sk<S, KT_>(proj_range, proj_argmax, unproj_range, px, py, offsets, batch,  \
                               total, H, W, knn, cutoff, nclasses, inv_gauss, out, pxy64, lab64, \
                               cofill, cofill_bytes, stream)
  switch (knn) {
    case 1: KNN_CALL(1);
    case 2: KNN_CALL(2);
    case 3: KNN_CALL(3);
    case 4: KNN_CALL(4);
    case 5: KNN_CALL(5);
    case 6: KNN_CALL(6);
    case 7: KNN_CALL(7);
    case 8: KNN_CALL(8);
    default: break;
  }
  if (knn <= 16 && S * S >= 16) KNN_CALL(16);
  constexpr int kCap = (S * S > 32) ? 32 : S * S;
  if (knn <= kCap) KNN_CALL(kCap);
#undef KNN_CALL
  set_error("unsupported knn=%d for search=%d", knn, S);
  return C3D_UNSUPPORTED;
}

// ------------------------------------------------------------ binning ----
// The vote reads a 5 x 5 window and up to k classes per point through the L1: with the points of
// a scan in arbitrary order every lane of a warp touches its own cache lines, and the kernel is
// bound by the L1's tag stage (one line per cycle: 15 gather instructions x 32 lines per warp).
// Binning the points by (row, 32-pixel column segment) first makes the lanes of a warp share
// lines.  Counting sort, three small kernels, per-scan (the order stays scan-major, so the CSR
// offsets still describe it): count + rank (one atomicAdd per point), per-scan exclusive scan of
// the bin counts, scatter of 16-byte records {range, x, y, original index}.  The order inside a
// bin is whatever the atomics gave: it only affects locality, the vote of a point does not
// depend on its position.
__device__ __forceinline__ int knn_bin_of(int x, int y, int nbx) { return y * nbx + (x >> 5); }

template <bool kI64>
__global__ void __launch_bounds__(256)
knn_bin_count_kernel(const void* __restrict__ px_, const void* __restrict__ py_,
                     const int32_t* __restrict__ offsets, int batch, int total, int nbx, int nb,
                     int32_t* __restrict__ cnt, int32_t* __restrict__ rank) {
  extern __shared__ int32_t s_off[];
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  int b = scan_of(s_off, batch, g);
  int x, y;
  if (kI64) { x = (int)reinterpret_cast<const long long*>(px_)[g]; y = (int)reinterpret_cast<const long long*>(py_)[g]; }
  else { x = reinterpret_cast<const int*>(px_)[g]; y = reinterpret_cast<const int*>(py_)[g]; }
  rank[g] = atomicAdd(&cnt[(size_t)b * nb + knn_bin_of(x, y, nbx)], 1);
}

// one CTA per scan: cnt[b*nb + bin] -> first sorted position of the bin (offsets[b] + exclusive prefix)
__global__ void __launch_bounds__(1024)
knn_bin_scan_kernel(const int32_t* __restrict__ offsets, int nb, int32_t* __restrict__ cnt) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* c = cnt + (size_t)b * nb;
  if (threadIdx.x == 0) s_carry = offsets[b];
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? c[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - v;
    if (i < nb) c[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
}

template <bool kI64>
__global__ void __launch_bounds__(256)
knn_bin_scatter_kernel(const float* __restrict__ unproj_range, const void* __restrict__ px_,
                       const void* __restrict__ py_, const int32_t* __restrict__ offsets, int batch,
                       int total, int nbx, int nb, const int32_t* __restrict__ start,
                       const int32_t* __restrict__ rank, float4* __restrict__ rec) {
  extern __shared__ int32_t s_off[];
  for (int i = threadIdx.x; i <= batch; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int b = scan_of(s_off, batch, g);
  int x, y;
  if (kI64) { x = (int)reinterpret_cast<const long long*>(px_)[g]; y = (int)reinterpret_cast<const long long*>(py_)[g]; }
  else { x = reinterpret_cast<const int*>(px_)[g]; y = reinterpret_cast<const int*>(py_)[g]; }
  const int pos = start[(size_t)b * nb + knn_bin_of(x, y, nbx)] + rank[g];
  rec[pos] = make_float4(__ldcs(unproj_range + g), __int_as_float(x), __int_as_float(y), __int_as_float(g));
}

}  // namespace c3d

using namespace c3d;

extern "C" size_t c3d_knn_sort_workspace_bytes(int batch, int64_t total_points, int proj_h, int proj_w) {
  if (batch <= 0 || total_points < 0 || proj_h <= 0 || proj_w <= 0) return 0;
  const size_t nb = (size_t)proj_h * ((proj_w + 31) / 32);
  return (((size_t)batch * nb * 4 + 255) & ~(size_t)255) + (size_t)total_points * 4 + 256;
}

extern "C" int c3d_knn_sort_points(const float* unproj_range, const void* px, const void* py,
                                   const int32_t* offsets, int batch, int64_t total_points, int proj_h,
                                   int proj_w, int pxy_is_i64, void* workspace, void* sorted_records,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size");
  C3D_REQUIRE(offsets && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "workspace must be 256 B aligned");
  if (total_points == 0) return C3D_OK;
  C3D_REQUIRE(unproj_range && px && py && sorted_records &&
              (reinterpret_cast<uintptr_t>(sorted_records) & 15) == 0, "null / misaligned per-point pointer");
  const int total = (int)total_points, nbx = (proj_w + 31) / 32, nb = proj_h * nbx;
  int32_t* cnt = reinterpret_cast<int32_t*>(workspace);
  const size_t cnt_bytes = ((size_t)batch * nb * 4 + 255) & ~(size_t)255;
  int32_t* rank = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + cnt_bytes);
  C3D_CUDA(cudaMemsetAsync(cnt, 0, (size_t)batch * nb * 4, stream));
  const int grid = (total + 255) / 256;
  const size_t smem = (size_t)(batch + 1) * sizeof(int32_t);
  int rc;
  { KernelTimer kt__("knn_bin_count_kernel", stream);
    if (pxy_is_i64) knn_bin_count_kernel<true><<<grid, 256, smem, stream>>>(px, py, offsets, batch, total, nbx, nb, cnt, rank);
    else knn_bin_count_kernel<false><<<grid, 256, smem, stream>>>(px, py, offsets, batch, total, nbx, nb, cnt, rank); }
  if ((rc = check_launch("knn_bin_count_kernel"))) return rc;
  { KernelTimer kt__("knn_bin_scan_kernel", stream);
    knn_bin_scan_kernel<<<batch, 1024, 0, stream>>>(offsets, nb, cnt); }
  if ((rc = check_launch("knn_bin_scan_kernel"))) return rc;
  { KernelTimer kt__("knn_bin_scatter_kernel", stream);
    if (pxy_is_i64) knn_bin_scatter_kernel<true><<<grid, 256, smem, stream>>>(unproj_range, px, py, offsets, batch, total, nbx, nb, cnt, rank, reinterpret_cast<float4*>(sorted_records));
    else knn_bin_scatter_kernel<false><<<grid, 256, smem, stream>>>(unproj_range, px, py, offsets, batch, total, nbx, nb, cnt, rank, reinterpret_cast<float4*>(sorted_records)); }
  return check_launch("knn_bin_scatter_kernel");
}

extern "C" int c3d_knn_batch(const float* proj_range, const void* proj_argmax,
                             const float* unproj_range, const void* px, const void* py,
                             const int32_t* offsets, int batch, int64_t total_points, int proj_h,
                             int proj_w, int knn, int search, float cutoff, int nclasses,
                             const float* inv_gauss, int pxy_is_i64, int label_is_i64,
                             void* out_labels, void* cofill_ptr, size_t cofill_bytes,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(search % 2 == 1, "Nearest neighbor kernel must be odd number");  // knn.py:72-73
  C3D_REQUIRE(batch > 0 && batch <= kMaxBatch, "batch must be in [1, %d]", kMaxBatch);
  C3D_REQUIRE(knn >= 1 && knn <= search * search && knn <= 32, "knn must be in [1, min(search^2, 32)]");
  C3D_REQUIRE(nclasses >= 2, "nclasses must be >= 2");
  C3D_REQUIRE(!(label_is_i64 & 2) || nclasses <= 256, "uint8 output needs nclasses <= 256");
  C3D_REQUIRE(total_points >= 0 && total_points < (1ll << 31), "total_points out of range");
  C3D_REQUIRE(proj_h > 0 && proj_w > 0, "bad image size");
  C3D_REQUIRE(proj_range && proj_argmax && offsets && inv_gauss, "null pointer argument");
  C3D_REQUIRE((cofill_ptr == nullptr) == (cofill_bytes == 0), "cofill_ptr and cofill_bytes go together");
  C3D_REQUIRE((reinterpret_cast<uintptr_t>(cofill_ptr) & 15) == 0 && cofill_bytes % 16 == 0,
              "co-scheduled fill must be 16 B aligned and a multiple of 16 B");
  if (total_points == 0) return cofill_ptr ? launch_fill(cofill_ptr, cofill_bytes, stream) : C3D_OK;
  if (pxy_is_i64 == 2) {   // px = the 16-byte records of c3d_knn_sort_points; unproj_range / py unused
    C3D_REQUIRE(px && out_labels && (reinterpret_cast<uintptr_t>(px) & 15) == 0, "null / misaligned records");
  } else {
    C3D_REQUIRE(unproj_range && px && py && out_labels, "null per-point pointer");
  }
#define KNN_ARGS knn, proj_range, proj_argmax, unproj_range, px, py, offsets, batch,            \
                 (int)total_points, proj_h, proj_w, cutoff, nclasses, inv_gauss, out_labels,     \
                 pxy_is_i64, label_is_i64, cofill_ptr, cofill_bytes, stream
  switch (search) {
    case 3: return launch_knn_s<3>(KNN_ARGS);
    case 5: return launch_knn_s<5>(KNN_ARGS);
    case 7: return launch_knn_s<7>(KNN_ARGS);
    case 9: return launch_knn_s<9>(KNN_ARGS);
    default: break;
  }
#undef KNN_ARGS
  set_error("unsupported KNN window: search=%d (supported: 3, 5, 7, 9)", search);
  return C3D_UNSUPPORTED;
}
