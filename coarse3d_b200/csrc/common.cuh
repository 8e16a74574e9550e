// Shared helpers for the coarse3d_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/coarse3d_b200.h"

namespace c3d {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kMaxBatch = 4096;

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// RAII device timer around one kernel launch; a no-op unless c3d_profile_enable.
class KernelTimer {
 public:
  KernelTimer(const char* name, cudaStream_t stream);
  ~KernelTimer();
 private:
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t a_, b_;
};

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return C3D_CUDA_ERROR;
  }
  return C3D_OK;
}

#define C3D_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::c3d::set_error(__VA_ARGS__);      \
      return C3D_INVALID_ARGUMENT;        \
    }                                     \
  } while (0)

#define C3D_CUDA(call)                                                   \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) {                                            \
      ::c3d::set_error("%s: %s", #call, cudaGetErrorString(e__));        \
      return C3D_CUDA_ERROR;                                             \
    }                                                                    \
  } while (0)

// Zero fill with 128-bit streaming stores (proto_loss.cu); shared with the KNN entry point.
int launch_fill(void* dst, size_t nbytes, cudaStream_t stream);

// Grid for a grid-stride kernel: whole waves of the 148 SMs.
inline int wave_grid(long long work_items, int threads, int ctas_per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  long long wave = (long long)kNumSMs * ctas_per_sm;
  if (blocks >= wave) return (int)wave;
  return (int)(blocks > 0 ? blocks : 1);
}

// Scan id of flat point index g: largest b with offsets[b] <= g.
__device__ __forceinline__ int scan_of(const int32_t* __restrict__ offs, int batch, int g) {
  int lo = 0, hi = batch;  // offs[lo] <= g < offs[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (offs[mid] <= g) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 16-byte asynchronous global -> shared copy (LDGSTS), bypassing registers and L1.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------ carried fill ----
// The dense-gradient zero fill of the loss backward (84 % of a step's bytes, pure HBM writes)
// has no dependence on anything but "done before the gradient rows are scattered", while every
// other kernel of the step is latency-, ALU- or L1-bound and leaves HBM mostly idle.  Any kernel
// can therefore CARRY a share of the fill: each CTA keeps a small zero page in shared memory and
// ONE thread issues cp.async.bulk shared->global copies of it (UBLKCP: no LSU wavefronts, a few
// issue slots) over the CTA's slice of the share, in instalments between its own phases.
struct FillShare {
  char* ptr;                    // this launch's byte range of the buffer to zero (16 B aligned), or null
  unsigned long long bytes;     // multiple of 16
};

__device__ __forceinline__ void bulk_store_evict_first(void* gdst, const void* ssrc, unsigned bytes) {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n"
               :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

// All threads of the CTA: zero the page and make it visible to the copy engine.  The caller's
// next __syncthreads() (before the first carrier_issue) completes the hand-over.
__device__ __forceinline__ void carrier_init(void* s_page, int page_bytes) {
  for (int i = threadIdx.x; i < page_bytes / 16; i += blockDim.x)
    reinterpret_cast<float4*>(s_page)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// One thread: the pages j = part, part + nparts, ... of CTA `cta`'s slice (of `nctas` equal,
// page-aligned slices of the share).
__device__ __forceinline__ void carrier_issue(const FillShare& f, const void* s_page, unsigned page_bytes,
                                              unsigned cta, unsigned nctas, int part, int nparts) {
  if (f.ptr == nullptr || f.bytes == 0) return;
  const unsigned long long pages = (f.bytes + page_bytes - 1) / page_bytes;
  const unsigned long long per = (pages + nctas - 1) / nctas;
  const unsigned long long lo = (unsigned long long)cta * per;
  const unsigned long long hi = lo + per < pages ? lo + per : pages;
  bool any = false;
  for (unsigned long long j = lo + part; j < hi; j += nparts) {
    const unsigned long long off = j * page_bytes;
    const unsigned long long left = f.bytes - off;
    bulk_store_evict_first(f.ptr + off, s_page, left < page_bytes ? (unsigned)left : page_bytes);
    any = true;
  }
  if (any) bulk_commit();
}

// A carrier WARP: the rows kernels (one 256-thread CTA per SM, phases separated by barriers) give
// the fill a ninth warp of its own instead of thread 0 -- a thread that issues hundreds of bulk
// copies blocks whenever the copy queue is full, and the whole CTA would wait for it at the next
// barrier.  The compute warps synchronise among themselves on named barrier 1 (rows_sync).
constexpr int kRowsThreads = 256;
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void carrier_warp_run(const FillShare& f, void* s_page, int page_bytes,
                                                 unsigned cta, unsigned nctas) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < page_bytes / 16; i += 32)
    reinterpret_cast<float4*>(s_page)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    carrier_issue(f, s_page, page_bytes, cta, nctas, 0, 1);
    bulk_wait_read_all();
  }
}

// Streaming 128-bit store (data written once, never re-read by this kernel).
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

}  // namespace c3d
