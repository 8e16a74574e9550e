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

// Streaming 128-bit store (data written once, never re-read by this kernel).
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

}  // namespace c3d
