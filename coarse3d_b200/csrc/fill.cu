// Background zero fill ("fill daemon") for the dense (B,D,H,W) gradient of the prototype
// loss (the reference's autograd zero-fills it inside index_put's backward,
// pc_processor/loss/contrast_pixel_loss.py:118-123 -> grad of `X[b, keep_indices]`).
//
// The fill is 84 % of the step's compulsory bytes and pure HBM-write work; every other
// kernel of the step is latency- or ALU-bound.  Instead of a grid that occupies every SM
// slot (and starves the small kernels of the other chains), the fill runs as a persistent
// kernel with a minimal footprint -- ONE warp per CTA, one CTA per SM, a few KB of shared
// memory -- launched first and resident for the whole step:
//   mode 0  (TMA)  one thread per CTA issues cp.async.bulk shared->global copies of a
//           zero page (UBLKCP): no LSU wavefronts, no issue slots to speak of;
//   mode 1  (STG)  the warp issues 128-bit streaming stores (512 B per instruction).
// The number of copies in flight per CTA bounds how hard the fill leans on HBM/L2 while
// the latency-bound kernels of the other streams run.
#include "common.cuh"

namespace c3d {

namespace {

__device__ __forceinline__ void bulk_store_ef(void* gdst, const void* ssrc, unsigned bytes,
                                              unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n"
               :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;\n" :: "n"(N) : "memory");
}

// kInflight: committed bulk groups allowed in flight per CTA before the issuing thread waits.
template <int kInflight>
__global__ void __launch_bounds__(32)
fill_daemon_tma_kernel(char* __restrict__ dst, unsigned long long nbytes, unsigned page_bytes) {
  extern __shared__ __align__(128) char s_page[];
  const int lane = threadIdx.x;
  for (unsigned i = lane; i < page_bytes / 16; i += 32)
    reinterpret_cast<float4*>(s_page)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // page visible to the copy engine
  __syncwarp();
  if (lane != 0) return;
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  const unsigned long long npages = (nbytes + page_bytes - 1) / page_bytes;
  // pages are dealt round-robin: at any moment the CTAs write one contiguous window
  for (unsigned long long p = blockIdx.x; p < npages; p += gridDim.x) {
    const unsigned long long off = p * page_bytes;
    const unsigned long long left = nbytes - off;
    const unsigned nb = left < page_bytes ? (unsigned)left : page_bytes;
    bulk_store_ef(dst + off, s_page, nb, pol);
    bulk_commit_group();
    bulk_wait_group<kInflight>();
  }
  bulk_wait_group<0>();   // writes complete before the CTA (and its zero page) retires
}

__global__ void __launch_bounds__(32)
fill_daemon_stg_kernel(float4* __restrict__ dst, unsigned long long n4) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr unsigned kChunk = 32 * 8;   // float4 per CTA round: 8 independent 512 B stores
  for (unsigned long long base = (unsigned long long)blockIdx.x * kChunk; base < n4;
       base += (unsigned long long)gridDim.x * kChunk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned long long i = base + j * 32 + threadIdx.x;
      if (i < n4) __stcs(dst + i, z);
    }
  }
}

}  // namespace
}  // namespace c3d

using namespace c3d;

extern "C" int c3d_zero_fill_background(void* dst, size_t nbytes, int mode, int ctas_per_sm,
                                        int page_bytes, int inflight, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(dst && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16 B aligned");
  C3D_REQUIRE(nbytes % 16 == 0, "nbytes must be a multiple of 16");
  C3D_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (TMA bulk stores) or 1 (128-bit stores)");
  C3D_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 8, "ctas_per_sm must be in [1, 8]");
  if (nbytes == 0) return C3D_OK;
  const int grid = kNumSMs * ctas_per_sm;
  if (mode == 1) {
    KernelTimer kt__("fill_daemon_kernel", stream);
    fill_daemon_stg_kernel<<<grid, 32, 0, stream>>>(reinterpret_cast<float4*>(dst), nbytes / 16);
    return check_launch("fill_daemon_kernel");
  }
  C3D_REQUIRE(page_bytes >= 1024 && page_bytes <= 64 * 1024 && page_bytes % 1024 == 0,
              "page_bytes must be a multiple of 1024 in [1024, 65536]");
  C3D_REQUIRE(inflight == 1 || inflight == 2 || inflight == 4 || inflight == 8 || inflight == 16,
              "inflight must be 1, 2, 4, 8 or 16");
  char* d = reinterpret_cast<char*>(dst);
  const unsigned pb = (unsigned)page_bytes;
#define C3D_DAEMON(N)                                                                          \
  do {                                                                                         \
    if (page_bytes > 48 * 1024)                                                                \
      C3D_CUDA(cudaFuncSetAttribute(fill_daemon_tma_kernel<N>,                                 \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, page_bytes)); \
    KernelTimer kt__("fill_daemon_kernel", stream);                                            \
    fill_daemon_tma_kernel<N><<<grid, 32, page_bytes, stream>>>(d, nbytes, pb);                \
  } while (0)
  switch (inflight) {
    case 1: C3D_DAEMON(1); break;
    case 2: C3D_DAEMON(2); break;
    case 4: C3D_DAEMON(4); break;
    case 8: C3D_DAEMON(8); break;
    default: C3D_DAEMON(16); break;
  }
#undef C3D_DAEMON
  return check_launch("fill_daemon_kernel");
}
