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
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" :: "n"(N) : "memory");
}

// kInflight: committed bulk groups allowed in flight per CTA before the issuing thread waits.
// kWait: 0 = wait for write completion of all but kInflight groups, 1 = wait only until their
// source page has been read (the page is constant, so this bounds nothing but the queue of
// un-started copies), 2 = no wait before the end.
template <int kInflight, int kWait>
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
    if (kWait == 0) bulk_wait_group<kInflight>();
    else if (kWait == 1) bulk_wait_group_read<kInflight>();
  }
  bulk_wait_group<0>();   // writes complete before the CTA (and its zero page) retires
}

// Placement-proof variant.  When the daemon is launched together with other grids (parallel
// branches of a CUDA graph), the block scheduler packs its tiny CTAs into whatever SM slots
// free up first -- several per SM on a few SMs -- and the fill crawls.  Here more CTAs than SMs
// are launched, each claims its SM (%smid) and retires at once if the SM already has
// `max_per_sm` daemon CTAs; pages are handed out by a global counter, so it does not matter
// which SMs ended up with a daemon.  ctrl: [0] next page, [1 + smid] claims (zeroed by the
// launcher); dbg (optional): per CTA {smid, first ns, last ns, pages}.
template <int kInflight>
__global__ void __launch_bounds__(32)
fill_daemon_claim_kernel(char* __restrict__ dst, unsigned long long nbytes, unsigned page_bytes,
                         int max_per_sm, unsigned chunk_pages, unsigned int* __restrict__ ctrl,
                         long long* __restrict__ dbg) {
  extern __shared__ __align__(128) char s_page[];
  const int lane = threadIdx.x;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  int claim = 0;
  if (lane == 0) claim = (int)atomicAdd(&ctrl[1 + smid], 1u);
  claim = __shfl_sync(0xffffffffu, claim, 0);
  if (claim >= max_per_sm) {
    if (dbg && lane == 0) { dbg[(size_t)blockIdx.x * 4] = smid; dbg[(size_t)blockIdx.x * 4 + 3] = -1; }
    return;
  }
  for (unsigned i = lane; i < page_bytes / 16; i += 32)
    reinterpret_cast<float4*>(s_page)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncwarp();
  if (lane != 0) return;
  unsigned long long pol, t0 = 0, t1 = 0;
  if (dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  const unsigned long long npages = (nbytes + page_bytes - 1) / page_bytes;
  long long done = 0;
  for (;;) {
    const unsigned long long p0 = atomicAdd(&ctrl[0], chunk_pages);
    if (p0 >= npages) break;
    const unsigned long long p1 = (p0 + chunk_pages < npages) ? p0 + chunk_pages : npages;
    for (unsigned long long p = p0; p < p1; ++p) {
      const unsigned long long off = p * page_bytes;
      const unsigned long long left = nbytes - off;
      const unsigned nb = left < page_bytes ? (unsigned)left : page_bytes;
      bulk_store_ef(dst + off, s_page, nb, pol);
      ++done;
    }
    bulk_commit_group();
    bulk_wait_group_read<kInflight>();
  }
  bulk_wait_group<0>();
  if (dbg) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    long long* d = dbg + (size_t)blockIdx.x * 4;
    d[0] = smid; d[1] = (long long)t0; d[2] = (long long)t1; d[3] = done;
  }
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

__global__ void delay_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

}  // namespace
}  // namespace c3d

using namespace c3d;

// Holds the stream for `ns` nanoseconds (one spinning thread): lets a kernel launched just
// before on another stream become resident on an otherwise idle GPU.
extern "C" int c3d_delay(unsigned long long ns, void* stream_) {
  delay_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(ns);
  return check_launch("delay_kernel");
}

extern "C" int c3d_zero_fill_daemon(void* dst, size_t nbytes, int max_per_sm, int launch_per_sm,
                                    int page_bytes, int chunk_pages, void* ctrl_ws, void* debug,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(dst && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16 B aligned");
  C3D_REQUIRE(nbytes % 16 == 0, "nbytes must be a multiple of 16");
  C3D_REQUIRE(ctrl_ws && (reinterpret_cast<uintptr_t>(ctrl_ws) & 3) == 0, "ctrl_ws: 1 KB of device scratch");
  C3D_REQUIRE(max_per_sm >= 1 && max_per_sm <= 8 && launch_per_sm >= max_per_sm && launch_per_sm <= 128,
              "need 1 <= max_per_sm <= 8 and max_per_sm <= launch_per_sm <= 128");
  C3D_REQUIRE(page_bytes >= 1024 && page_bytes <= 48 * 1024 && page_bytes % 1024 == 0,
              "page_bytes must be a multiple of 1024 in [1024, 49152]");
  C3D_REQUIRE(chunk_pages >= 1 && chunk_pages <= 64, "chunk_pages must be in [1, 64]");
  if (nbytes == 0) return C3D_OK;
  C3D_CUDA(cudaMemsetAsync(ctrl_ws, 0, 1024, stream));
  KernelTimer kt__("fill_daemon_kernel", stream);
  fill_daemon_claim_kernel<8><<<kNumSMs * launch_per_sm, 32, page_bytes, stream>>>(
      reinterpret_cast<char*>(dst), nbytes, (unsigned)page_bytes, max_per_sm, (unsigned)chunk_pages,
      reinterpret_cast<unsigned int*>(ctrl_ws), reinterpret_cast<long long*>(debug));
  return check_launch("fill_daemon_kernel");
}

extern "C" int c3d_zero_fill_background(void* dst, size_t nbytes, int mode, int ctas_per_sm,
                                        int page_bytes, int inflight, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  C3D_REQUIRE(dst && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16 B aligned");
  C3D_REQUIRE(nbytes % 16 == 0, "nbytes must be a multiple of 16");
  const int wait_mode = mode / 10;   // TMA: 0 completion waits, 1 source-read waits, 2 no waits
  mode %= 10;
  C3D_REQUIRE((mode == 0 || mode == 1) && wait_mode >= 0 && wait_mode <= 2,
              "mode must be 0 (TMA bulk stores; +10 / +20 selects the wait form) or 1 (128-bit stores)");
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
#define C3D_DAEMON_W(N, W)                                                                       \
  do {                                                                                         \
    if (page_bytes > 48 * 1024)                                                                \
      C3D_CUDA(cudaFuncSetAttribute(fill_daemon_tma_kernel<N, W>,                              \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, page_bytes)); \
    KernelTimer kt__("fill_daemon_kernel", stream);                                            \
    fill_daemon_tma_kernel<N, W><<<grid, 32, page_bytes, stream>>>(d, nbytes, pb);             \
  } while (0)
#define C3D_DAEMON(N)                                    \
  do {                                                   \
    if (wait_mode == 0) C3D_DAEMON_W(N, 0);              \
    else if (wait_mode == 1) C3D_DAEMON_W(N, 1);         \
    else C3D_DAEMON_W(N, 2);                             \
  } while (0)
  switch (inflight) {
    case 1: C3D_DAEMON(1); break;
    case 2: C3D_DAEMON(2); break;
    case 4: C3D_DAEMON(4); break;
    case 8: C3D_DAEMON(8); break;
    default: C3D_DAEMON(16); break;
  }
#undef C3D_DAEMON
#undef C3D_DAEMON_W
  return check_launch("fill_daemon_kernel");
}
