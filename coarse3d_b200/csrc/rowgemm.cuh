// Register-tiled SIMT "rows x bank" products shared by the prototype loss and the
// EMA update.
//
// A CTA of 256 threads owns a group of 16 feature rows (A, 16 x D, shared memory) and
// multiplies it with a tile of bank rows staged in shared memory:
//
//   tile_logits : C[16][rows] = A . bank_tile^T      thread (rg, cg) -> 4 rows x 6 cols
//   tile_gradT  : dA[16][D]  += G[16][rows] . bank_tile   thread (rgp, ct) -> kRP rows x 4 d
//
// The first version of these kernels gave each warp ONE row and re-read every bank
// element from shared memory once per row: ncu showed them shared-memory-bandwidth
// bound (1920 wavefronts per row per product, L1TEX busiest unit, issue 28 %).
// Here every bank element read is reused for 4 rows from registers, and A / G reads
// are warp-wide broadcasts, so the products become FMA-bound (~7 k cycles per 16 rows
// at D=128, K=380 instead of ~31 k per 8 rows).
//
// Bank tile layout: row r, 16-byte chunk c.  If D % 32 == 0 the chunk is stored at
// c ^ (r & 7) (XOR swizzle, row stride D): 8 consecutive rows at one logical chunk hit
// 8 different 16 B bank groups, so the 128-bit loads of `tile_logits` (lanes on
// consecutive rows) are conflict-free without padding -- padding would push the
// KITTI-shaped bank (380 x 128 f32) past the 227 KB shared-memory limit.  Otherwise
// rows are padded to D + 4 floats.
#pragma once
#include "common.cuh"

namespace c3d {

constexpr int kGroupRows = 16;   // feature rows per CTA group
constexpr int kColsPerThread = 6;  // tile_logits: cols cg + 64*i, i < 6  => tile_rows <= 384
constexpr int kMaxTileRows = 64 * kColsPerThread;

struct BankLayout {
  int D, ld, swz;  // ld: row stride in floats; swz: XOR swizzle on/off
  __host__ __device__ static BankLayout make(int D) {
    BankLayout b; b.D = D; b.swz = (D % 32 == 0) ? 1 : 0; b.ld = b.swz ? D : D + 4; return b;
  }
  __device__ __forceinline__ int off(int r, int c) const {  // float offset of chunk c of row r
    return r * ld + ((swz ? (c ^ (r & 7)) : c) << 2);
  }
};

// bank_n rows [r0, r0 + rows) -> shared memory, all 16 B chunks in flight (cp.async).
// `issue` only starts the copies; cp_async_wait_all() (+ a barrier) completes them, so
// the caller can overlap the staging with its own global loads.
__device__ __forceinline__ void stage_bank_tile_issue(float* s_bank, const float* __restrict__ bank_n,
                                                      int r0, int rows, const BankLayout& L,
                                                      int nt = kRowsThreads /* staging threads */) {
  const int d4 = L.D >> 2;
  if (d4 <= nt && nt % d4 == 0) {  // fixed chunk per thread: no division in the loop
    const int c = threadIdx.x % d4, rstep = nt / d4;
    for (int r = threadIdx.x / d4; r < rows; r += rstep)
      cp_async16(s_bank + L.off(r, c), bank_n + (size_t)(r0 + r) * L.D + c * 4);
  } else {
    for (int i = threadIdx.x; i < rows * d4; i += nt) {
      const int r = i / d4, c = i - r * d4;
      cp_async16(s_bank + L.off(r, c), bank_n + (size_t)(r0 + r) * L.D + c * 4);
    }
  }
}
__device__ __forceinline__ void stage_bank_tile(float* s_bank, const float* __restrict__ bank_n, int r0,
                                                int rows, const BankLayout& L) {
  stage_bank_tile_issue(s_bank, bank_n, r0, rows, L);
  cp_async_wait_all();
}

// acc[r][i] = sum_d A[rg*4 + r][d] * bank[cg + 64 i][d].  Columns >= `rows` are
// computed on a clamped (valid) row and must be discarded by the caller: keeping the
// inner loop branch-free lets the compiler batch the ten 128-bit loads of an iteration
// (with per-column guards it serialised load -> use, 4x slower).
template <int kCPT = kColsPerThread>
__device__ __forceinline__ void tile_logits(const float* s_A, const float* s_bank, int rows,
                                            const BankLayout& L, float (&acc)[4][kCPT]) {
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
  int rowoff[kCPT], sw[kCPT];
#pragma unroll
  for (int i = 0; i < kCPT; ++i) {
    const int c = min(cg + 64 * i, rows - 1);
    rowoff[i] = c * L.ld;
    sw[i] = L.swz ? (c & 7) : 0;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < kCPT; ++i) acc[r][i] = 0.f;
  const int d4 = L.D >> 2;
  const float4* a_base = reinterpret_cast<const float4*>(s_A) + (size_t)(rg * 4) * d4;
#pragma unroll 2
  for (int j = 0; j < d4; ++j) {
    float4 a[4], b[kCPT];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = a_base[r * d4 + j];  // warp-wide broadcast
#pragma unroll
    for (int i = 0; i < kCPT; ++i)
      b[i] = *reinterpret_cast<const float4*>(s_bank + rowoff[i] + ((j ^ sw[i]) << 2));
#pragma unroll
    for (int i = 0; i < kCPT; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][i] += a[r].x * b[i].x; acc[r][i] += a[r].y * b[i].y;
        acc[r][i] += a[r].z * b[i].z; acc[r][i] += a[r].w * b[i].w;
      }
    }
  }
}

// Partial dA[16][D] += G[16][ks .. ke) . bank_tile for this thread's k range.
// Thread (kh, rgp, ct): k-split kh of kKS, rows rgp*kRP .. +kRP-1, chunks ct + CT*q with
// CT = 256 / (kKS * 16 / kRP) chunk-threads.  kKS > 1 keeps all eight warps busy when
// D/4 < 64 (D = 128: kKS = 2, kRP = 4, CT = 32); the caller adds the kKS partials.
// G is read 4 k at a time as warp-wide broadcasts (k0, ldg multiples of 4).
template <int kKS, int kRP, int kDch>
__device__ __forceinline__ void tile_gradT(const float* s_G, int ldg, int k0, const float* s_bank,
                                           int rows, const BankLayout& L, float4 (&acc)[kRP][kDch]) {
  constexpr int kRowGroups = kGroupRows / kRP;
  constexpr int CT = 256 / (kKS * kRowGroups);
  const int ct = threadIdx.x % CT;
  const int rgp = (threadIdx.x / CT) % kRowGroups;
  const int kh = threadIdx.x / (CT * kRowGroups);
  const int d4 = L.D >> 2;
  if (ct >= d4) return;
  const int per = (((rows + kKS - 1) / kKS) + 3) & ~3;   // k per split, multiple of 4
  const int ks = min(kh * per, rows), ke = min(ks + per, rows);
  const float* g_base = s_G + (size_t)(rgp * kRP) * ldg + k0;
  const int ke4 = ks + ((ke - ks) & ~3);
  for (int k = ks; k < ke4; k += 4) {
    float4 g[kRP];
#pragma unroll
    for (int r = 0; r < kRP; ++r) g[r] = *reinterpret_cast<const float4*>(g_base + r * ldg + k);
#pragma unroll
    for (int q = 0; q < kDch; ++q) {
      const int ch = ct + CT * q;
      if (ch < d4) {
        const float* bp = s_bank + (size_t)k * L.ld;
        const int m = L.swz ? (k & 7) : 0;   // k % 4 == 0: rows k..k+3 swizzle with m..m+3
        const float4 b0 = *reinterpret_cast<const float4*>(bp + ((L.swz ? (ch ^ m) : ch) << 2));
        const float4 b1 = *reinterpret_cast<const float4*>(bp + L.ld + ((L.swz ? (ch ^ (m + 1)) : ch) << 2));
        const float4 b2 = *reinterpret_cast<const float4*>(bp + 2 * L.ld + ((L.swz ? (ch ^ (m + 2)) : ch) << 2));
        const float4 b3 = *reinterpret_cast<const float4*>(bp + 3 * L.ld + ((L.swz ? (ch ^ (m + 3)) : ch) << 2));
#pragma unroll
        for (int r = 0; r < kRP; ++r) {
          acc[r][q].x += g[r].x * b0.x; acc[r][q].y += g[r].x * b0.y; acc[r][q].z += g[r].x * b0.z; acc[r][q].w += g[r].x * b0.w;
          acc[r][q].x += g[r].y * b1.x; acc[r][q].y += g[r].y * b1.y; acc[r][q].z += g[r].y * b1.z; acc[r][q].w += g[r].y * b1.w;
          acc[r][q].x += g[r].z * b2.x; acc[r][q].y += g[r].z * b2.y; acc[r][q].z += g[r].z * b2.z; acc[r][q].w += g[r].z * b2.w;
          acc[r][q].x += g[r].w * b3.x; acc[r][q].y += g[r].w * b3.y; acc[r][q].z += g[r].w * b3.z; acc[r][q].w += g[r].w * b3.w;
        }
      }
    }
  }
  for (int k = ke4; k < ke; ++k) {
#pragma unroll
    for (int q = 0; q < kDch; ++q) {
      const int ch = ct + CT * q;
      if (ch < d4) {
        const float4 b = *reinterpret_cast<const float4*>(s_bank + L.off(k, ch));
#pragma unroll
        for (int r = 0; r < kRP; ++r) {
          const float g = g_base[r * ldg + k];
          acc[r][q].x += g * b.x; acc[r][q].y += g * b.y; acc[r][q].z += g * b.z; acc[r][q].w += g * b.w;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- tensor cores ----
// The same two products on the tensor cores: mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, both TF32; a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, fp32 accumulate), which keeps
// the products at fp32 accuracy (the dropped lo.lo term is ~2^-22 relative), as the loss's 1e-5
// bar needs.  A 16-row group is exactly one M tile.  Operand fragments come straight from the
// shared-memory tiles: the XOR-swizzled bank layout is conflict-free for the B fragments of the
// logits product (8 rows x 4 consecutive floats) and 2-way for the gradient product; A rows are
// padded to D + 4 floats (kMmaPadA) so that the 8 rows of a fragment hit 8 different bank groups.
constexpr int kMmaPadA = 4;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// s_L[16][r0 .. r0 + rows) = (A[16][D] . bank_tile^T) * out_scale.  Warp w owns the 8-column tiles
// w, w + 8, ...; kNT = tiles per warp (8 * 8 * kNT >= tile rows).
template <int kNT>
__device__ __forceinline__ void mma_tile_logits(const float* s_A, int lda, const float* s_bank, int rows,
                                                const BankLayout& L, float* s_L, int ldl, int r0,
                                                float temperature) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[kNT][4];
  int brow[kNT];
#pragma unroll
  for (int j = 0; j < kNT; ++j) {
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    brow[j] = min(8 * (warp + 8 * j) + g, rows - 1);     // clamped: out-of-range columns are discarded
  }
  for (int k0 = 0; k0 < L.D; k0 += 8) {
    uint32_t ah[4], al[4];
    split_tf32(s_A[g * lda + k0 + t], ah[0], al[0]);
    split_tf32(s_A[(g + 8) * lda + k0 + t], ah[1], al[1]);
    split_tf32(s_A[g * lda + k0 + t + 4], ah[2], al[2]);
    split_tf32(s_A[(g + 8) * lda + k0 + t + 4], ah[3], al[3]);
    const int c0 = k0 >> 2;
#pragma unroll
    for (int j = 0; j < kNT; ++j) {
      if (8 * (warp + 8 * j) >= rows) continue;          // warp-uniform
      uint32_t bh[2], bl[2];
      split_tf32(s_bank[L.off(brow[j], c0) + t], bh[0], bl[0]);
      split_tf32(s_bank[L.off(brow[j], c0 + 1) + t], bh[1], bl[1]);
      mma_3x(acc[j], ah, al, bh, bl);
    }
  }
#pragma unroll
  for (int j = 0; j < kNT; ++j) {
    const int n0 = 8 * (warp + 8 * j) + 2 * t;
    if (n0 < rows) { s_L[g * ldl + r0 + n0] = acc[j][0] / temperature; s_L[(g + 8) * ldl + r0 + n0] = acc[j][2] / temperature; }
    if (n0 + 1 < rows) { s_L[g * ldl + r0 + n0 + 1] = acc[j][1] / temperature; s_L[(g + 8) * ldl + r0 + n0 + 1] = acc[j][3] / temperature; }
  }
}

// acc (this warp's 8-wide slices of dA[16][D]) += G[16][r0 .. r0 + rows) . bank_tile.  Warp w owns the
// feature slices d0 = 8 (w + 8 j), j < kNT.
template <int kNT>
__device__ __forceinline__ void mma_tile_gradT(const float* s_G, int ldg, int r0, const float* s_bank, int rows,
                                               const BankLayout& L, float (&acc)[kNT][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int k0 = 0; k0 < rows; k0 += 8) {
    const bool v0 = k0 + t < rows, v1 = k0 + t + 4 < rows;   // the last step of a tile may be partial
    uint32_t ah[4], al[4];
    split_tf32(v0 ? s_G[g * ldg + r0 + k0 + t] : 0.f, ah[0], al[0]);
    split_tf32(v0 ? s_G[(g + 8) * ldg + r0 + k0 + t] : 0.f, ah[1], al[1]);
    split_tf32(v1 ? s_G[g * ldg + r0 + k0 + t + 4] : 0.f, ah[2], al[2]);
    split_tf32(v1 ? s_G[(g + 8) * ldg + r0 + k0 + t + 4] : 0.f, ah[3], al[3]);
    const int ra = min(k0 + t, rows - 1), rb = min(k0 + t + 4, rows - 1);
#pragma unroll
    for (int j = 0; j < kNT; ++j) {
      const int d0 = 8 * (warp + 8 * j);
      if (d0 >= L.D) continue;                             // warp-uniform
      const int d = d0 + g;
      uint32_t bh[2], bl[2];
      split_tf32(v0 ? s_bank[L.off(ra, d >> 2) + (d & 3)] : 0.f, bh[0], bl[0]);
      split_tf32(v1 ? s_bank[L.off(rb, d >> 2) + (d & 3)] : 0.f, bh[1], bl[1]);
      mma_3x(acc[j], ah, al, bh, bl);
    }
  }
}

// Shared-memory plan: [bank tile][A / dA : 16 x D][L : 16 x ldl]; returns -1 if even a
// 64-row tile does not fit.
struct RowsPlan { int tile_rows, n_tiles, ldl; size_t smem; };
inline int plan_rows16(int D, int K, RowsPlan* out, int pad_a = 0) {
  const size_t budget = 227 * 1024 - 1024;  // 1 KB for the kernel's static shared memory
  const BankLayout L = BankLayout::make(D);
  const int ldl = (K + 3) & ~3;
  // the L region also holds the k-split partial sums of the gradient product
  // ((kKS-1) x 16 x D floats; kKS = 4 for D <= 64, 2 for D <= 128, see the launch table)
  const size_t part = (D <= 64) ? (size_t)3 * kGroupRows * D : (D <= 128 ? (size_t)kGroupRows * D : 0);
  size_t lfloats = (size_t)kGroupRows * ldl;
  if (lfloats < part) lfloats = part;
  const size_t fixed = ((size_t)kGroupRows * (D + pad_a) + lfloats) * 4;
  const size_t row = (size_t)L.ld * 4;
  if (fixed + 64 * row > budget) return -1;
  long long tr = (long long)((budget - fixed) / row);
  if (tr > kMaxTileRows) tr = kMaxTileRows;
  int n_tiles = (int)((K + tr - 1) / tr);
  int tile_rows = (K + n_tiles - 1) / n_tiles;       // balance the tiles
  tile_rows = (tile_rows + 7) & ~7;
  if (tile_rows > tr) tile_rows = (int)tr;
  out->tile_rows = tile_rows;
  out->n_tiles = (K + tile_rows - 1) / tile_rows;
  out->ldl = ldl;
  out->smem = fixed + (size_t)tile_rows * row;
  return 0;
}

}  // namespace c3d
