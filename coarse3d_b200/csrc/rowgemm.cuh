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
                                                      int r0, int rows, const BankLayout& L) {
  const int d4 = L.D >> 2;
  for (int i = threadIdx.x; i < rows * d4; i += blockDim.x) {
    const int r = i / d4, c = i - r * d4;
    cp_async16(s_bank + L.off(r, c), bank_n + (size_t)(r0 + r) * L.D + c * 4);
  }
}
__device__ __forceinline__ void stage_bank_tile(float* s_bank, const float* __restrict__ bank_n, int r0,
                                                int rows, const BankLayout& L) {
  stage_bank_tile_issue(s_bank, bank_n, r0, rows, L);
  cp_async_wait_all();
}

// acc[r][i] = sum_d A[rg*4 + r][d] * bank[cg + 64 i][d]   (rows >= `rows` give 0)
__device__ __forceinline__ void tile_logits(const float* s_A, const float* s_bank, int rows,
                                            const BankLayout& L, float (&acc)[4][kColsPerThread]) {
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < kColsPerThread; ++i) acc[r][i] = 0.f;
  const int d4 = L.D >> 2;
  const float4* a_base = reinterpret_cast<const float4*>(s_A) + (size_t)(rg * 4) * d4;
  for (int j = 0; j < d4; ++j) {
    float4 a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = a_base[r * d4 + j];  // warp-wide broadcast
#pragma unroll
    for (int i = 0; i < kColsPerThread; ++i) {
      const int c = cg + 64 * i;
      if (c < rows) {
        const float4 b = *reinterpret_cast<const float4*>(s_bank + L.off(c, j));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][i] += a[r].x * b.x; acc[r][i] += a[r].y * b.y;
          acc[r][i] += a[r].z * b.z; acc[r][i] += a[r].w * b.w;
        }
      }
    }
  }
}

// dA[16][D] += G[16][k0 .. k0+rows) . bank_tile.  Thread (rgp, ct): kRP rows
// rgp*kRP.., chunks ct + CT*q (CT = 16*kRP chunk-threads per row group, so all 256
// threads work for D = 64*kRP/... see launch table).  G is read 4 k at a time as
// warp-wide broadcasts; requires k0 % 4 == 0 and ldg % 4 == 0.
template <int kRP, int kDch>
__device__ __forceinline__ void tile_gradT(const float* s_G, int ldg, int k0, const float* s_bank,
                                           int rows, const BankLayout& L, float4 (&acc)[kRP][kDch]) {
  constexpr int CT = 16 * kRP;
  const int ct = threadIdx.x % CT, rgp = threadIdx.x / CT;
  const int d4 = L.D >> 2;
  if (ct >= d4) return;
  const float* g_base = s_G + (size_t)(rgp * kRP) * ldg + k0;
  const int rows4 = rows & ~3;
  for (int k = 0; k < rows4; k += 4) {
    float4 g[kRP];
#pragma unroll
    for (int r = 0; r < kRP; ++r) g[r] = *reinterpret_cast<const float4*>(g_base + r * ldg + k);
#pragma unroll
    for (int q = 0; q < kDch; ++q) {
      const int ch = ct + CT * q;
      if (ch < d4) {
        const float4 b0 = *reinterpret_cast<const float4*>(s_bank + L.off(k, ch));
        const float4 b1 = *reinterpret_cast<const float4*>(s_bank + L.off(k + 1, ch));
        const float4 b2 = *reinterpret_cast<const float4*>(s_bank + L.off(k + 2, ch));
        const float4 b3 = *reinterpret_cast<const float4*>(s_bank + L.off(k + 3, ch));
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
  for (int k = rows4; k < rows; ++k) {
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

// Shared-memory plan: [bank tile][A / dA : 16 x D][L : 16 x ldl]; returns -1 if even a
// 64-row tile does not fit.
struct RowsPlan { int tile_rows, n_tiles, ldl; size_t smem; };
inline int plan_rows16(int D, int K, RowsPlan* out) {
  const size_t budget = 227 * 1024 - 1024;  // 1 KB for the kernel's static shared memory
  const BankLayout L = BankLayout::make(D);
  const int ldl = (K + 3) & ~3;
  const size_t fixed = ((size_t)kGroupRows * D + (size_t)kGroupRows * ldl) * 4;
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
