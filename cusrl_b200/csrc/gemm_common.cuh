// Shared definitions of the K-major tcgen05 dense-layer kernels (gemm_tf32.cu: 1-SM MMA + TMA multicast;
// the cta_group::2 variant was measured slower in round 1 and removed in round 2).
#pragma once
#include "tc_common.cuh"

namespace cusrl_b200 {

using namespace tc;

constexpr int BM = 128;           // rows per CTA tile (UMMA M)
constexpr int BK = 32;            // fp32 per k-block = 128 bytes = one SWIZZLE_128B span
constexpr int UMMA_K = 8;         // tf32 elements per tcgen05.mma
constexpr int kGemmThreads = 512;
constexpr int kEpiWarpBytes = 32 * 32 * 4;          // one 32-row x 32-column fp32 chunk per epilogue warp
constexpr int kEpiStageBytes = 8 * kEpiWarpBytes;    // 8 epilogue warps
constexpr int kSmemBudget = 192 * 1024;              // pipeline stages (the epilogue staging and barriers come on top)

enum { EPI_BIAS_ACT = 0, EPI_ACT_GRAD = 1 };

struct GemmParams {
  float* out;
  int64_t ldo;
  const float* bias;   // EPI_BIAS_ACT (may be null)
  const float* aux;    // EPI_ACT_GRAD: post-activation output of the layer below (may be null: plain copy)
  int64_t ldaux;
  int M, N, K, act;
  int num_m_tiles, num_n_tiles;
  int num_items;  // work items of a cluster: (pair of M tiles) x (N tile)
  float* colsum;  // EPI_ACT_GRAD, optional: [gridDim.x][4][N] column sums of the output (bias gradient partials)
};

// expm1(z) for z <= 0 in ~11 instructions, both branches evaluated and selected (no divergence): libdevice's expm1f costs
// ~25 instructions per element and made the 8 epilogue warps the bottleneck of the whole GEMM (measured).
// |z| < 1/8: degree-6 Taylor polynomial (truncation < 4e-8 relative); otherwise ex2.approx(z log2 e) - 1, whose
// 2-ulp error in the exponential is < 2.4e-7 absolute on a result of magnitude > 0.117.
__device__ __forceinline__ float expm1_neg(float z) {
  float poly = fmaf(z, 1.f / 720.f, 1.f / 120.f);
  poly = fmaf(poly, z, 1.f / 24.f);
  poly = fmaf(poly, z, 1.f / 6.f);
  poly = fmaf(poly, z, 0.5f);
  poly = fmaf(poly, z, 1.f);
  poly *= z;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 1.4426950408889634f));
  const float big = e - 1.f;
  return z > -0.125f ? poly : big;
}
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return z > 0.f ? z : expm1_neg(z);  // ELU(alpha=1), torch.nn.functional.elu
  if (act == 2) return fmaxf(z, 0.f);
  return z;
}
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  if (act == 1) return y > 0.f ? 1.f : y + 1.f;  // d/dz ELU(z) = exp(z) = y + 1 for z <= 0
  if (act == 2) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}


// One 32-row x 32-column chunk of the epilogue, executed by one warp (lane = accumulator row within the chunk):
// TMEM -> registers -> bias + activation (forward) or x act'(aux) (data gradient) -> shared-memory staging in the
// SWIZZLE_128B pattern -> ONE bulk tensor store, which writes full 128-byte lines and clips rows >= M / columns >= N.
// The bias / aux operands are fetched while the TMEM load is in flight, and the wait for the previous chunk's bulk
// store (it must have finished READING the staging buffer) comes after the arithmetic, so the two overlap.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const CUtensorMap* tmOut, uint8_t* stg, uint32_t taddr,
                                               int row0, int col0, int lane) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  float4 e[8];
  if (EPI == EPI_BIAS_ACT) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      e[q] = (p.bias && col0 + 4 * q < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * q))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (p.aux) {
    // coalesced: load i of lane l fetches 16-byte unit (l % 8) of row 4 i + l / 8, i.e. every instruction reads four
    // full 128-byte lines (row-per-lane loads touch 32 lines per instruction and saturated the L1 pipe, ncu)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3), c = col0 + 4 * (lane & 7);
      e[i] = (row0 + r < p.M && c < p.N) ? __ldg(reinterpret_cast<const float4*>(p.aux + (int64_t)(row0 + r) * p.ldaux + c))
                                        : make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
  tmem_ld_wait();
  float4 v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                       __uint_as_float(r[4 * q + 3]));
    if (EPI == EPI_BIAS_ACT) {
      v[q].x = act_fwd(v[q].x + e[q].x, p.act), v[q].y = act_fwd(v[q].y + e[q].y, p.act);
      v[q].z = act_fwd(v[q].z + e[q].z, p.act), v[q].w = act_fwd(v[q].w + e[q].w, p.act);
    }
  }
  if (lane == 0) tma_store_wait_read();  // the previous chunk's store has finished reading the staging buffer
  __syncwarp();
  float4* srow = reinterpret_cast<float4*>(stg + lane * 128);
  const int sw = lane & 7;  // SWIZZLE_128B: 16-byte unit q of row r lives at unit q ^ (r % 8)
  if (EPI == EPI_ACT_GRAD && p.aux) {
    // transpose the coalesced aux fragments through the staging buffer: afterwards e[q] is unit q of this lane's row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3);
      reinterpret_cast<float4*>(stg + r * 128)[(lane & 7) ^ (r & 7)] = e[i];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = srow[q ^ sw];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q].x *= act_grad_from_output(e[q].x, p.act), v[q].y *= act_grad_from_output(e[q].y, p.act);
      v[q].z *= act_grad_from_output(e[q].z, p.act), v[q].w *= act_grad_from_output(e[q].w, p.act);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) srow[q ^ sw] = v[q];
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmOut, stg, col0, row0);
    tma_store_commit();
  }
  if (EPI == EPI_ACT_GRAD && p.colsum) {
    // bias gradient of the layer below = column sums of this output: lane c adds up column c of the staged chunk (rows
    // >= M hold exact zeros: their accumulators are products of zero-filled A rows) in a fixed order and accumulates
    // into the slot owned by this (CTA, row quarter, column) -- one owner per slot, so plain read-modify-write.
    const float* col = reinterpret_cast<const float*>(stg) + (lane & 3);
    const int unit = lane >> 2;
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) sum += col[r * 32 + ((unit ^ (r & 7)) << 2)];
    if (col0 + lane < p.N) {
      float* slot = p.colsum + ((int64_t)blockIdx.x * 4 + ((threadIdx.x >> 5) & 3)) * p.N + col0 + lane;
      *slot += sum;
    }
  }
}

// grid size the launchers use for `num_items` work items (2 CTAs per cluster)
int gemm_grid_ctas(int num_items);

}  // namespace cusrl_b200
