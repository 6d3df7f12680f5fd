// K6 (weight gradient), precision 2 ("f16x3"): dW[N,K] = dZ[M,N]^T @ X[M,K] on tcgen05 kind::f16 with fp16 hi / lo split
// operands (f16x3_common.cuh), reduction over the huge batch dimension M.
//
// Both operands are MN-major for the tensor core (the reduction index = batch row is the slow dimension of the row-major
// pairs): a k-block of 64 batch rows is staged as 64-feature chunks of [64 rows][128 B] (TMA box {64 features, 64 rows},
// SWIZZLE_128B) and described with MN-major SWIZZLE_128B descriptors (LBO = chunk stride, SBO = 1024 B between 8-row
// atoms); one MMA consumes 16 batch rows = two atoms.  Unlike the tf32 kernel (gemm_wgrad_tf32.cu) nothing is split in
// shared memory -- both halves of both operands arrive by TMA -- so the eight warps that used to rewrite the tiles now
// all drain accumulators: the CTA's slab is cut into segments of WGF_SEG k-blocks that alternate between two TMEM
// accumulators (the tensor core's fp32 adder truncates: a long chain drifts), and every finished segment is folded into the
// CTA's partial tile with round-to-nearest adds while the next one runs.  Split-K over CTAs, deterministic second-stage
// reduction (which also applies the exact power-of-two 1 / (s_dZ s_X)).
#include "f16x3_common.cuh"

namespace cusrl_b200 {

using namespace tc;

extern int g_f16_prefetch;  // gemm_f16x3.cu

constexpr int WGF_BM = 128;        // output features per tile (UMMA M)
constexpr int WGF_BKB = 64;        // batch rows per k-block
constexpr int WGF_CHUNK = 64;      // features per 128-byte swizzle span
constexpr int WGF_THREADS = 384;   // warps: 0 TMA, 1 MMA, 2 TMEM, 3 idle, 4-11 fold
constexpr int WGF_SEG = 16;        // k-blocks (= 1024 batch rows) accumulated inside the tensor core before a fold

struct WgradF16Params {
  float* partial;        // [splits][num_m_tiles*128][ldp]   raw accumulators (scaled by s_dZ s_X)
  int64_t ldp;
  int M, N, K;
  int num_m_tiles, num_n_tiles, splits;
  int rows_per_split;    // multiple of WGF_BKB
  int prefetch;          // k-blocks the L2 prefetch runs ahead of the loads (0 = off)
};

template <int BN>
struct WgradF16Cfg {
  static constexpr int CHUNK_BYTES = WGF_BKB * 128;                    // 8 KB
  static constexpr int A_BYTES = (WGF_BM / WGF_CHUNK) * CHUNK_BYTES;   // 16 KB
  static constexpr int B_BYTES = (BN / WGF_CHUNK) * CHUNK_BYTES;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (220 * 1024 / STAGE_BYTES) > 6 ? 6 : (220 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
  static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

template <int BN>
__global__ void __launch_bounds__(WGF_THREADS, 1)
wgrad_f16x3_kernel(const __grid_constant__ CUtensorMap tmDZhi, const __grid_constant__ CUtensorMap tmDZlo,
                   const __grid_constant__ CUtensorMap tmXhi, const __grid_constant__ CUtensorMap tmXlo, const WgradF16Params p) {
  using Cfg = WgradF16Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CHUNK_BYTES = Cfg::CHUNK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage layout: [A hi | B hi | A lo | B lo]
  auto sAhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sBhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES; };
  auto sBlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;   // [2] segment accumulator complete
  uint64_t* tempty = tfull + 2;          // [2] segment accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % (p.num_m_tiles * p.num_n_tiles);
  const int sp = blockIdx.x / (p.num_m_tiles * p.num_n_tiles);
  const int f0 = (tile / p.num_n_tiles) * WGF_BM;   // first output feature (row of dW)
  const int k0 = (tile % p.num_n_tiles) * BN;       // first input feature (column of dW)
  const int row_begin = sp * p.rows_per_split;
  int row_end = row_begin + p.rows_per_split;
  if (row_end > p.M) row_end = p.M;
  const int num_kb = row_end > row_begin ? (row_end - row_begin + WGF_BKB - 1) / WGF_BKB : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDZhi);
    tma_prefetch_desc(&tmDZlo);
    tma_prefetch_desc(&tmXhi);
    tma_prefetch_desc(&tmXlo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // both operands stream from HBM and two 96 KB stages keep too few requests in flight to cover its latency: a second
      // cursor pulls the tiles of k-block kb + p.prefetch into L2 (see gemm_f16x3.cu)
      auto prefetch_kb = [&](int kb) {
        if (kb >= num_kb) return;
        const int r0 = row_begin + kb * WGF_BKB;
#pragma unroll
        for (int c = 0; c < WGF_BM / WGF_CHUNK; ++c) {
          tma_prefetch_2d(&tmDZhi, f0 + c * WGF_CHUNK, r0);
          tma_prefetch_2d(&tmDZlo, f0 + c * WGF_CHUNK, r0);
        }
#pragma unroll
        for (int c = 0; c < BN / WGF_CHUNK; ++c) {
          tma_prefetch_2d(&tmXhi, k0 + c * WGF_CHUNK, r0);
          tma_prefetch_2d(&tmXlo, k0 + c * WGF_CHUNK, r0);
        }
      };
      for (int i = 0; i < p.prefetch; ++i) prefetch_kb(i);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int r0 = row_begin + kb * WGF_BKB;
        if (p.prefetch > 0) prefetch_kb(kb + p.prefetch);
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        // rows beyond M are zero-filled by TMA and contribute nothing; a slab is a multiple of 64 rows, so a k-block never
        // straddles two slabs
#pragma unroll
        for (int c = 0; c < WGF_BM / WGF_CHUNK; ++c) {
          tma_load_2d(sAhi(s) + c * CHUNK_BYTES, &tmDZhi, f0 + c * WGF_CHUNK, r0, &full[s]);
          tma_load_2d(sAlo(s) + c * CHUNK_BYTES, &tmDZlo, f0 + c * WGF_CHUNK, r0, &full[s]);
        }
#pragma unroll
        for (int c = 0; c < BN / WGF_CHUNK; ++c) {
          tma_load_2d(sBhi(s) + c * CHUNK_BYTES, &tmXhi, k0 + c * WGF_CHUNK, r0, &full[s]);
          tma_load_2d(sBlo(s) + c * CHUNK_BYTES, &tmXlo, k0 + c * WGF_CHUNK, r0, &full[s]);
        }
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      constexpr uint32_t idesc = make_idesc_f16(WGF_BM, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int seg = kb / WGF_SEG, a = seg & 1, kin = kb - seg * WGF_SEG;
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        if (kin == 0) {
          mbar_wait(&tempty[a], ((seg >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t ahi = smem_u32(sAhi(s)), bhi = smem_u32(sBhi(s)), alo = smem_u32(sAlo(s)), blo = smem_u32(sBlo(s));
        // one MMA consumes 16 batch rows = two 8-row swizzle atoms (2 x 1024 B) of every 64-feature chunk
#pragma unroll
        for (int k = 0; k < WGF_BKB / 16; ++k) {
          const uint32_t off = (uint32_t)k * 2048;
          const uint64_t dah = make_smem_desc_sw128(ahi + off, CHUNK_BYTES, 1024, 2);
          const uint64_t dbh = make_smem_desc_sw128(bhi + off, CHUNK_BYTES, 1024, 2);
          mma_f16_ss(d_tmem, dah, dbh, idesc, (kin > 0 || k > 0) ? 1u : 0u);
          mma_f16_ss(d_tmem, make_smem_desc_sw128(alo + off, CHUNK_BYTES, 1024, 2), dbh, idesc, 1u);
          mma_f16_ss(d_tmem, dah, make_smem_desc_sw128(blo + off, CHUNK_BYTES, 1024, 2), idesc, 1u);
        }
        mma_commit(&empty[s]);
        if (kin == WGF_SEG - 1 || kb == num_kb - 1) mma_commit(&tfull[a]);
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // fold: every finished segment is added into this CTA's partial tile (the first one overwrites).  Two warps per TMEM
    // lane quarter, each takes half of the tile's columns.
    const int ew = warp & 3, half = (warp - 4) >> 2;
    const int frow = f0 + ew * 32 + lane;  // output feature handled by this thread
    float* prow = p.partial + ((int64_t)sp * p.num_m_tiles * WGF_BM + frow) * p.ldp + k0;
    const int num_seg = (num_kb + WGF_SEG - 1) / WGF_SEG;
    const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
    if (num_seg == 0) {
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 4)
        if (k0 + c0 < p.ldp) *reinterpret_cast<uint4*>(prow + c0) = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll 1
    for (int seg = 0; seg < num_seg; ++seg) {
      const int a = seg & 1;
      mbar_wait(&tfull[a], (seg >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, r);
        float4 prev[8];
        if (seg > 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            prev[j] = (k0 + c0 + 4 * j < p.ldp) ? *reinterpret_cast<const float4*>(prow + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (seg > 0) v.x += prev[j].x, v.y += prev[j].y, v.z += prev[j].z, v.w += prev[j].w;
          if (k0 + c0 + 4 * j < p.ldp) *reinterpret_cast<float4*>(prow + c0 + 4 * j) = v;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// dW[n,k] (+)= (sum_s partial[s][n][k]) / (s_dZ s_X)  in a fixed order; accumulate != 0 adds to the existing gradient.
__global__ void wgrad_f16_reduce_kernel(const float* __restrict__ partial, int splits, int64_t split_stride, int64_t ldp,
                                        const float* __restrict__ bound_dz, const float* __restrict__ bound_x,
                                        float* __restrict__ dW, int64_t lddw, int N, int K, int accumulate) {
  const float inv = 1.f / (f16x3_scale(__ldg(bound_dz)) * f16x3_scale(__ldg(bound_x)));
  if ((K & 3) == 0 && (ldp & 3) == 0 && (lddw & 3) == 0 && (split_stride & 3) == 0 && aligned_to(partial, 16) && aligned_to(dW, 16)) {
    // A group of four lanes owns four adjacent k (one float4): lane q of the group sums splits q, q + 4, q + 8, ... and the four
    // sums are combined by two shuffles -- a fixed association, so the result is deterministic.  One thread per element
    // walking all ~37 splits left the kernel latency-bound (14 us for 19 MB of L2-resident partials at the 512 x 256 layer).
    const int K4 = K >> 2, total4 = N * K4;
    const int q = threadIdx.x & 3;
    const int stride = (gridDim.x * blockDim.x) >> 2;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; i < ((total4 + stride - 1) / stride) * stride; i += stride) {
      const bool in = i < total4;
      const int n = in ? i / K4 : 0, k = in ? (i - n * K4) << 2 : 0;
      const float* src = partial + (int64_t)n * ldp + k;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
#pragma unroll 4
        for (int sp = q; sp < splits; sp += 4) {
          const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)sp * split_stride);
          acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      if (in && q == 0) {
        float4* dst = reinterpret_cast<float4*>(dW + (int64_t)n * lddw + k);
        float4 o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (accumulate) {
          const float4 d = *dst;
          o.x += d.x, o.y += d.y, o.z += d.z, o.w += d.w;
        }
        *dst = o;
      }
    }
    return;
  }
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i - n * K;
    const float* src = partial + (int64_t)n * ldp + k;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[(int64_t)s * split_stride];
    acc *= inv;
    float* dst = dW + (int64_t)n * lddw + k;
    *dst = accumulate ? *dst + acc : acc;
  }
}

template <int BN>
static int launch_wgrad_f16(const CUtensorMap& tZh, const CUtensorMap& tZl, const CUtensorMap& tXh, const CUtensorMap& tXl,
                            const WgradF16Params& p, cudaStream_t s) {
  using Cfg = WgradF16Cfg<BN>;
  auto kern = wgrad_f16x3_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("wgrad_f16x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.num_m_tiles * p.num_n_tiles * p.splits;
  kern<<<grid, WGF_THREADS, Cfg::SMEM_BYTES, s>>>(tZh, tZl, tXh, tXl, p);
  return check_launch("wgrad_f16x3_kernel");
}

static void wgrad_f16_plan(int64_t M, int64_t N, int64_t K, int* bn, int* mt, int* nt, int* splits, int* rows_per_split,
                           int64_t* ldp) {
  *bn = K > 128 ? 256 : 128;
  *mt = (int)((N + WGF_BM - 1) / WGF_BM);
  *nt = (int)((K + *bn - 1) / *bn);
  const int tiles = *mt * *nt;
  int sp = sm_count() / tiles;
  if (sp < 1) sp = 1;
  int64_t rps = ((M + sp - 1) / sp + WGF_BKB - 1) / WGF_BKB * WGF_BKB;
  sp = (int)((M + rps - 1) / rps);
  *splits = sp;
  *rows_per_split = (int)rps;
  *ldp = (int64_t)*nt * *bn;
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

size_t cusrl_b200_wgrad_f16x3_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int bn, mt, nt, splits, rps;
  int64_t ldp;
  wgrad_f16_plan(M, N, K, &bn, &mt, &nt, &splits, &rps, &ldp);
  return (size_t)splits * mt * WGF_BM * ldp * sizeof(float) + 256;
}

int cusrl_b200_linear_wgrad_f16x3(const uint16_t* dZhi, const uint16_t* dZlo, int64_t lddz, const float* dz_bound,
                                  const uint16_t* Xhi, const uint16_t* Xlo, int64_t ldx, const float* x_bound, float* dW,
                                  int64_t lddw, int64_t M, int64_t N, int64_t K, int accumulate, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  CUSRL_REQUIRE(dZhi && dZlo && Xhi && Xlo && dz_bound && x_bound && dW && workspace, CUSRL_B200_EINVAL, "wgrad_f16x3: null pointer");
  CUSRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N <= 65536 && K <= 65536, CUSRL_B200_EINVAL,
                "wgrad_f16x3: bad problem size");
  CUSRL_REQUIRE((lddz % 8) == 0 && (ldx % 8) == 0 && lddz >= N && ldx >= K && lddw >= K, CUSRL_B200_EALIGN,
                "wgrad_f16x3: pair leading dimensions must be multiples of 8 halves and cover the row");
  CUSRL_REQUIRE(aligned_to(dZhi, 16) && aligned_to(dZlo, 16) && aligned_to(Xhi, 16) && aligned_to(Xlo, 16) && aligned_to(workspace, 16),
                CUSRL_B200_EALIGN, "wgrad_f16x3: operands and workspace must be 16-byte aligned");
  CUSRL_REQUIRE(workspace_bytes >= cusrl_b200_wgrad_f16x3_workspace_bytes(M, N, K), CUSRL_B200_ESCRATCH,
                "wgrad_f16x3: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int bn, mt, nt, splits, rps;
  int64_t ldp;
  wgrad_f16_plan(M, N, K, &bn, &mt, &nt, &splits, &rps, &ldp);
  const uint64_t Np = (uint64_t)((N + 7) / 8 * 8), Kp = (uint64_t)((K + 7) / 8 * 8);
  CUtensorMap tZh, tZl, tXh, tXl;
  if (int e = encode_tmap_2d_f16(&tZh, dZhi, Np, (uint64_t)M, (uint64_t)lddz, WGF_CHUNK, WGF_BKB, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tZl, dZlo, Np, (uint64_t)M, (uint64_t)lddz, WGF_CHUNK, WGF_BKB, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tXh, Xhi, Kp, (uint64_t)M, (uint64_t)ldx, WGF_CHUNK, WGF_BKB, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tXl, Xlo, Kp, (uint64_t)M, (uint64_t)ldx, WGF_CHUNK, WGF_BKB, TMAP_SW128)) return e;
  WgradF16Params p{};
  p.partial = (float*)workspace, p.ldp = ldp, p.M = (int)M, p.N = (int)N, p.K = (int)K;
  p.num_m_tiles = mt, p.num_n_tiles = nt, p.splits = splits, p.rows_per_split = rps;
  p.prefetch = g_f16_prefetch;
  int e = bn == 256 ? launch_wgrad_f16<256>(tZh, tZl, tXh, tXl, p, s) : launch_wgrad_f16<128>(tZh, tZl, tXh, tXl, p, s);
  if (e) return e;
  const int64_t split_stride = (int64_t)mt * WGF_BM * ldp;
  int blocks = (int)((N * K + 255) / 256);   // four lanes per float4 of the output
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  wgrad_f16_reduce_kernel<<<blocks, 256, 0, s>>>(p.partial, splits, split_stride, ldp, dz_bound, x_bound, dW, lddw, (int)N,
                                                 (int)K, accumulate);
  return check_launch("wgrad_f16_reduce_kernel");
}

}  // extern "C"
