// K6 (weight gradient): dW[N,K] = dZ[M,N]^T @ X[M,K] on tcgen05, reduction over the huge batch dimension M.
//
// Both operands are "MN-major" for the tensor core: the reduction index (batch row) is the slow dimension of the
// row-major activations, so each k-block of 32 batch rows is staged as 32-feature chunks of [32 rows][128 B]
// (TMA box {32 features, 32 rows}, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) and described with UMMA MN-major descriptors of
// layout type SWIZZLE_128B_BASE32B -- the only shared-memory layout the tensor core accepts for MN-major tf32
// operands: 32-byte chunks XOR-ed with (row % 4), atoms of 4 rows x 128 B (LBO = chunk stride, SBO = 4-row group stride).
//
// Split-K over CTAs: every CTA owns one 128 x BN output tile and a contiguous slab of batch rows and produces one
// partial tile in a workspace; a second kernel adds the partials in a fixed order (deterministic) into the gradient
// arena.  The tensor core's accumulator adder TRUNCATES, so a long accumulation chain drifts (7e-5 relative after
// 10 k batch rows, measured): the slab is therefore cut into segments of WG_SEG k-blocks that alternate between two
// TMEM accumulators, and the four epilogue warps fold every finished segment into the CTA's partial tile with
// round-to-nearest fp32 adds (read-modify-write of 128 KB that stays in L2) while the next segment is running.  3xTF32 splits BOTH operands in shared
// memory (they are activations / gradients, there is nothing to pre-split).
#include "tc_common.cuh"

namespace cusrl_b200 {

using namespace tc;

constexpr int WG_BM = 128;        // output features per tile (UMMA M)
constexpr int WG_BKB = 32;        // batch rows per k-block
constexpr int WG_CHUNK = 32;      // features per 128-byte swizzle span
constexpr int WG_THREADS = 512;
constexpr int WG_SEG = 32;        // k-blocks (= 1024 batch rows) accumulated inside the tensor core before a drain
                                  // (64 measured 9.8x cuBLAS-fp32's error at M = 393 216, N x K = 128 x 256: over the 8x bar
                                  //  of tests/test_gemm_gpu.py; 32 halves the truncation drift for ~6 % more time)

struct WgradParams {
  float* partial;        // [splits][num_m_tiles*128][ldp]
  int64_t ldp;           // padded K (multiple of 4)
  int M, N, K;           // batch rows, output features, input features
  int num_m_tiles, num_n_tiles, splits;
  int rows_per_split;    // multiple of WG_BKB
};

template <int BN, int PASSES>
struct WgradCfg {
  static constexpr int A_BYTES = WG_BM * WG_BKB * 4;   // 16 KB: 4 chunks x [32 rows][128 B]
  static constexpr int B_BYTES = BN * WG_BKB * 4;      // BN/32 chunks
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (PASSES == 3 ? 2 : 1);
  static constexpr int STAGES = (220 * 1024 / STAGE_BYTES) > 6 ? 6 : (220 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
  static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using Cfg = WgradCfg<BN, PASSES>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CHUNK_BYTES = WG_BKB * 128;  // 4096
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage layout: [A | B | A lo | B lo]  (A, B are the raw TMA destinations = the hi operands: the MMA truncates)
  auto sA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sB = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sLo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* split = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* tfull = bars + 3 * STAGES;   // [2] segment accumulator complete
  uint64_t* tempty = tfull + 2;          // [2] segment accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % (p.num_m_tiles * p.num_n_tiles);
  const int sp = blockIdx.x / (p.num_m_tiles * p.num_n_tiles);
  const int f0 = (tile / p.num_n_tiles) * WG_BM;   // first output feature (row of dW)
  const int k0 = (tile % p.num_n_tiles) * BN;      // first input feature (column of dW)
  const int row_begin = sp * p.rows_per_split;
  int row_end = row_begin + p.rows_per_split;
  if (row_end > p.M) row_end = p.M;
  const int num_kb = row_end > row_begin ? (row_end - row_begin + WG_BKB - 1) / WG_BKB : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDZ);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 8);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int r0 = row_begin + kb * WG_BKB;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
        // rows beyond M (or beyond this split's slab end, when the slab is not a multiple of 32 -- it always is,
        // except for the global tail) are zero-filled by TMA and contribute nothing
#pragma unroll
        for (int c = 0; c < WG_BM / WG_CHUNK; ++c) tma_load_2d(sA(s) + c * CHUNK_BYTES, &tmDZ, f0 + c * WG_CHUNK, r0, &full[s]);
#pragma unroll
        for (int c = 0; c < BN / WG_CHUNK; ++c) tma_load_2d(sB(s) + c * CHUNK_BYTES, &tmX, k0 + c * WG_CHUNK, r0, &full[s]);
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      constexpr uint32_t idesc = make_idesc_tf32(WG_BM, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int seg = kb / WG_SEG, a = seg & 1, kin = kb - seg * WG_SEG;
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        if (kin == 0) {
          mbar_wait(&tempty[a], ((seg >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA(s)), b_addr = smem_u32(sB(s));
        const uint32_t alo_addr = smem_u32(sLo(s)), blo_addr = alo_addr + Cfg::A_BYTES;
        // hi x hi first (raw fp32 tiles: the tensor core truncates to TF32), the two lo passes once the splitter is done.
        // One MMA consumes 8 batch rows = two 4-row swizzle atoms (2 x 512 B) of every 32-feature chunk.
#pragma unroll
        for (int k = 0; k < WG_BKB / 8; ++k) {
          const uint32_t off = (uint32_t)k * 1024;
          mma_tf32_ss(d_tmem, make_smem_desc_sw128(a_addr + off, CHUNK_BYTES, 512, 1),
                      make_smem_desc_sw128(b_addr + off, CHUNK_BYTES, 512, 1), idesc, (kin > 0 || k > 0) ? 1u : 0u);
        }
        if (PASSES == 3) {
          mbar_wait(&split[s], ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < WG_BKB / 8; ++k) {
            const uint32_t off = (uint32_t)k * 1024;
            const uint64_t da = make_smem_desc_sw128(a_addr + off, CHUNK_BYTES, 512, 1);
            const uint64_t db = make_smem_desc_sw128(b_addr + off, CHUNK_BYTES, 512, 1);
            mma_tf32_ss(d_tmem, make_smem_desc_sw128(alo_addr + off, CHUNK_BYTES, 512, 1), db, idesc, 1u);
            mma_tf32_ss(d_tmem, da, make_smem_desc_sw128(blo_addr + off, CHUNK_BYTES, 512, 1), idesc, 1u);
          }
        }
        mma_commit(&empty[s]);
        if (kin == WG_SEG - 1 || kb == num_kb - 1) mma_commit(&tfull[a]);
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // epilogue: fold every finished segment into this CTA's partial tile (the first one overwrites)
    const int ew = warp - 4;
    const int frow = f0 + ew * 32 + lane;  // output feature handled by this thread
    float* prow = p.partial + ((int64_t)sp * p.num_m_tiles * WG_BM + frow) * p.ldp + k0;
    const int num_seg = (num_kb + WG_SEG - 1) / WG_SEG;
    if (num_seg == 0) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 4)
        if (k0 + c0 < p.ldp) *reinterpret_cast<uint4*>(prow + c0) = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll 1
    for (int seg = 0; seg < num_seg; ++seg) {
      const int a = seg & 1;
      mbar_wait(&tfull[a], (seg >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
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
  } else if (PASSES == 3 && warp >= 8) {
    // splitter: lo = x - trunc_tf32(x) of both operand tiles, written to the mirror buffer at identical offsets
    const int t = threadIdx.x - 256;  // 0..255
    constexpr int VECS = (Cfg::A_BYTES + Cfg::B_BYTES) / 16;
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[s], ph);
      const uint4* hi = reinterpret_cast<const uint4*>(sA(s));
      float4* lo = reinterpret_cast<float4*>(sLo(s));
#pragma unroll 4
      for (int idx = t; idx < VECS; idx += 256) {
        const uint4 x = hi[idx];
        lo[idx] = make_float4(__uint_as_float(x.x) - __uint_as_float(x.x & 0xffffe000u),
                              __uint_as_float(x.y) - __uint_as_float(x.y & 0xffffe000u),
                              __uint_as_float(x.z) - __uint_as_float(x.z & 0xffffe000u),
                              __uint_as_float(x.w) - __uint_as_float(x.w & 0xffffe000u));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
      if (++s == STAGES) s = 0, ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// dW[n,k] (+)= sum_s partial[s][n][k]  in a fixed order; accumulate != 0 adds to the existing gradient.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int64_t split_stride, int64_t ldp,
                                    float* __restrict__ dW, int64_t lddw, int N, int K, int accumulate) {
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i - n * K;
    const float* src = partial + (int64_t)n * ldp + k;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[(int64_t)s * split_stride];
    float* dst = dW + (int64_t)n * lddw + k;
    *dst = accumulate ? *dst + acc : acc;
  }
}

// db[n] (+)= sum_m dZ[m, n]   (bias gradient); block partials in double, fixed-order finalisation
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dz, int64_t ld, int M, int N,
                                                             int rows_per_block, float* __restrict__ partial) {
  // thread = (row lane, column): 256 threads = 8 row lanes x 32 columns per pass
  const int col_lane = threadIdx.x & 31, row_lane = threadIdx.x >> 5;
  __shared__ float red[8][33];
  const int r0 = blockIdx.x * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  if ((N & 3) == 0 && (ld & 3) == 0 && aligned_to(dz, 16) && aligned_to(partial, 16)) {
    // wide rows (the LSTM's [T * Nb, 4H] gate gradients: 100 MB): four columns per thread, a warp reads 512 contiguous bytes
    // of a row, the block's rows are walked with four loads in flight per thread; no shared memory, no barriers.  The
    // 32-column passes below reached 1.9 TB/s on that shape.
    for (int c = threadIdx.x * 4; c < N; c += 4 * blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = dz + (int64_t)r0 * ld + c;
      int r = r0;
      for (; r + 4 <= r1; r += 4, src += 4 * ld) {
        const float4 a = ldg_stream4(src), b = ldg_stream4(src + ld), c2 = ldg_stream4(src + 2 * ld), d = ldg_stream4(src + 3 * ld);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
        acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
        acc.x += c2.x, acc.y += c2.y, acc.z += c2.z, acc.w += c2.w;
        acc.x += d.x, acc.y += d.y, acc.z += d.z, acc.w += d.w;
      }
      for (; r < r1; ++r, src += ld) {
        const float4 a = ldg_stream4(src);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      }
      *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * N + c) = acc;
    }
    return;
  }
  for (int c0 = 0; c0 < N; c0 += 32) {
    const int c = c0 + col_lane;
    float acc = 0.f;
    if (c < N)
      for (int r = r0 + row_lane; r < r1; r += 8) acc += dz[(int64_t)r * ld + c];
    red[row_lane][col_lane] = acc;
    __syncthreads();
    if (row_lane == 0 && c < N) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][col_lane];
      partial[(int64_t)blockIdx.x * N + c] = s;
    }
    __syncthreads();
  }
}
// one warp per output column: lanes stride over the block partials (fixed order -> deterministic)
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblocks, int N, float* __restrict__ db,
                                    int accumulate) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= N) return;
  double acc = 0.0;
  for (int b = lane; b < nblocks; b += 32) acc += partial[(int64_t)b * N + c];
  acc = warp_sum(acc);
  if (lane == 0) db[c] = accumulate ? db[c] + (float)acc : (float)acc;
}

int colsum_finalize(const float* partial, int nblocks, int N, float* db, int accumulate, cudaStream_t s) {
  colsum_final_kernel<<<(unsigned)(((int64_t)N * 32 + 255) / 256), 256, 0, s>>>(partial, nblocks, N, db, accumulate);
  return check_launch("colsum_final_kernel");
}

template <int BN, int PASSES>
static int launch_wgrad(const CUtensorMap& tDZ, const CUtensorMap& tX, const WgradParams& p, cudaStream_t s) {
  using Cfg = WgradCfg<BN, PASSES>;
  auto kern = wgrad_tf32_kernel<BN, PASSES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = p.num_m_tiles * p.num_n_tiles * p.splits;
  kern<<<grid, WG_THREADS, Cfg::SMEM_BYTES, s>>>(tDZ, tX, p);
  return check_launch("wgrad_tf32_kernel");
}

static void wgrad_plan(int64_t M, int64_t N, int64_t K, int* bn, int* mt, int* nt, int* splits, int* rows_per_split,
                       int64_t* ldp) {
  *bn = K > 128 ? 256 : 128;
  *mt = (int)((N + WG_BM - 1) / WG_BM);
  *nt = (int)((K + *bn - 1) / *bn);
  const int tiles = *mt * *nt;
  int sp = sm_count() / tiles;
  if (sp < 1) sp = 1;
  int64_t rps = ((M + sp - 1) / sp + WG_BKB - 1) / WG_BKB * WG_BKB;
  sp = (int)((M + rps - 1) / rps);
  *splits = sp;
  *rows_per_split = (int)rps;
  *ldp = (int64_t)*nt * *bn;
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

size_t cusrl_b200_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int bn, mt, nt, splits, rps;
  int64_t ldp;
  wgrad_plan(M, N, K, &bn, &mt, &nt, &splits, &rps, &ldp);
  const size_t partial = (size_t)splits * mt * WG_BM * ldp * sizeof(float);
  const size_t colsum = (size_t)1184 * (size_t)N * sizeof(float);
  return partial + colsum + 256;
}

size_t cusrl_b200_colsum_workspace_bytes(int64_t N) { return N > 0 ? (size_t)1184 * (size_t)N * sizeof(float) : 0; }

int cusrl_b200_colsum_f32(const float* dZ, int64_t lddz, int64_t M, int64_t N, float* db, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream) {
  CUSRL_REQUIRE(dZ && db && workspace, CUSRL_B200_EINVAL, "colsum: null pointer");
  CUSRL_REQUIRE(M > 0 && N > 0 && lddz >= N && M < (1ll << 31), CUSRL_B200_EINVAL, "colsum: bad sizes");
  CUSRL_REQUIRE(workspace_bytes >= cusrl_b200_colsum_workspace_bytes(N), CUSRL_B200_ESCRATCH, "colsum: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int nb = sm_count() * 8;
  if (nb > 1184) nb = 1184;
  int rows_per_block = (int)((M + nb - 1) / nb);
  nb = (int)((M + rows_per_block - 1) / rows_per_block);
  colsum_partial_kernel<<<nb, 256, 0, s>>>(dZ, lddz, (int)M, (int)N, rows_per_block, (float*)workspace);
  if (int e = check_launch("colsum_partial_kernel")) return e;
  return colsum_finalize((const float*)workspace, nb, (int)N, db, accumulate, s);
}

int cusrl_b200_linear_wgrad_tf32(const float* dZ, int64_t lddz, const float* X, int64_t ldx, float* dW, int64_t lddw,
                                 float* db, int64_t M, int64_t N, int64_t K, int precision, int accumulate,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  CUSRL_REQUIRE(dZ && X && dW && workspace, CUSRL_B200_EINVAL, "wgrad: null pointer");
  CUSRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N <= 65536 && K <= 65536, CUSRL_B200_EINVAL,
                "wgrad: bad problem size");
  CUSRL_REQUIRE(precision == 1 || precision == 3, CUSRL_B200_EINVAL, "wgrad: precision must be 1 or 3");
  CUSRL_REQUIRE((lddz % 4) == 0 && (ldx % 4) == 0 && lddz >= N && ldx >= K && lddw >= K, CUSRL_B200_EALIGN,
                "wgrad: activation leading dimensions must be multiples of 4 floats and cover the row");
  CUSRL_REQUIRE(aligned_to(dZ, 16) && aligned_to(X, 16) && aligned_to(workspace, 16), CUSRL_B200_EALIGN,
                "wgrad: dZ, X and workspace must be 16-byte aligned");
  CUSRL_REQUIRE(workspace_bytes >= cusrl_b200_wgrad_workspace_bytes(M, N, K), CUSRL_B200_ESCRATCH, "wgrad: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int bn, mt, nt, splits, rps;
  int64_t ldp;
  wgrad_plan(M, N, K, &bn, &mt, &nt, &splits, &rps, &ldp);
  CUtensorMap tDZ, tX;
  if (int e = encode_tmap_2d_f32(&tDZ, dZ, (uint64_t)N, (uint64_t)M, (uint64_t)lddz, WG_CHUNK, WG_BKB, TMAP_SW128_ATOM32)) return e;
  if (int e = encode_tmap_2d_f32(&tX, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, WG_CHUNK, WG_BKB, TMAP_SW128_ATOM32)) return e;
  WgradParams p{};
  p.partial = (float*)workspace, p.ldp = ldp, p.M = (int)M, p.N = (int)N, p.K = (int)K;
  p.num_m_tiles = mt, p.num_n_tiles = nt, p.splits = splits, p.rows_per_split = rps;
  int e = CUSRL_B200_EUNSUPPORTED;
  if (bn == 256 && precision == 3) e = launch_wgrad<256, 3>(tDZ, tX, p, s);
  else if (bn == 128 && precision == 3) e = launch_wgrad<128, 3>(tDZ, tX, p, s);
  else if (bn == 256 && precision == 1) e = launch_wgrad<256, 1>(tDZ, tX, p, s);
  else if (bn == 128 && precision == 1) e = launch_wgrad<128, 1>(tDZ, tX, p, s);
  if (e) return e;
  const int64_t split_stride = (int64_t)mt * WG_BM * ldp;
  int blocks = (int)((N * K + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  wgrad_reduce_kernel<<<blocks, 256, 0, s>>>(p.partial, splits, split_stride, ldp, dW, lddw, (int)N, (int)K, accumulate);
  if (int e2 = check_launch("wgrad_reduce_kernel")) return e2;
  if (db) {
    float* cs = p.partial + (int64_t)splits * split_stride;
    int nb = sm_count() * 8;
    if (nb > 1184) nb = 1184;
    int rows_per_block = (int)((M + nb - 1) / nb);
    nb = (int)((M + rows_per_block - 1) / rows_per_block);
    colsum_partial_kernel<<<nb, 256, 0, s>>>(dZ, lddz, (int)M, (int)N, rows_per_block, cs);
    if (int e3 = check_launch("colsum_partial_kernel")) return e3;
    if (int e4 = colsum_finalize(cs, nb, (int)N, db, accumulate, s)) return e4;
  }
  return 0;
}

}  // extern "C"
