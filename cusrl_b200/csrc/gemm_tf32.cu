// K6: dense layers of the MLP actor / critic on the 5th-generation tensor cores.
//
//   D[M,N] = epilogue( A[M,K] @ B[N,K]^T ),  A and B both K-major (row-major with K contiguous), fp32 in HBM.
//
// * forward   : A = X,  B = W        epilogue  y = act(d + bias)                      (nn.Linear + ELU/ReLU)
// * data grad : A = dZ, B = W^T copy epilogue  dx = d * act'(y_prev)                  (autograd of the layer below)
//
// Persistent, warp-specialised CTAs (one per SM) launched as 2-CTA clusters:
//   warp 0       TMA producer  : cp.async.bulk.tensor 128B-swizzled tiles -> shared-memory ring, mbarrier tx.  The two
//                                CTAs of a cluster work on two M tiles that share the same weight tile: each loads its
//                                own A tile and HALF of the B tile, multicast to both CTAs, so every weight byte leaves
//                                L2 once per cluster (the weights are re-read for every M tile: L2 bandwidth, not HBM,
//                                is what bounds the main loop otherwise)
//   warp 1       MMA issuer    : one elected thread issues tcgen05.mma.kind::tf32, accumulators in TMEM
//   warp 2       TMEM allocator
//   warps 4-11   epilogue      : tcgen05.ld -> registers -> bias/activation -> swizzled staging -> bulk tensor store
//   warps 12-15  splitter (3xTF32 only): writes lo = A - trunc_tf32(A) next to the A tile (A itself is untouched)
// Two TMEM accumulator buffers (2 x BN columns) let the epilogue of tile i overlap the main loop of tile i+1.
//
// Precision.  The reference computes these layers with fp32 SGEMM.  PASSES = 3 is the error-compensated
// "3xTF32" scheme: x = hi + lo with hi, lo both TF32-representable, D = Ahi*Bhi + Ahi*Blo + Alo*Bhi (fp32
// accumulate), which restores ~fp32 accuracy at one third of the TF32 tensor rate.  The tensor core truncates its
// fp32 inputs to TF32 (measured, tools/tf32_rounding_probe.py), so the raw fp32 activation tile IS the hi operand
// and lo = x - (x & ~0x1fff) is exact; weights are pre-split once per optimizer step.
//
// What bounds it (round 1, A/B experiments with parts of the kernel disabled, M = 393216, K = 512, N = 256):
//   * 1xTF32: 227 us; main loop alone (epilogue off) 162 us  = HBM (1.2 GB at ~5.3 TB/s);
//   * 3xTF32: 437 us; main loop alone 377 us = 1.12 us per 32-deep k-block against 0.78 us of MMA time.  Per k-block
//     a CTA INGESTS 16 KB of A + 64 KB of W hi/lo (~85-95 GB/s per SM, the L2 -> SM port) and its shared memory serves
//     80 KB of TMA writes + 32 KB of splitter traffic + 144 KB of MMA operand reads (12 MMAs x 12 KB), i.e. 256 KB per
//     1536 MMA cycles = 167 B/clk against 128 B/clk: the SM's own data paths, not HBM and not the tensor pipe
//     (58 % busy under ncu), set the pace.  Tried and measured without gain: 4 stages of 16-float k-blocks, an
//     in-kernel weight split (fewer bytes in, more shared-memory traffic), cta_group::2 (a CTA-pair MMA variant, removed in round 2; the peer's
//     half of B still crosses the SM boundary).  What DID matter was the epilogue: libdevice expm1f (-25 %) and
//     row-per-thread 16-byte global stores.
// PASSES = 1 is plain single-pass TF32.
#include "gemm_common.cuh"

namespace cusrl_b200 {

// k-block depth.  16-float (SWIZZLE_64B) k-blocks give the 3xTF32 kernel 4 pipeline stages instead of 2 in the same
// shared memory; measured on B200 it changes nothing (474/426/165 us vs 485/437/153 us for the three Anymal-C layers),
// so the pipeline is not latency-bound and the simpler 32-float blocks stay.  Return 16 for PASSES == 3 to re-test.
template <int PASSES>
constexpr int gemm_bk() { return 32; }

template <int BN, int PASSES>
struct GemmCfg {
  static constexpr int BK = gemm_bk<PASSES>();
  static constexpr uint32_t LAYOUT = BK == 32 ? 2u : 4u;   // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static constexpr uint32_t SBO = 8 * BK * 4;               // stride between 8-row groups
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (PASSES == 3 ? 2 : 1);
  static constexpr int STAGES = (kSmemBudget / STAGE_BYTES) > 6 ? 6 : (kSmemBudget / STAGE_BYTES);
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kEpiStageBytes + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
  static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

template <int BN, int PASSES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, PASSES>;
  constexpr int BK = Cfg::BK;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int HALF_B_BYTES = Cfg::B_BYTES / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  auto sA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sB = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES * (PASSES == 3 ? 2 : 1); };
  auto sBlo = [&](int s) { return sB(s) + Cfg::B_BYTES; };
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;  // 8 epilogue warps x one 32x32 fp32 chunk (SWIZZLE_128B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + kEpiStageBytes);
  uint64_t* full = bars;                  // TMA bytes landed (own A + both halves of B)
  uint64_t* split = bars + STAGES;        // A tile split into hi/lo (3xTF32)
  uint64_t* empty = bars + 2 * STAGES;    // MMAs of BOTH CTAs reading the stage retired
  uint64_t* tfull = bars + 3 * STAGES;    // accumulator complete
  uint64_t* tempty = tfull + 2;           // accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_k_blocks = (p.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (PASSES == 3) tma_prefetch_desc(&tmBlo);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 4);
      mbar_init(&empty[s], 2);  // one tcgen05.commit from each CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers must be initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (N tile, pair of M tiles); this CTA takes M tile 2*pair + rank (possibly past the end: TMA zero-fills,
  // the epilogue skips the rows, and the CTA still loads and multicasts its half of B for its peer)
  auto item_m0 = [&](int item) { return ((item / p.num_n_tiles) * 2 + (int)rank) * BM; };
  auto item_n0 = [&](int item) { return (item % p.num_n_tiles) * BN; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const int m0 = item_m0(item), n0 = item_n0(item);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES * (PASSES == 3 ? 2 : 1));
          tma_load_2d(sA(s), &tmA, kb * BK, m0, &full[s]);
          tma_load_2d_mc(sB(s) + rank * HALF_B_BYTES, &tmB, kb * BK, n0 + (int)rank * (BN / 2), &full[s], (uint16_t)3);
          if (PASSES == 3)
            tma_load_2d_mc(sBlo(s) + rank * HALF_B_BYTES, &tmBlo, kb * BK, n0 + (int)rank * (BN / 2), &full[s], (uint16_t)3);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int local = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters, ++local) {
        const int a = local & 1;
        const uint32_t aph = (local >> 1) & 1;
        mbar_wait(&tempty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA(s)), b_addr = smem_u32(sB(s));
          const uint32_t alo_addr = smem_u32(sAlo(s)), blo_addr = smem_u32(sBlo(s));
          // The tensor core TRUNCATES fp32 inputs to TF32 (measured: tools/tf32_rounding_probe.py), so the raw fp32 A
          // tile is the "hi" operand as it lands; the two passes that do not need A_lo are issued immediately and
          // the splitter is off the critical path.  K-major SWIZZLE_128B operands: 8-row groups are 1024 B apart
          // (SBO); advancing K by 8 tf32 = +32 B.
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t off = (uint32_t)k * UMMA_K * 4;
            const uint64_t da = make_smem_desc_sw128(a_addr + off, 16, Cfg::SBO, Cfg::LAYOUT);
            const uint64_t db = make_smem_desc_sw128(b_addr + off, 16, Cfg::SBO, Cfg::LAYOUT);
            mma_tf32_ss(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (PASSES == 3) mma_tf32_ss(d_tmem, da, make_smem_desc_sw128(blo_addr + off, 16, Cfg::SBO, Cfg::LAYOUT), idesc, 1u);
          }
          if (PASSES == 3) {
            mbar_wait(&split[s], ph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t off = (uint32_t)k * UMMA_K * 4;
              mma_tf32_ss(d_tmem, make_smem_desc_sw128(alo_addr + off, 16, Cfg::SBO, Cfg::LAYOUT), make_smem_desc_sw128(b_addr + off, 16, Cfg::SBO, Cfg::LAYOUT),
                          idesc, 1u);
            }
          }
          // the stage may be refilled (by either CTA's multicast) only when both CTAs are done reading it
          mma_commit_mc(&empty[s], (uint16_t)3);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
        mma_commit(&tfull[a]);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================== epilogue ==========================================
    const int ew = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    const int half = (warp - 4) >> 2;   // which half of the tile's columns
    uint8_t* stg = epi_stage + (warp - 4) * kEpiWarpBytes;
    int local = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters, ++local) {
      const int a = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      const int row0 = item_m0(item) + ew * 32, n0 = item_n0(item);
      mbar_wait(&tfull[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(ew * 32) << 16);
      if (row0 < p.M) {
#pragma unroll 1
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32)
          if (n0 + c0 < p.N) epilogue_chunk<EPI>(p, &tmOut, stg, taddr + (uint32_t)c0, row0, n0 + c0, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
    if (lane == 0) tma_store_wait_read();  // shared memory must outlive the last bulk store's read
  } else if (PASSES == 3 && warp >= 12) {
    // ===================================== splitter (3xTF32) =================================
    const int t = threadIdx.x - 384;  // 0..127
    int s = 0;
    uint32_t ph = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full[s], ph);
        const uint4* hi = reinterpret_cast<const uint4*>(sA(s));
        float4* lo = reinterpret_cast<float4*>(sAlo(s));
#pragma unroll
        for (int i = 0; i < Cfg::A_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;  // elementwise at identical offsets: the swizzled layout is preserved
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
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // do not exit while the peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight operand preparation: hi / lo split (and transposed copies for the data-gradient GEMM) of a small
// weight matrix, run once per optimizer step.
// ------------------------------------------------------------------------------------------------
__global__ void weight_prep_kernel(const float* __restrict__ w, int N, int K, float* __restrict__ hi, float* __restrict__ lo,
                                   int ld, float* __restrict__ hi_t, float* __restrict__ lo_t, int ldt) {
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i - n * K;
    const float x = w[i];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const float l = x - h;
    hi[(int64_t)n * ld + k] = h;
    lo[(int64_t)n * ld + k] = l;
    if (hi_t) {
      hi_t[(int64_t)k * ldt + n] = h;
      lo_t[(int64_t)k * ldt + n] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int encode_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                       uint32_t box_inner, uint32_t box_outer, int swizzle) {
  EncodeTiledFn enc = get_encoder();
  CUSRL_REQUIRE(enc != nullptr, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == TMAP_SW128_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                : (swizzle == TMAP_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUSRL_REQUIRE(r == CUDA_SUCCESS, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// Un-swizzled 2D map over a row-major [outer, inner] array of 1-byte or 4-byte elements (HBM-bound staging kernels:
// rows land in shared memory exactly as they lie in global memory).
int encode_tmap_2d_plain(CUtensorMap* map, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld_bytes,
                         uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encoder();
  CUSRL_REQUIRE(enc != nullptr, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CUSRL_REQUIRE(elem_bytes == 1 || elem_bytes == 4, CUSRL_B200_EINVAL, "tensor map: element size must be 1 or 4 bytes");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUSRL_REQUIRE(r == CUDA_SUCCESS, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled (plain) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tc


int gemm_grid_ctas(int num_items) {
  const int max_clusters = sm_count() / 2;
  return 2 * (num_items < max_clusters ? num_items : max_clusters);
}

template <int BN, int PASSES, int EPI>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tBlo, const CUtensorMap& tOut,
                       const GemmParams& p, cudaStream_t s) {
  using Cfg = GemmCfg<BN, PASSES>;
  auto kern = gemm_tf32_kernel<BN, PASSES, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gemm: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  kern<<<gemm_grid_ctas(p.num_items), kGemmThreads, Cfg::SMEM_BYTES, s>>>(tA, tB, tBlo, tOut, p);  // cluster dims (2,1,1) are compiled in
  return check_launch("gemm_tf32_kernel");
}

static int gemm_dispatch(int bn, int precision, int epi, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tBlo,
                         const CUtensorMap& tOut, const GemmParams& p, cudaStream_t s);
// db (+)= sum over `nblocks` partial rows, fixed order (gemm_wgrad_tf32.cu)
int colsum_finalize(const float* partial, int nblocks, int N, float* db, int accumulate, cudaStream_t s);

// Shared implementation of forward (EPI_BIAS_ACT) and data-gradient (EPI_ACT_GRAD) calls.
static int gemm_kmajor(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* out, int64_t ldo,
                       const float* bias, const float* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int act,
                       int precision, int epi, void* stream, float* colsum_db = nullptr, int colsum_accumulate = 0,
                       void* workspace = nullptr, size_t workspace_bytes = 0) {
  CUSRL_REQUIRE(A && Bhi && out, CUSRL_B200_EINVAL, "linear: null pointer");
  CUSRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N <= 65536 && K <= 65536, CUSRL_B200_EINVAL,
                "linear: bad problem size");
  CUSRL_REQUIRE(precision == 1 || (precision == 3 && Blo), CUSRL_B200_EINVAL, "linear: precision must be 1 or 3 (3 needs W_lo)");
  CUSRL_REQUIRE(act >= 0 && act <= 2, CUSRL_B200_EINVAL, "linear: unknown activation code %d", act);
  CUSRL_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0 && (ldo % 4) == 0 && (N % 4) == 0 && (!aux || (ldaux % 4) == 0),
                CUSRL_B200_EALIGN, "linear: leading dimensions and N must be multiples of 4 floats");
  CUSRL_REQUIRE(lda >= K && ldb >= K && ldo >= N, CUSRL_B200_EINVAL, "linear: leading dimension smaller than the row");
  CUSRL_REQUIRE(aligned_to(A, 16) && aligned_to(Bhi, 16) && aligned_to(out, 16) && (!Blo || aligned_to(Blo, 16)) &&
                    (!bias || aligned_to(bias, 16)) && (!aux || aligned_to(aux, 16)),
                CUSRL_B200_EALIGN, "linear: pointers must be 16-byte aligned");
  const int bn = N > 128 ? 256 : 128;
  CUtensorMap tA, tB, tBlo;
  const bool deep = gemm_bk<3>() == 16 && precision == 3;  // 16-float k-blocks, SWIZZLE_64B
  const uint32_t bk = deep ? 16 : 32;
  const int sw = deep ? TMAP_SW64 : TMAP_SW128;
  if (int e = encode_tmap_2d_f32(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, bk, BM, sw)) return e;
  // each CTA of a cluster loads (and multicasts) one half of the B tile: box = bn/2 rows
  if (int e = encode_tmap_2d_f32(&tB, Bhi, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bk, (uint32_t)bn / 2, sw)) return e;
  if (int e = encode_tmap_2d_f32(&tBlo, Blo ? Blo : Bhi, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bk, (uint32_t)bn / 2, sw)) return e;
  // epilogue: 32 x 32 fp32 chunks staged in shared memory and written by bulk tensor stores
  CUtensorMap tOut;
  if (int e = encode_tmap_2d_f32(&tOut, out, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, 32, 32)) return e;
  GemmParams p{};
  p.out = out, p.ldo = ldo, p.bias = bias, p.aux = aux, p.ldaux = ldaux;
  p.M = (int)M, p.N = (int)N, p.K = (int)K, p.act = act;
  p.num_m_tiles = (int)((M + BM - 1) / BM);
  p.num_n_tiles = (int)((N + bn - 1) / bn);
  p.num_items = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = gemm_grid_ctas(p.num_items);
  if (colsum_db) {
    const size_t need = (size_t)ctas * 4 * (size_t)N * sizeof(float);
    CUSRL_REQUIRE(workspace && workspace_bytes >= need && aligned_to(workspace, 16), CUSRL_B200_ESCRATCH,
                  "linear_dgrad: workspace too small for the bias-gradient partials");
    cudaError_t me = cudaMemsetAsync(workspace, 0, need, s);
    CUSRL_REQUIRE(me == cudaSuccess, (int)me, "linear_dgrad: cudaMemsetAsync: %s", cudaGetErrorString(me));
    p.colsum = (float*)workspace;
  }
  int rc = gemm_dispatch(bn, precision, epi, tA, tB, tBlo, tOut, p, s);
  if (rc || !colsum_db) return rc;
  return colsum_finalize((const float*)workspace, ctas * 4, (int)N, colsum_db, colsum_accumulate, s);
}

static int gemm_dispatch(int bn, int precision, int epi, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tBlo,
                         const CUtensorMap& tOut, const GemmParams& p, cudaStream_t s) {
#define CUSRL_GEMM_CASE(BN_, P_, E_) \
  if (bn == BN_ && precision == P_ && epi == E_) return launch_gemm<BN_, P_, E_>(tA, tB, tBlo, tOut, p, s);
  CUSRL_GEMM_CASE(256, 3, EPI_BIAS_ACT)
  CUSRL_GEMM_CASE(128, 3, EPI_BIAS_ACT)
  CUSRL_GEMM_CASE(256, 1, EPI_BIAS_ACT)
  CUSRL_GEMM_CASE(128, 1, EPI_BIAS_ACT)
  CUSRL_GEMM_CASE(256, 3, EPI_ACT_GRAD)
  CUSRL_GEMM_CASE(128, 3, EPI_ACT_GRAD)
  CUSRL_GEMM_CASE(256, 1, EPI_ACT_GRAD)
  CUSRL_GEMM_CASE(128, 1, EPI_ACT_GRAD)
#undef CUSRL_GEMM_CASE
  set_last_error("linear: no kernel for this configuration");
  return CUSRL_B200_EUNSUPPORTED;
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_weight_prep_f32(const float* W, int64_t N, int64_t K, float* hi, float* lo, int64_t ld, float* hi_t,
                               float* lo_t, int64_t ldt, void* stream) {
  CUSRL_REQUIRE(W && hi && lo, CUSRL_B200_EINVAL, "weight_prep: null pointer");
  CUSRL_REQUIRE(N > 0 && K > 0 && ld >= K && N * K < (1ll << 31), CUSRL_B200_EINVAL, "weight_prep: bad sizes");
  CUSRL_REQUIRE((hi_t == nullptr) == (lo_t == nullptr) && (!hi_t || ldt >= N), CUSRL_B200_EINVAL,
                "weight_prep: transposed outputs must be given together with ldt >= N");
  const int total = (int)(N * K);
  int blocks = (total + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  weight_prep_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, (int)N, (int)K, hi, lo, (int)ld, hi_t, lo_t, (int)ldt);
  return check_launch("weight_prep_kernel");
}

int cusrl_b200_linear_fwd_tf32(const float* X, int64_t ldx, const float* W_hi, const float* W_lo, int64_t ldw,
                               const float* bias, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act,
                               int precision, void* stream) {
  return gemm_kmajor(X, ldx, W_hi, W_lo, ldw, Y, ldy, bias, nullptr, 0, M, N, K, act, precision, EPI_BIAS_ACT, stream);
}

size_t cusrl_b200_dgrad_workspace_bytes(int64_t K) {
  if (K <= 0) return 0;
  return (size_t)sm_count() * 4 * (size_t)K * sizeof(float);
}

int cusrl_b200_linear_dgrad_tf32(const float* dY, int64_t lddy, const float* WT_hi, const float* WT_lo, int64_t ldwt,
                                 const float* Xact, int64_t ldxa, float* dX, int64_t lddx, int64_t M, int64_t N, int64_t K,
                                 int act, int precision, float* db_below, int accumulate, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  // dX[M,K] = dY[M,N] @ W[N,K]: as a K-major GEMM the reduction runs over N and B is the transposed copy WT[K,N]
  return gemm_kmajor(dY, lddy, WT_hi, WT_lo, ldwt, dX, lddx, nullptr, Xact, ldxa, M, /*N=*/K, /*K=*/N, act, precision,
                     EPI_ACT_GRAD, stream, db_below, accumulate, workspace, workspace_bytes);
}

}  // extern "C"
