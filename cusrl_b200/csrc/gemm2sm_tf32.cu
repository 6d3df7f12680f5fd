// K6, CTA-pair variant: the K-major dense-layer GEMM with tcgen05.mma.cta_group::2.
//
// Why: the 1-SM kernel (gemm_tf32.cu) is bound by shared-memory bandwidth -- every tf32 MMA (128 x 256 x 8) re-reads
// 4 KB of A and 8 KB of B from shared memory per 128 cycles.  With cta_group::2 one MMA spans a CTA pair (M = 256):
// each SM supplies its own 128 rows of A and only HALF of the B tile from its shared memory (the halves are exchanged
// over the SM-pair link), i.e. 8 KB instead of 12 KB per MMA-time, and the B half also halves the stage footprint
// (3 pipeline stages instead of 2 for 3xTF32).
//
// Roles per CTA (640 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator
// (cta_group::2 alloc in both CTAs), warps 4-11 epilogue (own 128 accumulator rows), warps 12-15 splitter (3xTF32).
// Barriers: A tiles complete on a LOCAL barrier (the local splitter waits on it); B halves of both CTAs complete on
// the LEADER's barrier (cp.async.bulk.tensor.cta_group::2); splitters and epilogues of both CTAs arrive on the
// leader's barriers (remote arrive for the peer); tcgen05.commit is multicast to both CTAs.
#include "gemm_common.cuh"

namespace cusrl_b200 {

constexpr int k2smThreads = 512;

template <int BN, int PASSES>
struct Gemm2smCfg {
  static constexpr int BK = 32;
  static constexpr int A_BYTES = BM * BK * 4;            // this CTA's 128 rows
  static constexpr int BH_BYTES = (BN / 2) * BK * 4;     // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = (A_BYTES + BH_BYTES) * (PASSES == 3 ? 2 : 1);
  static constexpr int STAGES = (kSmemBudget / STAGE_BYTES) > 6 ? 6 : (kSmemBudget / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kEpiStageBytes + 512 + 1024;
  static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

template <int BN, int PASSES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2smThreads, 1)
gemm2sm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut,
                    const GemmParams p) {
  using Cfg = Gemm2smCfg<BN, PASSES>;
  constexpr int BK = Cfg::BK;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage layout: [A | B half | A_lo | B_lo half]
  auto sA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sB = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::BH_BYTES; };
  auto sBlo = [&](int s) { return sAlo(s) + Cfg::A_BYTES; };
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + kEpiStageBytes);
  uint64_t* afull = bars;                 // local : this CTA's A tile landed                      (3xTF32)
  uint64_t* bfull = bars + STAGES;        // leader: B halves (1xTF32: A tiles too) of both CTAs landed
  uint64_t* split = bars + 2 * STAGES;    // leader: A_lo written by the splitters of both CTAs    (3xTF32)
  uint64_t* empty = bars + 3 * STAGES;    // local : MMAs reading the stage retired (multicast commit)
  uint64_t* apeer = bars + 4 * STAGES;    // leader: the PEER's A tile landed (relayed by the peer's warp 3)      (3xTF32)
  uint64_t* tfull = bars + 5 * STAGES;    // local : accumulator complete (multicast commit)
  uint64_t* tempty = tfull + 2;           // leader: accumulator drained by the epilogues of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k_blocks = (p.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (PASSES == 3) tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&bfull[s], 1);
      mbar_init(&split[s], 8);    // 4 splitter warps x 2 CTAs
      mbar_init(&empty[s], 1);
      mbar_init(&apeer[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 16);  // 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_m0 = [&](int item) { return ((item / p.num_n_tiles) * 2 + (int)rank) * BM; };
  auto item_n0 = [&](int item) { return (item % p.num_n_tiles) * BN; };

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const int m0 = item_m0(item), n0 = item_n0(item) + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          const uint32_t leader_bfull = mapa_u32(smem_u32(&bfull[s]), 0);
          if (PASSES == 3) {
            mbar_expect_tx(&afull[s], Cfg::A_BYTES);
            tma_load_2d(sA(s), &tmA, kb * BK, m0, &afull[s]);
            if (leader) mbar_expect_tx(&bfull[s], 4 * Cfg::BH_BYTES);  // hi + lo halves of both CTAs
            tma_load_2d_2sm(sB(s), &tmB, kb * BK, n0, leader_bfull);
            tma_load_2d_2sm(sBlo(s), &tmBlo, kb * BK, n0, leader_bfull);
          } else {
            if (leader) mbar_expect_tx(&bfull[s], 2 * (Cfg::A_BYTES + Cfg::BH_BYTES));
            tma_load_2d_2sm(sA(s), &tmA, kb * BK, m0, leader_bfull);
            tma_load_2d_2sm(sB(s), &tmB, kb * BK, n0, leader_bfull);
          }
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN, 0, 0);  // M = 256 across the pair
      int s = 0;
      uint32_t ph = 0;
      int local = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters, ++local) {
        const int a = local & 1;
        const uint32_t aph = (local >> 1) & 1;
        mbar_wait_cluster(&tempty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait_cluster(&bfull[s], ph);
          if (PASSES == 3) {
            mbar_wait(&afull[s], ph);
            mbar_wait_cluster(&apeer[s], ph);
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA(s)), b_addr = smem_u32(sB(s));
          const uint32_t alo_addr = smem_u32(sAlo(s)), blo_addr = smem_u32(sBlo(s));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t off = (uint32_t)k * UMMA_K * 4;
            const uint64_t da = make_smem_desc_sw128(a_addr + off, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + off, 16, 1024);
            mma_tf32_ss_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (PASSES == 3) mma_tf32_ss_2sm(d_tmem, da, make_smem_desc_sw128(blo_addr + off, 16, 1024), idesc, 1u);
          }
          if (PASSES == 3) {
            // the raw fp32 A tiles are the "hi" operands as they land (the tensor core truncates to TF32); only the
            // third pass needs the splitters of both CTAs
            mbar_wait_cluster(&split[s], ph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t off = (uint32_t)k * UMMA_K * 4;
              mma_tf32_ss_2sm(d_tmem, make_smem_desc_sw128(alo_addr + off, 16, 1024), make_smem_desc_sw128(b_addr + off, 16, 1024),
                              idesc, 1u);
            }
          }
          mma_commit_2sm_mc(&empty[s], (uint16_t)3);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
        mma_commit_2sm_mc(&tfull[a], (uint16_t)3);
      }
    }
  } else if (warp == 3) {
    // ===================================== peer: relay "my A tile landed" to the leader ======
    if (PASSES == 3 && !leader && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&afull[s], ph);
          mbar_arrive_cluster(mapa_u32(smem_u32(&apeer[s]), 0));
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================== epilogue (both CTAs, own 128 rows) ================
    const int ew = warp & 3;
    const int half = (warp - 4) >> 2;
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
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[a]), 0));
    }
    if (lane == 0) tma_store_wait_read();
  } else if (PASSES == 3 && warp >= 12) {
    // ===================================== splitter (3xTF32, both CTAs) ======================
    const int t = threadIdx.x - 384;  // 0..127
    int s = 0;
    uint32_t ph = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&afull[s], ph);
        const uint4* hi = reinterpret_cast<const uint4*>(sA(s));
        float4* lo = reinterpret_cast<float4*>(sAlo(s));
#pragma unroll
        for (int i = 0; i < Cfg::A_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;
          const uint4 x = hi[idx];
          lo[idx] = make_float4(__uint_as_float(x.x) - __uint_as_float(x.x & 0xffffe000u),
                                __uint_as_float(x.y) - __uint_as_float(x.y & 0xffffe000u),
                                __uint_as_float(x.z) - __uint_as_float(x.z & 0xffffe000u),
                                __uint_as_float(x.w) - __uint_as_float(x.w & 0xffffe000u));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&split[s]), 0));
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int BN, int PASSES, int EPI>
static int launch_one(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tBlo, const CUtensorMap& tOut,
                      const GemmParams& p, cudaStream_t s) {
  using Cfg = Gemm2smCfg<BN, PASSES>;
  auto kern = gemm2sm_tf32_kernel<BN, PASSES, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gemm2sm: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int max_clusters = sm_count() / 2;
  const int clusters = p.num_items < max_clusters ? p.num_items : max_clusters;
  kern<<<2 * clusters, k2smThreads, Cfg::SMEM_BYTES, s>>>(tA, tB, tBlo, tOut, p);
  return check_launch("gemm2sm_tf32_kernel");
}

int launch_gemm_2sm(int bn, int precision, int epi, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tBlo,
                    const CUtensorMap& tOut, const GemmParams& p, cudaStream_t s) {
#define CUSRL_CASE(BN_, P_, E_) \
  if (bn == BN_ && precision == P_ && epi == E_) return launch_one<BN_, P_, E_>(tA, tB, tBlo, tOut, p, s);
  CUSRL_CASE(256, 3, EPI_BIAS_ACT)
  CUSRL_CASE(128, 3, EPI_BIAS_ACT)
  CUSRL_CASE(256, 1, EPI_BIAS_ACT)
  CUSRL_CASE(128, 1, EPI_BIAS_ACT)
  CUSRL_CASE(256, 3, EPI_ACT_GRAD)
  CUSRL_CASE(128, 3, EPI_ACT_GRAD)
  CUSRL_CASE(256, 1, EPI_ACT_GRAD)
  CUSRL_CASE(128, 1, EPI_ACT_GRAD)
#undef CUSRL_CASE
  set_last_error("gemm2sm: no kernel for this configuration");
  return CUSRL_B200_EUNSUPPORTED;
}

}  // namespace cusrl_b200
