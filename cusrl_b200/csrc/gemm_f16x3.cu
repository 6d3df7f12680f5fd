// K6, precision 2 ("f16x3"): the dense layers on tcgen05 kind::f16 with fp16 hi / lo split operands (f16x3_common.cuh).
//
//   D[M,N] = epilogue( (Ah Bh^T + Ah Bl^T + Al Bh^T) / (sA sB) ),   A [M,K] and B [N,K] K-major fp16 pairs in HBM
//
// * forward   : A = X pair,  B = W pair      y = act(d + bias)        -> fp16 pair (next layer's operand) or fp32 (heads)
// * data grad : A = dZ pair, B = W^T pair    dx = d * act'(y_below)   -> fp16 pair (+ column sums = bias gradient below)
//
// Same persistent warp-specialised structure as gemm_tf32.cu (2-CTA clusters, each CTA loads its own A tile and half of
// the B tile, multicast to both), with three differences that follow from the operand format:
//   * both halves of BOTH operands arrive by TMA already split -- no splitter warps, no generic-proxy traffic on the
//     operand tiles (round 1: 48 % LSU shared-memory wavefronts on the weight gradient, 38 % here);
//   * a k-block is 64 halves (one 128-byte swizzle span), i.e. TWICE the reduction depth per stage for the same bytes, and
//     an MMA covers K = 16: per unit of K half the instructions, half the shared-memory operand reads, half the tensor time;
//   * the epilogue rescales by the exact power of two 1/(sA sB) and, for pair outputs, splits the result again with the
//     scale of the output's ANALYTIC bound (computed here from device scalars, published for the consumers).
#include "f16x3_common.cuh"
#include "gemm_common.cuh"

namespace cusrl_b200 {

constexpr int FBK = 64;            // halves per k-block = 128 bytes
constexpr int F_UMMA_K = 16;       // halves per tcgen05.mma
// warps: 0 TMA, 1 MMA, 2 TMEM, 3 idle, 4.. epilogue.  The epilogue (TMEM -> activation -> fp16 split -> store) is what bounds
// the layers with many outputs per reduction step: with 128 x 256 tiles only 8 epilogue warps fit (4 KB of staging each next
// to two 96 KB stages) and the 235->512 layer takes 453 us at M = 393 216; 128 x 128 tiles leave room for 16.
template <int BN>
constexpr int f16_epi_warps() { return BN == 128 ? 16 : 8; }
template <int BN>
constexpr int f16_threads() { return 128 + 32 * f16_epi_warps<BN>(); }

enum { OUT_F32 = 0, OUT_PAIR = 1 };

struct F16GemmParams {
  const float* bias;        // EPI_BIAS_ACT (may be null)
  const __half* aux_hi;     // EPI_ACT_GRAD: pair of the post-activation output of the layer below (null: plain product)
  const __half* aux_lo;
  int64_t ldaux;            // halves
  const float* bound_a;     // device scalars: bound of A, weight statistics (float[4]), bound of aux
  const float* wstats;
  const float* bound_aux;
  float* bound_out;         // OUT_PAIR: the analytic bound of the output, written by CTA 0
  int M, N, K, act;
  int num_m_tiles, num_n_tiles, num_items;
  float* colsum;            // EPI_ACT_GRAD, optional: [gridDim.x][4][N] column sums of the fp32 output values
  int prefetch;             // k-blocks the L2 prefetch of the A tiles runs ahead of the loads (0 = off)
};

template <int BN>
struct F16Cfg {
  static constexpr int A_BYTES = BM * FBK * 2;     // 16 KB
  static constexpr int B_BYTES = BN * FBK * 2;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);   // hi + lo of both operands
  static constexpr int EPI_WARPS = f16_epi_warps<BN>();
  static constexpr int EPI_BYTES = EPI_WARPS * kEpiWarpBytes;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES = ((226 * 1024 - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES) > 4 ? 4 : ((226 * 1024 - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

// all 32 column sums of a 32 x 32 block held one row per lane: after five exchange rounds lane l owns column l's sum
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n / 2) {
        const float send = up ? v[j] : v[j + n / 2];
        const float keep = up ? v[j + n / 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
  }
  return v[0];
}

// fp16 hi / lo split of two fp32 values with ONE conversion instruction per half2: hi = cvt.rn.f16x2(a, b); the fp32 value
// of hi is rebuilt with integer arithmetic (round-to-nearest-even at mantissa bit 13) instead of converting back, which is
// exact wherever hi is a normal fp16 number; below 2^-14 (2^-29 of the tensor's bound) it differs from the stored hi by at
// most 2^-25, i.e. 2^-40 of the bound -- the same floor lo's gradual underflow has.
__device__ __forceinline__ float f16_round_as_f32(float a) {
  const uint32_t b = __float_as_uint(a);
  return __uint_as_float((b + 0xfffu + ((b >> 13) & 1u)) & 0xffffe000u);
}
__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  lo = __floats2half2_rn(a - f16_round_as_f32(a), b - f16_round_as_f32(b));
}

// One 32-row x 32-column chunk of the epilogue, one warp, lane = accumulator row within the chunk.  Measured alternatives
// to the bulk tensor stores on a B200 (us for the 235->512 / 512->256 / 256->128 layers at M = 393 216): bulk stores
// 434 / 295 / 153, row-per-lane direct global stores 544 / 347 / 168, shared-memory transpose + coalesced global stores
// 472 / 320 / 170 -- the epilogue is not what bounds the kernel (see the A-tile L2 prefetch in the producer).
// Staging (4 KB per warp): OUT_F32  one 32 x 32 fp32 block, SWIZZLE_128B;
//                          OUT_PAIR two 32 x 32 fp16 blocks (hi at +0, lo at +2048), 64-byte rows, SWIZZLE_64B.
template <int EPI, int OUT>
__device__ __forceinline__ void f16_epilogue_chunk(const F16GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmOutLo,
                                                   uint8_t* stg, uint32_t taddr, int row0, int col0, int lane, float inv_ab,
                                                   float inv_aux, float s_out) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  float4 e[8];
  uint4 ah[4], al[4];
  if (EPI == EPI_BIAS_ACT) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      e[q] = (p.bias && col0 + 4 * q < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * q))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (p.aux_hi) {
    // coalesced: load i of lane l fetches 16-byte unit (l % 4) of row 8 i + l / 4 (each instruction reads eight 64-byte
    // row segments); out-of-range rows / columns read as zeros (act'(0) is finite, the product is clipped by the store)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + (lane >> 2), c = col0 + 8 * (lane & 3);
      const bool ok = row0 + rr < p.M && c < p.ldaux;
      const int64_t off = (int64_t)(row0 + rr) * p.ldaux + c;
      ah[i] = ok ? __ldg(reinterpret_cast<const uint4*>(p.aux_hi + off)) : make_uint4(0u, 0u, 0u, 0u);
      al[i] = ok ? __ldg(reinterpret_cast<const uint4*>(p.aux_lo + off)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tmem_ld_wait();
  // From here on v holds the output ALREADY MULTIPLIED by the output scale s_out (1 for fp32 outputs): the epilogue is
  // instruction-bound (ncu: ~1000 warp instructions per 32 x 32 chunk with two epilogue warps per scheduler, tensor pipe
  // 31 % on the 235->512 layer), so every multiply that can be folded is.
  float v[32];
  if (EPI == EPI_BIAS_ACT) {
    const float* ef = reinterpret_cast<const float*>(e);
    const float neg_s = -s_out;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float z = fmaf(__uint_as_float(r[j]), inv_ab, ef[j]);   // x W^T + b
      const float t = OUT == OUT_PAIR ? z * s_out : z;
      if (p.act == 1) {
        // ELU: z <= 0 -> exp(z) - 1 with ex2.approx (2-ulp exponential: < 2.4e-7 ABSOLUTE on a value of magnitude < 1,
        // i.e. fp32 rounding level of the O(1) sums it feeds; the 3xTF32 kernel's extra Taylor branch for |z| < 1/8
        // would cost 8 more instructions per element here)
        float ex;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * 1.4426950408889634f));
        v[j] = z > 0.f ? t : fmaf(ex, s_out, neg_s);
      } else if (p.act == 2) {
        v[j] = fmaxf(t, 0.f);
      } else {
        v[j] = t;
      }
    }
  } else {
    const float k = OUT == OUT_PAIR ? inv_ab * s_out : inv_ab;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * k;
  }
  if (lane == 0) tma_store_wait_read();  // the previous chunk's stores have finished reading the staging buffer
  __syncwarp();
  if (EPI == EPI_ACT_GRAD && p.aux_hi) {
    // transpose the coalesced aux fragments through the staging buffer (hi block at +0, lo block at +2048; 64-byte rows,
    // unit u of row r at u ^ ((r >> 1) & 3)), then every lane reads its own row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + (lane >> 2), u = (lane & 3) ^ ((rr >> 1) & 3);
      reinterpret_cast<uint4*>(stg + rr * 64)[u] = ah[i];
      reinterpret_cast<uint4*>(stg + 2048 + rr * 64)[u] = al[i];
    }
    __syncwarp();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 h = reinterpret_cast<const uint4*>(stg + lane * 64)[u ^ sw];
      const uint4 l = reinterpret_cast<const uint4*>(stg + 2048 + lane * 64)[u ^ sw];
      const __half2* h2 = reinterpret_cast<const __half2*>(&h);
      const __half2* l2 = reinterpret_cast<const __half2*>(&l);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fh = __half22float2(h2[k]), fl = __half22float2(l2[k]);
        v[8 * u + 2 * k + 0] *= act_grad_from_output((fh.x + fl.x) * inv_aux, p.act);
        v[8 * u + 2 * k + 1] *= act_grad_from_output((fh.y + fl.y) * inv_aux, p.act);
      }
    }
    __syncwarp();  // every lane has read its row before the staging buffer is reused for the output
  }
  if (OUT == OUT_F32) {
    float4* srow = reinterpret_cast<float4*>(stg + lane * 128);
    const int sw = lane & 7;  // SWIZZLE_128B: 16-byte unit q of row r lives at unit q ^ (r % 8)
#pragma unroll
    for (int q = 0; q < 8; ++q) srow[q ^ sw] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte unit u of 64-byte row r lives at unit u ^ ((r >> 1) % 4)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint4 h, l;
      __half2* h2 = reinterpret_cast<__half2*>(&h);
      __half2* l2 = reinterpret_cast<__half2*>(&l);
#pragma unroll
      for (int k = 0; k < 4; ++k) split2(v[8 * u + 2 * k], v[8 * u + 2 * k + 1], h2[k], l2[k]);
      reinterpret_cast<uint4*>(stg + lane * 64)[u ^ sw] = h;
      reinterpret_cast<uint4*>(stg + 2048 + lane * 64)[u ^ sw] = l;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmOut, stg, col0, row0);
    if (OUT == OUT_PAIR) tma_store_2d(tmOutLo, stg + 2048, col0, row0);
    tma_store_commit();
  }
  if (EPI == EPI_ACT_GRAD && p.colsum) {
    // bias gradient of the layer below = column sums of this output (rows >= M hold exact zeros: their accumulators are
    // products of zero-filled A rows); one owner per (CTA, row quarter, column) slot, so plain read-modify-write
    const float sum = warp_column_sums(v, lane) * (OUT == OUT_PAIR ? 1.f / s_out : 1.f);  // v carries the output scale
    if (col0 + lane < p.N) {
      float* slot = p.colsum + ((int64_t)blockIdx.x * 4 + ((threadIdx.x >> 5) & 3)) * p.N + col0 + lane;
      *slot += sum;
    }
  }
}

template <int BN, int EPI, int OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(f16_threads<BN>(), 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOutLo, const F16GemmParams p) {
  using Cfg = F16Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int HALF_B_BYTES = Cfg::B_BYTES / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  auto sAhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sBhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
  auto sBlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::EPI_BYTES);
  uint64_t* full = bars;                  // TMA bytes landed (own A pair + both halves of the B pair)
  uint64_t* empty = bars + STAGES;        // MMAs of BOTH CTAs reading the stage retired
  uint64_t* tfull = bars + 2 * STAGES;    // accumulator complete
  uint64_t* tempty = tfull + 2;           // accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_k_blocks = (p.K + FBK - 1) / FBK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  // scales: exact powers of two derived from device-resident bounds (f16x3_common.cuh)
  const float bound_a = __ldg(p.bound_a);
  const float inv_ab = 1.f / (f16x3_scale(bound_a) * f16x3_scale(__ldg(p.wstats + WSTAT_AMAX)));
  const float inv_aux = (EPI == EPI_ACT_GRAD && p.bound_aux) ? 1.f / f16x3_scale(__ldg(p.bound_aux)) : 1.f;
  float s_out = 1.f;
  if (OUT == OUT_PAIR) {
    float b;
    if (EPI == EPI_BIAS_ACT) {
      // |act(x W^T + b)| <= bound(x) max_n sum_k |W_nk| + max|b|; ELU(z) lies in (-1, z]
      b = bound_a * __ldg(p.wstats + WSTAT_ROW_L1) + __ldg(p.wstats + WSTAT_BIAS_MAX);
      if (p.act == 1) b = fmaxf(b, 1.f);
    } else {
      // |(dz W) act'| <= bound(dz) max_k sum_n |W_nk|  (here B = W^T, so that is its row L1 norm = W's column norm; act' <= 1)
      b = bound_a * __ldg(p.wstats + WSTAT_COL_L1);
    }
    b *= 1.001f;  // the L1 norms are rounded sums and the products above are rounded: stay an upper bound
    s_out = f16x3_scale(b);
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.bound_out) *p.bound_out = b;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAhi);
    tma_prefetch_desc(&tmAlo);
    tma_prefetch_desc(&tmBhi);
    tma_prefetch_desc(&tmBlo);
    tma_prefetch_desc(&tmOut);
    if (OUT == OUT_PAIR) tma_prefetch_desc(&tmOutLo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);  // one tcgen05.commit from each CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], Cfg::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers must be initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_m0 = [&](int item) { return ((item / p.num_n_tiles) * 2 + (int)rank) * BM; };
  auto item_n0 = [&](int item) { return (item % p.num_n_tiles) * BN; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // The A tiles stream from HBM (~1 us away) while the B tiles come from L2, and two 96 KB stages keep only ~32 KB of A
      // requests in flight per SM -- below the bandwidth-delay product of HBM (measured: 1.8 us per k-block against 0.83 us
      // of tensor time).  A second cursor therefore runs p.prefetch k-blocks AHEAD of the loads and pulls the A tiles into
      // L2 with bulk prefetches, so that the loads themselves see L2 latency.
      int pf_item = cluster_id, pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_item >= p.num_items) return;
        const int pm0 = item_m0(pf_item);
        tma_prefetch_2d(&tmAhi, pf_kb * FBK, pm0);
        tma_prefetch_2d(&tmAlo, pf_kb * FBK, pm0);
        if (++pf_kb == num_k_blocks) pf_kb = 0, pf_item += num_clusters;
      };
      for (int i = 0; i < p.prefetch; ++i) prefetch_next();
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const int m0 = item_m0(item), n0 = item_n0(item);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (p.prefetch > 0) prefetch_next();
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          tma_load_2d(sAhi(s), &tmAhi, kb * FBK, m0, &full[s]);
          tma_load_2d(sAlo(s), &tmAlo, kb * FBK, m0, &full[s]);
          tma_load_2d_mc(sBhi(s) + rank * HALF_B_BYTES, &tmBhi, kb * FBK, n0 + (int)rank * (BN / 2), &full[s], (uint16_t)3);
          tma_load_2d_mc(sBlo(s) + rank * HALF_B_BYTES, &tmBlo, kb * FBK, n0 + (int)rank * (BN / 2), &full[s], (uint16_t)3);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0, 0);
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
          const uint32_t ahi = smem_u32(sAhi(s)), alo = smem_u32(sAlo(s)), bhi = smem_u32(sBhi(s)), blo = smem_u32(sBlo(s));
          // K-major SWIZZLE_128B operands: 8-row groups are 1024 B apart (SBO); advancing K by 16 halves = +32 B
#pragma unroll
          for (int k = 0; k < FBK / F_UMMA_K; ++k) {
            const uint32_t off = (uint32_t)k * F_UMMA_K * 2;
            const uint64_t dah = make_smem_desc_sw128(ahi + off, 16, 1024, 2);
            const uint64_t dbh = make_smem_desc_sw128(bhi + off, 16, 1024, 2);
            mma_f16_ss(d_tmem, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            mma_f16_ss(d_tmem, dah, make_smem_desc_sw128(blo + off, 16, 1024, 2), idesc, 1u);
            mma_f16_ss(d_tmem, make_smem_desc_sw128(alo + off, 16, 1024, 2), dbh, idesc, 1u);
          }
          // the stage may be refilled (by either CTA's multicast) only when both CTAs are done reading it
          mma_commit_mc(&empty[s], (uint16_t)3);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
        mma_commit(&tfull[a]);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int ew = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    constexpr int COLS = BN / (Cfg::EPI_WARPS / 4);   // columns of the tile this warp owns
    const int part = (warp - 4) >> 2;
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
        for (int c0 = part * COLS; c0 < (part + 1) * COLS; c0 += 32)
          if (n0 + c0 < p.N)
            f16_epilogue_chunk<EPI, OUT>(p, &tmOut, &tmOutLo, stg, taddr + (uint32_t)c0, row0, n0 + c0, lane, inv_ab, inv_aux, s_out);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
    if (lane == 0) tma_store_wait_read();  // shared memory must outlive the last bulk store's read
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
// host
// ------------------------------------------------------------------------------------------------
int g_f16_prefetch = 0;  // measured on a B200: prefetching the A tiles into L2 never helps (0: 307 us, 4: 313, 16: 347 for 512->256)
int g_f16_bn = 0;        // 0 = choose by N (256 when N > 128), 128 / 256 = force (cusrl_b200_f16x3_set_tile)
;  // cusrl_b200_f16x3_set_prefetch; shared with gemm_wgrad_f16x3.cu

int colsum_finalize(const float* partial, int nblocks, int N, float* db, int accumulate, cudaStream_t s);  // gemm_wgrad_tf32.cu

template <int BN, int EPI, int OUT>
static int launch_f16_gemm(const CUtensorMap& tAh, const CUtensorMap& tAl, const CUtensorMap& tBh, const CUtensorMap& tBl,
                           const CUtensorMap& tO, const CUtensorMap& tOl, const F16GemmParams& p, cudaStream_t s) {
  using Cfg = F16Cfg<BN>;
  auto kern = gemm_f16x3_kernel<BN, EPI, OUT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gemm_f16x3: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  kern<<<gemm_grid_ctas(p.num_items), f16_threads<BN>(), Cfg::SMEM_BYTES, s>>>(tAh, tAl, tBh, tBl, tO, tOl, p);
  return check_launch("gemm_f16x3_kernel");
}

// Shared implementation of the forward and data-gradient entry points.  K = reduction length (valid columns of A and B),
// N = valid output columns; pair outputs are written over ld-padded rows (padding columns come out as exact zeros).
static int f16_gemm(const uint16_t* Ahi, const uint16_t* Alo, int64_t lda, const float* bound_a, const uint16_t* Bhi,
                    const uint16_t* Blo, int64_t ldb, const float* wstats, const float* bias, const uint16_t* aux_hi,
                    const uint16_t* aux_lo, int64_t ldaux, const float* bound_aux, float* out32, int64_t ldo32, uint16_t* out_hi,
                    uint16_t* out_lo, int64_t ldo16, float* bound_out, int64_t M, int64_t N, int64_t K, int act, int epi,
                    float* colsum_db, int colsum_accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  const bool pair = out_hi != nullptr;
  CUSRL_REQUIRE(Ahi && Alo && Bhi && Blo && bound_a && wstats && (out32 || (out_hi && out_lo)), CUSRL_B200_EINVAL,
                "linear_f16x3: null pointer");
  CUSRL_REQUIRE(!(out32 && pair), CUSRL_B200_EINVAL, "linear_f16x3: give the fp32 output or the pair output, not both");
  CUSRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N <= 65536 && K <= 65536, CUSRL_B200_EINVAL,
                "linear_f16x3: bad problem size");
  CUSRL_REQUIRE(act >= 0 && act <= 2, CUSRL_B200_EINVAL, "linear_f16x3: unknown activation code %d", act);
  CUSRL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= K && ldb >= K, CUSRL_B200_EALIGN,
                "linear_f16x3: operand leading dimensions must be multiples of 8 halves covering K");
  CUSRL_REQUIRE((N % 4) == 0, CUSRL_B200_EALIGN, "linear_f16x3: the number of output features must be a multiple of 4");
  CUSRL_REQUIRE(pair ? ((ldo16 % 8) == 0 && ldo16 >= N && bound_out) : ((ldo32 % 4) == 0 && ldo32 >= N), CUSRL_B200_EALIGN,
                "linear_f16x3: output leading dimension / bound");
  CUSRL_REQUIRE(!aux_hi || (aux_lo && bound_aux && (ldaux % 8) == 0 && ldaux >= N), CUSRL_B200_EINVAL,
                "linear_f16x3: aux pair needs lo, bound and a leading dimension that is a multiple of 8 covering N");
  CUSRL_REQUIRE(aligned_to(Ahi, 16) && aligned_to(Alo, 16) && aligned_to(Bhi, 16) && aligned_to(Blo, 16) &&
                    (!out32 || aligned_to(out32, 16)) && (!pair || (aligned_to(out_hi, 16) && aligned_to(out_lo, 16))) &&
                    (!bias || aligned_to(bias, 16)) && (!aux_hi || (aligned_to(aux_hi, 16) && aligned_to(aux_lo, 16))),
                CUSRL_B200_EALIGN, "linear_f16x3: pointers must be 16-byte aligned");
  const int bn = g_f16_bn ? g_f16_bn : (N > 128 ? 256 : 128);
  const uint64_t Kp = (uint64_t)((K + 7) / 8 * 8);
  CUtensorMap tAh, tAl, tBh, tBl, tO, tOl;
  if (int e = encode_tmap_2d_f16(&tAh, Ahi, Kp, (uint64_t)M, (uint64_t)lda, FBK, BM, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tAl, Alo, Kp, (uint64_t)M, (uint64_t)lda, FBK, BM, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tBh, Bhi, Kp, (uint64_t)N, (uint64_t)ldb, FBK, (uint32_t)bn / 2, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tBl, Blo, Kp, (uint64_t)N, (uint64_t)ldb, FBK, (uint32_t)bn / 2, TMAP_SW128)) return e;
  if (pair) {
    const uint64_t Np = (uint64_t)((N + 7) / 8 * 8);
    if (int e = encode_tmap_2d_f16(&tO, out_hi, Np, (uint64_t)M, (uint64_t)ldo16, 32, 32, TMAP_SW64)) return e;
    if (int e = encode_tmap_2d_f16(&tOl, out_lo, Np, (uint64_t)M, (uint64_t)ldo16, 32, 32, TMAP_SW64)) return e;
  } else {
    if (int e = encode_tmap_2d_f32(&tO, out32, (uint64_t)N, (uint64_t)M, (uint64_t)ldo32, 32, 32)) return e;
    tOl = tO;
  }
  F16GemmParams p{};
  p.bias = bias, p.aux_hi = (const __half*)aux_hi, p.aux_lo = (const __half*)aux_lo, p.ldaux = ldaux;
  p.bound_a = bound_a, p.wstats = wstats, p.bound_aux = bound_aux, p.bound_out = bound_out;
  p.M = (int)M, p.N = (int)N, p.K = (int)K, p.act = act;
  p.prefetch = g_f16_prefetch;
  p.num_m_tiles = (int)((M + BM - 1) / BM);
  p.num_n_tiles = (int)((N + bn - 1) / bn);
  p.num_items = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = gemm_grid_ctas(p.num_items);
  if (colsum_db) {
    const size_t need = (size_t)ctas * 4 * (size_t)N * sizeof(float);
    CUSRL_REQUIRE(workspace && workspace_bytes >= need && aligned_to(workspace, 16), CUSRL_B200_ESCRATCH,
                  "linear_dgrad_f16x3: workspace too small for the bias-gradient partials");
    cudaError_t me = cudaMemsetAsync(workspace, 0, need, s);
    CUSRL_REQUIRE(me == cudaSuccess, (int)me, "linear_dgrad_f16x3: cudaMemsetAsync: %s", cudaGetErrorString(me));
    p.colsum = (float*)workspace;
  }
  int rc = CUSRL_B200_EUNSUPPORTED;
#define CUSRL_F16_CASE(BN_, E_, O_) \
  if (bn == BN_ && epi == E_ && (int)pair == O_) rc = launch_f16_gemm<BN_, E_, O_>(tAh, tAl, tBh, tBl, tO, tOl, p, s);
  CUSRL_F16_CASE(256, EPI_BIAS_ACT, OUT_PAIR)
  CUSRL_F16_CASE(128, EPI_BIAS_ACT, OUT_PAIR)
  CUSRL_F16_CASE(256, EPI_BIAS_ACT, OUT_F32)
  CUSRL_F16_CASE(128, EPI_BIAS_ACT, OUT_F32)
  CUSRL_F16_CASE(256, EPI_ACT_GRAD, OUT_PAIR)
  CUSRL_F16_CASE(128, EPI_ACT_GRAD, OUT_PAIR)
  CUSRL_F16_CASE(256, EPI_ACT_GRAD, OUT_F32)
  CUSRL_F16_CASE(128, EPI_ACT_GRAD, OUT_F32)
#undef CUSRL_F16_CASE
  if (rc || !colsum_db) return rc;
  return colsum_finalize((const float*)workspace, ctas * 4, (int)N, colsum_db, colsum_accumulate, s);
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_f16x3_set_tile(int bn) {
  if (bn != 0 && bn != 128 && bn != 256) return CUSRL_B200_EINVAL;
  g_f16_bn = bn;
  return 0;
}

int cusrl_b200_f16x3_set_prefetch(int k_blocks) {
  if (k_blocks < 0 || k_blocks > 64) return CUSRL_B200_EINVAL;
  g_f16_prefetch = k_blocks;
  return 0;
}

int cusrl_b200_linear_fwd_f16x3(const uint16_t* Xhi, const uint16_t* Xlo, int64_t ldx, const float* x_bound, const uint16_t* Whi,
                                const uint16_t* Wlo, int64_t ldw, const float* w_stats, const float* bias, float* Y, int64_t ldy,
                                uint16_t* Yhi, uint16_t* Ylo, int64_t ldyh, float* y_bound, int64_t M, int64_t N, int64_t K,
                                int act, void* stream) {
  return f16_gemm(Xhi, Xlo, ldx, x_bound, Whi, Wlo, ldw, w_stats, bias, nullptr, nullptr, 0, nullptr, Y, ldy, Yhi, Ylo, ldyh,
                  y_bound, M, N, K, act, EPI_BIAS_ACT, nullptr, 0, nullptr, 0, stream);
}

int cusrl_b200_linear_dgrad_f16x3(const uint16_t* dYhi, const uint16_t* dYlo, int64_t lddy, const float* dy_bound,
                                  const uint16_t* WThi, const uint16_t* WTlo, int64_t ldwt, const float* w_stats,
                                  const uint16_t* Xact_hi, const uint16_t* Xact_lo, int64_t ldxa, const float* xact_bound,
                                  float* dX, int64_t lddx32, uint16_t* dXhi, uint16_t* dXlo, int64_t lddx, float* dx_bound,
                                  int64_t M, int64_t N, int64_t K, int act, float* db_below, int accumulate, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  // dX[M,K] = dY[M,N] @ W[N,K]: as a K-major GEMM the reduction runs over N and B is the transposed pair WT[K,N]
  return f16_gemm(dYhi, dYlo, lddy, dy_bound, WThi, WTlo, ldwt, w_stats, nullptr, Xact_hi, Xact_lo, ldxa, xact_bound, dX, lddx32,
                  dXhi, dXlo, lddx, dx_bound, M, /*N=*/K, /*K=*/N, act, EPI_ACT_GRAD, db_below, accumulate, workspace,
                  workspace_bytes, stream);
}

}  // extern "C"
