// K7, sequence-resident form: ONE launch runs all T steps of an LSTM layer (reference: nn.LSTM inside
// cusrl/nn/module/rnn.py:62-97,264-299, driven over episode-segmented sequences by cusrl/nn/utils/recurrent.py:160-272).
//
// The per-step form (lstm_kernels.cu + one K6 GEMM per step) spends its time in launch-sized work: at the recurrent-PPO
// minibatch (Nb = 1024 columns, H = 256) a step is a [1024 x 256] x [256 x 1024] product -- 32 CTAs of the generic GEMM,
// 16 us -- plus a 5 us cell kernel, 48 launches per layer, ~9 000 per PPO iteration.  Here the recurrence stays on the SMs:
//
//   work split   a ROW TILE of 128 batch columns is advanced through time by H/16 CTAs, each owning 16 hidden units, i.e.
//                the 64 gate columns (i, f, g, o of its units) of the recurrent product;  8 row tiles x 16 slices = 128 CTAs
//                at the minibatch shape, all co-resident (one CTA per SM, grid <= #SMs), row-tile groups loop over more
//                tiles when Nb is larger (rollout / statistics passes).
//   operands     W_hh slice [64 x H] as an fp16 hi/lo pair (f16x3_common.cuh) RESIDENT in shared memory for the whole
//                launch (64 KB at H = 256); h_{t-1} of the row tile [128 x H] as a pair arrives by TMA every step (128 KB)
//                from a double-buffered exchange array in L2 that the slices' epilogues write.
//   step         3 x (H/16) tcgen05 kind::f16 MMAs (M = 128, N = 64) into one TMEM accumulator -> epilogue warps:
//                tcgen05.ld, + input projection (+ b_hh), gate nonlinearities, c_t (carried in REGISTERS across steps),
//                h_t, in-line reset where done[t] (the reference's split / pad / scatter), all saved tensors of the
//                backward pass, and h_t re-split into the exchange pair.
//   hand-over    per (row tile, step) one counter in global memory: every slice's epilogue publishes its part of h_t
//                (__threadfence, then one relaxed add), the TMA producers of the tile's CTAs acquire-poll it, order the
//                generic-proxy writes before their async-proxy reads (fence.proxy.async) and load.  No cluster, no
//                cooperative launch, no host involvement.
//
// Numerics: fp32-equivalent recurrent product (three fp16 MMAs, fp32 accumulation), |h| < 1 fixes the scale of the h pair
// (2^14), W_hh's scale comes from its exact amax (weight_prep_f16); everything else is the fp32 arithmetic of
// lstm_cell_fwd_kernel.  Shapes: H a multiple of 64, H <= 256 (the operands must fit in shared memory); others take the
// per-step path.
#include "f16x3_common.cuh"
#include "gemm_common.cuh"

namespace cusrl_b200 {

constexpr int LS_HS = 16;                 // hidden units per CTA
constexpr int LS_N = 4 * LS_HS;           // gate columns per CTA = UMMA N
constexpr int LS_KB = 64;                 // halves per k-block (one 128-byte swizzle span)
constexpr int LS_UMMA_K = 16;             // halves per tcgen05.mma
constexpr int LS_MAX_KB = 4;              // H <= 256
constexpr int LS_EPI_WARPS = 8;
constexpr int LS_THREADS = 128 + 32 * LS_EPI_WARPS;
constexpr int LS_W_KB_BYTES = LS_N * LS_KB * 2;    // 8 KB per half per k-block
constexpr int LS_A_KB_BYTES = BM * LS_KB * 2;      // 16 KB per half per k-block

struct LstmSeqFwdParams {
  const float* xp;        // [T * Nb, 4H] input projection of every step (b_ih included), row pitch ldxp
  int64_t ldxp;
  const float* b_hh;      // [4H] or null
  const float* h0;        // [Nb, H] state entering step 0 (null: zeros)
  const float* c0;
  const uint8_t* done;    // [T, Nb] or null
  float* gates;           // [T, Nb, 4H] activated gates
  float* cseq;            // [T, Nb, H]  c_t
  float* out;             // [T, Nb, H]  h_t
  float* hin;             // [T, Nb, H]  state entering step t: h0, then h_{t-1} (1 - done_{t-1})   (nullable together with cin)
  float* cin;
  __half* hx_hi;          // [2, Nbp, H] exchange pair (Nbp = tiles * 128), parity t & 1 holds the state entering step t
  __half* hx_lo;
  const float* wstats;    // weight_prep_f16 statistics of W_hh
  unsigned int* flags;    // [tiles, T + 1], zeroed by the launcher
  int T, Nb, H, tiles, slices, groups, nbp;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy writes (made visible to this thread by the acquire above) -> ordered before this thread's async-proxy reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * LS_EPI_WARPS) : "memory"); }

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void store_pair8(__half* hi, __half* lo, const float (&v)[8], float s) {
  uint4 h, l;
  __half2* h2 = reinterpret_cast<__half2*>(&h);
  __half2* l2 = reinterpret_cast<__half2*>(&l);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = v[2 * k] * s, b = v[2 * k + 1] * s;
    h2[k] = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2[k]);
    l2[k] = __floats2half2_rn(a - back.x, b - back.y);
  }
  *reinterpret_cast<uint4*>(hi) = h;
  *reinterpret_cast<uint4*>(lo) = l;
}

__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                    const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                    const LstmSeqFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = p.H / LS_KB;
  uint8_t* sW = smem;                                        // [nkb][hi | lo][64 rows x 128 B]
  uint8_t* sA = smem + LS_MAX_KB * 2 * LS_W_KB_BYTES;        // [nkb][hi | lo][128 rows x 128 B]
  uint8_t* tail = sA + LS_MAX_KB * 2 * LS_A_KB_BYTES;
  float* s_bias = reinterpret_cast<float*>(tail);            // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* tfull = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.slices, group = blockIdx.x / p.slices;
  const int T = p.T, H = p.H;

  if (threadIdx.x < LS_N) {
    const int q = threadIdx.x / LS_HS, jj = threadIdx.x % LS_HS;
    s_bias[threadIdx.x] = p.b_hh ? __ldg(p.b_hh + q * H + slice * LS_HS + jj) : 0.f;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
    tma_prefetch_desc(&tmAhi);
    tma_prefetch_desc(&tmAlo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(nkb * 2 * LS_W_KB_BYTES));
      for (int kb = 0; kb < nkb; ++kb)
        for (int q = 0; q < 4; ++q) {
          // 16 rows of gate q: two 8-row swizzle atoms, placed where a 64-row box would have put them
          tma_load_2d(sW + (kb * 2 + 0) * LS_W_KB_BYTES + q * LS_HS * 128, &tmWhi, kb * LS_KB, q * H + slice * LS_HS, w_full);
          tma_load_2d(sW + (kb * 2 + 1) * LS_W_KB_BYTES + q * LS_HS * 128, &tmWlo, kb * LS_KB, q * H + slice * LS_HS, w_full);
        }
      uint32_t it = 0;
      for (int tile = group; tile < p.tiles; tile += p.groups) {
        const unsigned int* flag = p.flags + (int64_t)tile * (T + 1);
        for (int t = 0; t < T; ++t, ++it) {
          // every slice of this row tile has published the state entering step t -- which also means that every CTA of the
          // tile, this one included, is done with step t - 1 (accumulator drained, operand tile no longer read)
          while (ld_acquire_u32(flag + t) < (unsigned int)p.slices) __nanosleep(32);
          fence_proxy_async_all();
          mbar_expect_tx(a_full, (uint32_t)(nkb * 2 * LS_A_KB_BYTES));
          const int row = (t & 1) * p.nbp + tile * BM;
          for (int kb = 0; kb < nkb; ++kb) {
            tma_load_2d(sA + (kb * 2 + 0) * LS_A_KB_BYTES, &tmAhi, kb * LS_KB, row, a_full);
            tma_load_2d(sA + (kb * 2 + 1) * LS_A_KB_BYTES, &tmAlo, kb * LS_KB, row, a_full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, LS_N, 0, 0);
      mbar_wait(w_full, 0);
      uint32_t it = 0;
      for (int tile = group; tile < p.tiles; tile += p.groups)
        for (int t = 0; t < T; ++t, ++it) {
          mbar_wait(a_full, it & 1);
          tc_fence_after();
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t ahi = smem_u32(sA + (kb * 2 + 0) * LS_A_KB_BYTES), alo = smem_u32(sA + (kb * 2 + 1) * LS_A_KB_BYTES);
            const uint32_t bhi = smem_u32(sW + (kb * 2 + 0) * LS_W_KB_BYTES), blo = smem_u32(sW + (kb * 2 + 1) * LS_W_KB_BYTES);
#pragma unroll
            for (int k = 0; k < LS_KB / LS_UMMA_K; ++k) {
              const uint32_t off = (uint32_t)k * LS_UMMA_K * 2;
              const uint64_t dah = make_smem_desc_sw128(ahi + off, 16, 1024, 2);
              const uint64_t dbh = make_smem_desc_sw128(bhi + off, 16, 1024, 2);
              mma_f16_ss(tmem_base, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              mma_f16_ss(tmem_base, dah, make_smem_desc_sw128(blo + off, 16, 1024, 2), idesc, 1u);
              mma_f16_ss(tmem_base, make_smem_desc_sw128(alo + off, 16, 1024, 2), dbh, idesc, 1u);
            }
          }
          mma_commit(tfull);
        }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int ew = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;        // which 8 of the CTA's 16 hidden units
    const int u0 = slice * LS_HS + half * 8; // first hidden unit of this thread
    const float s_h = f16x3_scale(1.f);
    const float inv_ab = 1.f / (s_h * f16x3_scale(__ldg(p.wstats + WSTAT_AMAX)));
    float bias[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[q][j] = s_bias[q * LS_HS + half * 8 + j];
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(half * 8);
    uint32_t it = 0;
    for (int tile = group; tile < p.tiles; tile += p.groups) {
      unsigned int* flag = p.flags + (int64_t)tile * (T + 1);
      const int row = tile * BM + ew * 32 + lane;
      const bool valid = row < p.Nb;
      float c[8], h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = 0.f, h[j] = 0.f;
      if (valid) {
        const int64_t o = (int64_t)row * H + u0;
        if (p.c0) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p.c0 + o)), b = __ldg(reinterpret_cast<const float4*>(p.c0 + o + 4));
          c[0] = a.x, c[1] = a.y, c[2] = a.z, c[3] = a.w, c[4] = b.x, c[5] = b.y, c[6] = b.z, c[7] = b.w;
        }
        if (p.h0) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p.h0 + o)), b = __ldg(reinterpret_cast<const float4*>(p.h0 + o + 4));
          h[0] = a.x, h[1] = a.y, h[2] = a.z, h[3] = a.w, h[4] = b.x, h[5] = b.y, h[6] = b.z, h[7] = b.w;
        }
        if (p.hin) {
          *reinterpret_cast<float4*>(p.hin + o) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(p.hin + o + 4) = make_float4(h[4], h[5], h[6], h[7]);
          *reinterpret_cast<float4*>(p.cin + o) = make_float4(c[0], c[1], c[2], c[3]);
          *reinterpret_cast<float4*>(p.cin + o + 4) = make_float4(c[4], c[5], c[6], c[7]);
        }
        const int64_t ox = (int64_t)row * H + u0;  // parity 0
        store_pair8(p.hx_hi + ox, p.hx_lo + ox, h, s_h);
      }
      __threadfence();
      epi_bar_sync();
      if (threadIdx.x == 128) atomicAdd(flag, 1u);

      for (int t = 0; t < T; ++t, ++it) {
        // operands that do not depend on the recurrence are requested before waiting for the accumulator
        float4 x[4][2];
        uint8_t dn = 0;
        if (valid) {
          const float* xr = p.xp + ((int64_t)t * p.Nb + row) * p.ldxp + u0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            x[q][0] = __ldg(reinterpret_cast<const float4*>(xr + q * H));
            x[q][1] = __ldg(reinterpret_cast<const float4*>(xr + q * H + 4));
          }
          if (p.done) dn = p.done[(int64_t)t * p.Nb + row];
        }
        mbar_wait(tfull, it & 1);
        tc_fence_after();
        uint32_t r[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x8(taddr + (uint32_t)(q * LS_HS), r[q]);
        tmem_ld_wait();
        tc_fence_before();
        if (valid) {
          float g[4][8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xv[8] = {x[q][0].x, x[q][0].y, x[q][0].z, x[q][0].w, x[q][1].x, x[q][1].y, x[q][1].z, x[q][1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float pre = xv[j] + fmaf(__uint_as_float(r[q][j]), inv_ab, bias[q][j]);
              g[q][j] = q == 2 ? tanhf(pre) : sigmoid_acc(pre);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            c[j] = g[1][j] * c[j] + g[0][j] * g[2][j];
            h[j] = g[3][j] * tanhf(c[j]);
          }
          const int64_t o = ((int64_t)t * p.Nb + row) * H + u0;
          float* gr = p.gates + ((int64_t)t * p.Nb + row) * 4 * H + u0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<float4*>(gr + q * H) = make_float4(g[q][0], g[q][1], g[q][2], g[q][3]);
            *reinterpret_cast<float4*>(gr + q * H + 4) = make_float4(g[q][4], g[q][5], g[q][6], g[q][7]);
          }
          *reinterpret_cast<float4*>(p.cseq + o) = make_float4(c[0], c[1], c[2], c[3]);
          *reinterpret_cast<float4*>(p.cseq + o + 4) = make_float4(c[4], c[5], c[6], c[7]);
          *reinterpret_cast<float4*>(p.out + o) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(p.out + o + 4) = make_float4(h[4], h[5], h[6], h[7]);
          if (t + 1 < T) {
            const float m = dn ? 0.f : 1.f;   // the state handed to step t + 1 restarts where the episode ended
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] *= m, h[j] *= m;
            if (p.hin) {
              const int64_t on = o + (int64_t)p.Nb * H;
              *reinterpret_cast<float4*>(p.hin + on) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(p.hin + on + 4) = make_float4(h[4], h[5], h[6], h[7]);
              *reinterpret_cast<float4*>(p.cin + on) = make_float4(c[0], c[1], c[2], c[3]);
              *reinterpret_cast<float4*>(p.cin + on + 4) = make_float4(c[4], c[5], c[6], c[7]);
            }
            const int64_t ox = ((int64_t)((t + 1) & 1) * p.nbp + row) * H + u0;
            store_pair8(p.hx_hi + ox, p.hx_lo + ox, h, s_h);
          }
        }
        // publish: this slice's part of the state entering step t + 1 is in L2, and this CTA has drained its accumulator
        __threadfence();
        epi_bar_sync();
        if (threadIdx.x == 128) atomicAdd(flag + t + 1, 1u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

constexpr int LS_SMEM_BYTES = LS_MAX_KB * 2 * (LS_W_KB_BYTES + LS_A_KB_BYTES) + 512 + 1024;

static bool lstm_seq_shape_ok(int64_t H) { return H > 0 && (H % LS_KB) == 0 && H <= LS_MAX_KB * LS_KB; }

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_lstm_seq_supported(int64_t H) { return lstm_seq_shape_ok(H) ? 1 : 0; }

size_t cusrl_b200_lstm_seq_workspace_bytes(int64_t T, int64_t Nb, int64_t H) {
  if (T <= 0 || Nb <= 0 || !lstm_seq_shape_ok(H)) return 0;
  const int64_t tiles = (Nb + BM - 1) / BM;
  const size_t flags = (size_t)((tiles * (T + 1) * 4 + 255) / 256 * 256);
  return flags + (size_t)(2 * 2 * tiles * BM * H) * sizeof(uint16_t);
}

int cusrl_b200_lstm_seq_fwd_f32(const float* xp, int64_t ldxp, const uint16_t* Whi, const uint16_t* Wlo, int64_t ldw,
                                const float* w_stats, const float* b_hh, const float* h0, const float* c0, const uint8_t* done,
                                float* gates, float* cseq, float* out, float* hin, float* cin, int64_t T, int64_t Nb, int64_t H,
                                void* workspace, size_t workspace_bytes, void* stream) {
  CUSRL_REQUIRE(xp && Whi && Wlo && w_stats && gates && cseq && out && workspace, CUSRL_B200_EINVAL, "lstm_seq_fwd: null pointer");
  CUSRL_REQUIRE((hin == nullptr) == (cin == nullptr), CUSRL_B200_EINVAL, "lstm_seq_fwd: hin / cin go together");
  CUSRL_REQUIRE(T > 0 && Nb > 0 && T < (1 << 20) && Nb < (1ll << 30), CUSRL_B200_EINVAL, "lstm_seq_fwd: bad sizes");
  CUSRL_REQUIRE(lstm_seq_shape_ok(H), CUSRL_B200_EUNSUPPORTED, "lstm_seq_fwd: H must be a multiple of 64, at most 256 (got %lld)",
                (long long)H);
  CUSRL_REQUIRE((ldxp % 4) == 0 && ldxp >= 4 * H && ldw >= H && (ldw % 8) == 0, CUSRL_B200_EALIGN, "lstm_seq_fwd: leading dimensions");
  CUSRL_REQUIRE(aligned_to(xp, 16) && aligned_to(Whi, 16) && aligned_to(Wlo, 16) && aligned_to(gates, 16) && aligned_to(cseq, 16) &&
                    aligned_to(out, 16) && (!hin || (aligned_to(hin, 16) && aligned_to(cin, 16))) && (!h0 || aligned_to(h0, 16)) &&
                    (!c0 || aligned_to(c0, 16)) && aligned_to(workspace, 256),
                CUSRL_B200_EALIGN, "lstm_seq_fwd: pointers must be 16-byte aligned (workspace: 256)");
  const size_t need = cusrl_b200_lstm_seq_workspace_bytes(T, Nb, H);
  CUSRL_REQUIRE(workspace_bytes >= need, CUSRL_B200_ESCRATCH, "lstm_seq_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  LstmSeqFwdParams p{};
  p.tiles = (int)((Nb + BM - 1) / BM);
  p.slices = (int)(H / LS_HS);
  const int max_groups = sm_count() / p.slices;
  CUSRL_REQUIRE(max_groups >= 1, CUSRL_B200_EUNSUPPORTED, "lstm_seq_fwd: not enough SMs for one row tile");
  p.groups = p.tiles < max_groups ? p.tiles : max_groups;
  p.nbp = p.tiles * BM;
  const size_t flag_bytes = (size_t)(((int64_t)p.tiles * (T + 1) * 4 + 255) / 256 * 256);
  p.flags = (unsigned int*)workspace;
  p.hx_hi = (__half*)((uint8_t*)workspace + flag_bytes);
  p.hx_lo = p.hx_hi + (size_t)2 * p.nbp * H;
  // flags AND exchange rows are cleared: rows >= Nb of the last tile are read by TMA (their results are discarded)
  cudaError_t me = cudaMemsetAsync(workspace, 0, need, s);
  CUSRL_REQUIRE(me == cudaSuccess, (int)me, "lstm_seq_fwd: cudaMemsetAsync: %s", cudaGetErrorString(me));
  p.xp = xp, p.ldxp = ldxp, p.b_hh = b_hh, p.h0 = h0, p.c0 = c0, p.done = done;
  p.gates = gates, p.cseq = cseq, p.out = out, p.hin = hin, p.cin = cin, p.wstats = w_stats;
  p.T = (int)T, p.Nb = (int)Nb, p.H = (int)H;
  CUtensorMap tWh, tWl, tAh, tAl;
  if (int e = encode_tmap_2d_f16(&tWh, Whi, (uint64_t)H, (uint64_t)(4 * H), (uint64_t)ldw, LS_KB, LS_HS, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tWl, Wlo, (uint64_t)H, (uint64_t)(4 * H), (uint64_t)ldw, LS_KB, LS_HS, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tAh, p.hx_hi, (uint64_t)H, (uint64_t)(2 * p.nbp), (uint64_t)H, LS_KB, BM, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tAl, p.hx_lo, (uint64_t)H, (uint64_t)(2 * p.nbp), (uint64_t)H, LS_KB, BM, TMAP_SW128)) return e;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM_BYTES);
    CUSRL_REQUIRE(e == cudaSuccess, (int)e, "lstm_seq_fwd: cudaFuncSetAttribute(%d bytes): %s", LS_SMEM_BYTES, cudaGetErrorString(e));
    configured = true;
  }
  lstm_seq_fwd_kernel<<<p.groups * p.slices, LS_THREADS, LS_SMEM_BYTES, s>>>(tWh, tWl, tAh, tAl, p);
  return check_launch("lstm_seq_fwd_kernel");
}

}  // extern "C"
