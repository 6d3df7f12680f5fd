// K7, sequence-resident form: ONE launch runs all T steps of an LSTM layer (reference: nn.LSTM inside
// cusrl/nn/module/rnn.py:62-97,264-299, driven over episode-segmented sequences by cusrl/nn/utils/recurrent.py:160-272).
//
// The per-step form (lstm_kernels.cu + one K6 GEMM per step) spends its time in launch-sized work: at the recurrent-PPO
// minibatch (Nb = 1024 columns, H = 256) a step is a [1024 x 256] x [256 x 1024] product -- 32 CTAs of the generic GEMM,
// 16 us -- plus a 5 us cell kernel, 48 launches per layer, ~9 000 per PPO iteration.  Here the recurrence stays on the SMs:
//
//   work split   a ROW TILE of 128 batch columns is advanced through time by H/16 CTAs, each owning 16 hidden units, i.e.
//                the 64 gate columns (i, f, g, o of its units) of the recurrent product;  8 row tiles x 16 slices = 128 CTAs
//                at the minibatch shape, all co-resident (one CTA per SM; the grid is bounded by the number of 4-CTA clusters
//                the device holds at once), row-tile groups loop over more tiles when Nb is larger (rollout / statistics).
//   operands     W_hh slice [64 x H] as an fp16 hi/lo pair (f16x3_common.cuh) RESIDENT in shared memory for the whole
//                launch (64 KB at H = 256); h_{t-1} of the row tile [128 x H] as a pair arrives by TMA every step (128 KB)
//                from a double-buffered exchange array in L2 that the slices' epilogues write -- read from L2 once per
//                4-CTA cluster and multicast; the step's slice of the input projection arrives by TMA too.
//   step         3 x (H/16) tcgen05 kind::f16 MMAs (M = 128, N = 64) into one TMEM accumulator -> 16 epilogue warps:
//                tcgen05.ld, + input projection (+ b_hh), gate nonlinearities, c_t (carried in REGISTERS across steps),
//                h_t, in-line reset where done[t] (the reference's split / pad / scatter), the tensors saved for the
//                backward pass, and h_t re-split into the exchange pair.
//   hand-over    per (row tile, step) one counter in global memory: every slice's epilogue stores its part of h_t, meets at
//                a named barrier and one thread adds to the counter with release semantics (the barrier makes the release
//                cumulative over the CTA's writes); the TMA producers of the tile's CTAs acquire-poll the counter, order the
//                generic-proxy writes before their async-proxy reads (fence.proxy.async) and load.  No cooperative launch,
//                no host involvement.
//   data layout  TMEM lanes are accumulator ROWS, so with the natural thread mapping a warp's global access touches 32 rows
//                = 32 L1 wavefronts per instruction; that, not latency or bandwidth, bounded the first versions (7 000
//                wavefront cycles per forward step, 27 600 per backward step).  Tensors only the backward twin reads (gates,
//                c_t, c_in, and the backward's partial products) therefore use a private tiled layout
//                [T][tile][H/4][(gate)][128 rows][4] in which a warp's rows are contiguous, and row-major outputs (h_t, h_in,
//                the exchange pair, dgates) are transposed through the operand tile while it is idle.
//                Inference calls (gates == null) save nothing and read / write a layer's slice of a flat memory in place.
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
constexpr int LS_EPI_WARPS = 16;
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
  float* hin;             // [T, Nb, H]  state entering step t: h0, then h_{t-1} (1 - done_{t-1})   (nullable)
  float* cin;             // PRIVATE layout, like gates and cseq: [T][tiles][H/4][(4 gates)][128 rows][4] -- only the backward kernel
                          // reads them, and in this layout the 32 rows of a warp are contiguous
  float* c_last;          // [Nb, .] row-major c_{T-1} and h_{T-1}, row pitch ld_last (nullable): the next memory
  float* h_last;
  int64_t ld0, ld_last;   // row pitches of h0 / c0 and of h_last / c_last (a layer's slice of a flat [N, layers * H] memory)
  __half* hx_hi;          // [2, Nbp, H] exchange pair (Nbp = tiles * 128), parity t & 1 holds the state entering step t
  __half* hx_lo;
  const float* wstats;    // weight_prep_f16 statistics of W_hh
  unsigned int* flags;    // [tiles, T + 1], zeroed by the launcher
  int T, Nb, H, tiles, slices, groups, nbp;
  int debug;
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
// the CTA's contribution to a hand-over counter: the epilogue barrier orders every epilogue thread's global writes before
// this one thread's release (cumulative), as in CUTLASS's semaphore release
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
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

__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void store_pair4(__half* hi, __half* lo, const float (&v)[4], float s) {
  uint2 h, l;
  __half2* h2 = reinterpret_cast<__half2*>(&h);
  __half2* l2 = reinterpret_cast<__half2*>(&l);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float a = v[2 * k] * s, b = v[2 * k + 1] * s;
    h2[k] = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2[k]);
    l2[k] = __floats2half2_rn(a - back.x, b - back.y);
  }
  *reinterpret_cast<uint2*>(hi) = h;
  *reinterpret_cast<uint2*>(lo) = l;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
}

// Clusters of LS_CLUSTER CTAs = consecutive slices of ONE row tile: the state tile every slice needs (128 KB, the same for
// all of them) is read from L2 once per cluster -- CTA r of the cluster loads k-block r (and r + 4, ...) and multicasts it to
// all four.  Without this the 16 slices of a tile pulled 16 x 128 KB = 2 MB per tile and step over the L2 fabric, 16 MB per
// step at the minibatch shape, which alone cost ~3 us of an 11 us step (measured).
constexpr int LS_CLUSTER = 4;

__global__ void __cluster_dims__(LS_CLUSTER, 1, 1) __launch_bounds__(LS_THREADS, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                    const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                    const __grid_constant__ CUtensorMap tmX, const LstmSeqFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = p.H / LS_KB;
  uint8_t* sW = smem;                                        // [nkb][hi | lo][64 rows x 128 B]
  uint8_t* sA = smem + LS_MAX_KB * 2 * LS_W_KB_BYTES;        // [nkb][hi | lo][128 rows x 128 B]
  uint8_t* sX = sA + LS_MAX_KB * 2 * LS_A_KB_BYTES;          // [4 gates][128 rows x 64 B] input projection of one step
  uint8_t* tail = sX + 4 * BM * LS_HS * 4;
  float* s_bias = reinterpret_cast<float*>(tail);            // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256);
  uint64_t* w_full = bars;
  uint64_t* tfull = bars + 1;
  uint64_t* x_full = bars + 2;
  uint64_t* a_full = bars + 3;             // one per k-block: the MMAs of a k-block start when ITS 32 KB have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + LS_MAX_KB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.slices, group = blockIdx.x / p.slices;
  const uint32_t crank = cluster_ctarank();
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
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(tfull, 1);
    mbar_init(x_full, 1);
    for (int kb = 0; kb < LS_MAX_KB; ++kb) mbar_init(&a_full[kb], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peers' barriers must be initialised before anything is multicast into them
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
      for (int tile = group; tile < p.tiles; tile += p.groups) {
        const unsigned int* flag = p.flags + (int64_t)tile * (T + 1);
        for (int t = 0; t < T; ++t) {
          // every slice of this row tile has published the state entering step t -- which also means that every CTA of the
          // tile, the four of this cluster included, is done with step t - 1 (accumulator drained, operand tile no longer
          // read, the k-block barriers in their next phase: a peer's multicast may land before this CTA arms them)
          while (ld_acquire_u32(flag + t) < (unsigned int)p.slices) {
          }
          fence_proxy_async_all();
          for (int kb = 0; kb < nkb; ++kb) mbar_expect_tx(&a_full[kb], (uint32_t)(2 * LS_A_KB_BYTES));
          const int row = (t & 1) * p.nbp + tile * BM;
          for (int kb = (int)crank; kb < nkb; kb += LS_CLUSTER) {
            tma_load_2d_mc(sA + (kb * 2 + 0) * LS_A_KB_BYTES, &tmAhi, kb * LS_KB, row, &a_full[kb], (uint16_t)((1u << LS_CLUSTER) - 1));
            tma_load_2d_mc(sA + (kb * 2 + 1) * LS_A_KB_BYTES, &tmAlo, kb * LS_KB, row, &a_full[kb], (uint16_t)((1u << LS_CLUSTER) - 1));
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
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait_cluster(&a_full[kb], it & 1);
            tc_fence_after();
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
    // Thread = (accumulator row, quad of 4 hidden units): TMEM lanes are rows, so a warp's global accesses with this mapping
    // touch 32 different rows -- 32 L1 wavefronts per instruction, which made the loads / stores of a step cost more LSU time
    // than everything else together (measured: 7 000 wavefront cycles per step).  Hence: the input projection arrives by TMA
    // (x_full), tensors only this kernel's backward twin reads (gates, c_t, c_in) use a layout in which a warp's 32 rows are
    // contiguous, and row-major outputs (h_t, the state entering step t + 1 and its fp16 pair) are transposed through the
    // idle operand tile in shared memory and written 16 bytes per thread along rows.
    const int ew = warp & 3;                 // TMEM lane quarter this warp may access
    const int part = (warp - 4) >> 2;        // which 4 of the CTA's 16 hidden units
    const int quad = slice * (LS_HS / 4) + part;
    const int rloc = ew * 32 + lane;
    const int etid = threadIdx.x - 128;      // 0..511
    const float s_h = f16x3_scale(1.f);
    const float inv_ab = 1.f / (s_h * f16x3_scale(__ldg(p.wstats + WSTAT_AMAX)));
    float bias[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) bias[q][j] = s_bias[q * LS_HS + part * 4 + j];
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(part * 4);
    const bool leader = etid == 0;
    const int nquads = H / 4;
    // staging inside the operand tile (idle between the MMAs of a step and the next step's loads)
    uint8_t* st_hx = sA;                      // [hi | lo][128 rows][16 halves]   8 KB
    float* st_out = reinterpret_cast<float*>(sA + 8192);    // [128 rows][16]      8 KB
    float* st_hin = reinterpret_cast<float*>(sA + 16384);   // [128 rows][16]      8 KB
    uint32_t it = 0;

    // row-major copy-out of the staged tiles: 16 bytes per thread, a warp covers 8 (fp32) or 16 (fp16) whole row segments
    auto flush = [&](int tile, int t_out, int par, bool with_out, bool with_hin) {
      {  // exchange pair -> parity `par`
        const int hl = etid >> 8, j = etid & 255, r = j >> 1, ch = j & 1;
        const uint4 v = *reinterpret_cast<const uint4*>(st_hx + hl * 4096 + r * 32 + ch * 16);
        __half* dst = (hl ? p.hx_lo : p.hx_hi) + ((int64_t)par * p.nbp + tile * BM + r) * H + slice * LS_HS + ch * 8;
        *reinterpret_cast<uint4*>(dst) = v;
      }
      const int r = etid >> 2, ch = etid & 3;
      const int grow = tile * BM + r;
      if (grow < p.Nb) {
        if (with_out)
          *reinterpret_cast<float4*>(p.out + ((int64_t)t_out * p.Nb + grow) * H + slice * LS_HS + ch * 4) =
              *reinterpret_cast<const float4*>(st_out + r * 16 + ch * 4);
        if (with_hin && p.hin)
          *reinterpret_cast<float4*>(p.hin + ((int64_t)(t_out + 1) * p.Nb + grow) * H + slice * LS_HS + ch * 4) =
              *reinterpret_cast<const float4*>(st_hin + r * 16 + ch * 4);
      }
    };
    auto load_x = [&](int tile, int t) {   // leader only: this CTA's 4 x [128 rows x 16] slices of the input projection
      mbar_expect_tx(x_full, 4 * BM * LS_HS * 4);
      for (int q = 0; q < 4; ++q) tma_load_2d(sX + q * (BM * LS_HS * 4), &tmX, q * H + slice * LS_HS, t * p.Nb + tile * BM, x_full);
    };

    for (int tile = group; tile < p.tiles; tile += p.groups) {
      unsigned int* flag = p.flags + (int64_t)tile * (T + 1);
      const int row = tile * BM + rloc;
      const bool valid = row < p.Nb;
      if (leader && tile == group) load_x(tile, 0);   // later tiles: requested at the end of the previous tile
      float c[4] = {0.f, 0.f, 0.f, 0.f}, h[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const int64_t o = (int64_t)row * p.ld0 + slice * LS_HS + part * 4;
        if (p.c0) ld4(p.c0 + o, c);
        if (p.h0) ld4(p.h0 + o, h);
      }
      store_pair4(reinterpret_cast<__half*>(st_hx) + rloc * 16 + part * 4, reinterpret_cast<__half*>(st_hx + 4096) + rloc * 16 + part * 4, h, s_h);
      st4(st_hin + rloc * 16 + part * 4, h);
      if (p.gates) st4(p.cin + ((((int64_t)0 * p.tiles + tile) * nquads + quad) * BM + rloc) * 4, c);
      epi_bar_sync();
      flush(tile, -1, 0, false, true);   // hin[0] = h0
      epi_bar_sync();
      if (leader) red_release_add(flag, 1u);

      for (int t = 0; t < T; ++t, ++it) {
        uint8_t dn = 0;
        if (valid && p.done) dn = p.done[(int64_t)t * p.Nb + row];
        mbar_wait(x_full, it & 1);
        float x[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // SWIZZLE_64B: 16-byte unit u of 64-byte row r lives at unit u ^ ((r >> 1) & 3)
          const float4 v = *reinterpret_cast<const float4*>(sX + q * (BM * LS_HS * 4) + rloc * 64 + ((part ^ ((rloc >> 1) & 3)) * 16));
          x[q][0] = v.x, x[q][1] = v.y, x[q][2] = v.z, x[q][3] = v.w;
        }
        mbar_wait(tfull, it & 1);
        tc_fence_after();
        uint32_t r[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x4(taddr + (uint32_t)(q * LS_HS), r[q]);
        tmem_ld_wait();
        tc_fence_before();
        float g[4][4], cs[4], hs[4];
        const float m = dn ? 0.f : 1.f;   // the state handed to step t + 1 restarts where the episode ended
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pre = x[q][j] + fmaf(__uint_as_float(r[q][j]), inv_ab, bias[q][j]);
            g[q][j] = q == 2 ? tanhf(pre) : sigmoid_acc(pre);
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cs[j] = g[1][j] * c[j] + g[0][j] * g[2][j];
          hs[j] = g[3][j] * tanhf(cs[j]);
          c[j] = cs[j] * m, h[j] = hs[j] * m;
        }
        const bool more = t + 1 < T;
        // stage the row-major outputs (the operand tile is idle: this step's MMAs have completed)
        store_pair4(reinterpret_cast<__half*>(st_hx) + rloc * 16 + part * 4, reinterpret_cast<__half*>(st_hx + 4096) + rloc * 16 + part * 4, h, s_h);
        st4(st_out + rloc * 16 + part * 4, hs);
        st4(st_hin + rloc * 16 + part * 4, h);
        epi_bar_sync();   // staged tiles complete; every thread has consumed the input-projection tile
        if (leader && (more || tile + p.groups < p.tiles)) load_x(more ? tile : tile + p.groups, more ? t + 1 : 0);
        flush(tile, t, (t + 1) & 1, !(p.debug & 1), more && !(p.debug & 1));
        // publish: this slice's part of the state entering step t + 1 is on its way to L2, this CTA has drained its
        // accumulator and no longer reads its operand tile
        epi_bar_sync();
        if (leader) red_release_add(flag + t + 1, 1u);
        if (!more && valid) {   // the next memory, row-major (once per launch)
          if (p.c_last) st4(p.c_last + (int64_t)row * p.ld_last + slice * LS_HS + part * 4, cs);
          if (p.h_last) st4(p.h_last + (int64_t)row * p.ld_last + slice * LS_HS + part * 4, hs);
        }
        if (p.gates && !(p.debug & 1)) {   // private layouts (rows of a warp contiguous), after the hand-over; inference: none
          float* gr = p.gates + ((((int64_t)t * p.tiles + tile) * nquads + quad) * 4 * BM + rloc) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) st4(gr + q * BM * 4, g[q]);
          st4(p.cseq + ((((int64_t)t * p.tiles + tile) * nquads + quad) * BM + rloc) * 4, cs);
          if (more) st4(p.cin + ((((int64_t)(t + 1) * p.tiles + tile) * nquads + quad) * BM + rloc) * 4, c);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();   // no CTA leaves while a peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int g_lstm_debug = 0;   // timing experiments only (cusrl_b200_lstm_seq_set_debug): results are wrong when non-zero

constexpr int LS_SMEM_BYTES = LS_MAX_KB * 2 * (LS_W_KB_BYTES + LS_A_KB_BYTES) + 4 * BM * LS_HS * 4 + 512 + 1024;

static bool lstm_seq_shape_ok(int64_t H) { return H > 0 && (H % LS_KB) == 0 && H <= LS_MAX_KB * LS_KB; }

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_lstm_seq_set_debug(int bits) {
  g_lstm_debug = bits;
  return 0;
}

int cusrl_b200_lstm_seq_supported(int64_t H) { return lstm_seq_shape_ok(H) ? 1 : 0; }

size_t cusrl_b200_lstm_seq_workspace_bytes(int64_t T, int64_t Nb, int64_t H) {
  if (T <= 0 || Nb <= 0 || !lstm_seq_shape_ok(H)) return 0;
  const int64_t tiles = (Nb + BM - 1) / BM;
  const size_t flags = (size_t)((tiles * (T + 1) * 4 + 255) / 256 * 256);
  return flags + (size_t)(2 * 2 * tiles * BM * H) * sizeof(uint16_t);
}

int cusrl_b200_lstm_seq_fwd_f32(const float* xp, int64_t ldxp, const uint16_t* Whi, const uint16_t* Wlo, int64_t ldw,
                                const float* w_stats, const float* b_hh, const float* h0, const float* c0, const uint8_t* done,
                                int64_t ld0, float* gates, float* cseq, float* out, float* hin, float* cin, float* h_last, float* c_last,
                                int64_t ld_last, int64_t T, int64_t Nb, int64_t H, void* workspace, size_t workspace_bytes,
                                void* stream) {
  CUSRL_REQUIRE(xp && Whi && Wlo && w_stats && out && workspace, CUSRL_B200_EINVAL, "lstm_seq_fwd: null pointer");
  CUSRL_REQUIRE((gates != nullptr) == (cseq != nullptr) && (gates != nullptr) == (cin != nullptr), CUSRL_B200_EINVAL,
                "lstm_seq_fwd: gates / cseq / cin are given together (training) or not at all (inference)");
  CUSRL_REQUIRE(((!h0 && !c0) || (ld0 >= H && (ld0 % 4) == 0)) && ((!h_last && !c_last) || (ld_last >= H && (ld_last % 4) == 0)),
                CUSRL_B200_EALIGN, "lstm_seq_fwd: memory row pitches must be multiples of 4 floats covering H");
  CUSRL_REQUIRE(T > 0 && Nb > 0 && T < (1 << 20) && Nb < (1ll << 30), CUSRL_B200_EINVAL, "lstm_seq_fwd: bad sizes");
  CUSRL_REQUIRE(lstm_seq_shape_ok(H), CUSRL_B200_EUNSUPPORTED, "lstm_seq_fwd: H must be a multiple of 64, at most 256 (got %lld)",
                (long long)H);
  CUSRL_REQUIRE((ldxp % 4) == 0 && ldxp >= 4 * H && ldw >= H && (ldw % 8) == 0, CUSRL_B200_EALIGN, "lstm_seq_fwd: leading dimensions");
  CUSRL_REQUIRE(aligned_to(xp, 16) && aligned_to(Whi, 16) && aligned_to(Wlo, 16) && (!gates || aligned_to(gates, 16)) &&
                    (!cseq || aligned_to(cseq, 16)) && aligned_to(out, 16) && (!hin || aligned_to(hin, 16)) &&
                    (!cin || aligned_to(cin, 16)) && (!h0 || aligned_to(h0, 16)) && (!c0 || aligned_to(c0, 16)) &&
                    (!c_last || aligned_to(c_last, 16)) && (!h_last || aligned_to(h_last, 16)) && aligned_to(workspace, 256),
                CUSRL_B200_EALIGN, "lstm_seq_fwd: pointers must be 16-byte aligned (workspace: 256)");
  const size_t need = cusrl_b200_lstm_seq_workspace_bytes(T, Nb, H);
  CUSRL_REQUIRE(workspace_bytes >= need, CUSRL_B200_ESCRATCH, "lstm_seq_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  LstmSeqFwdParams p{};
  p.tiles = (int)((Nb + BM - 1) / BM);
  p.slices = (int)(H / LS_HS);
  static bool configured = false;
  static int max_clusters = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM_BYTES);
    CUSRL_REQUIRE(e == cudaSuccess, (int)e, "lstm_seq_fwd: cudaFuncSetAttribute(%d bytes): %s", LS_SMEM_BYTES, cudaGetErrorString(e));
    // every CTA of a row tile spins on its peers: the whole grid must be co-resident, and clusters cannot straddle GPCs, so
    // the grid is bounded by the number of clusters the device can hold at once, not by the SM count
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sm_count() / LS_CLUSTER * LS_CLUSTER));
    cfg.blockDim = dim3(LS_THREADS);
    cfg.dynamicSmemBytes = LS_SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = LS_CLUSTER, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_seq_fwd_kernel, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = (sm_count() - 24) / LS_CLUSTER;   // conservative: up to three SMs per GPC unusable by 4-CTA clusters
    }
    max_clusters = n;
    configured = true;
  }
  const int max_groups = max_clusters * LS_CLUSTER / p.slices;
  CUSRL_REQUIRE(max_groups >= 1, CUSRL_B200_EUNSUPPORTED, "lstm_seq_fwd: not enough SMs for one row tile");
  p.groups = p.tiles < max_groups ? p.tiles : max_groups;
  p.nbp = p.tiles * BM;
  const size_t flag_bytes = (size_t)(((int64_t)p.tiles * (T + 1) * 4 + 255) / 256 * 256);
  p.flags = (unsigned int*)workspace;
  p.hx_hi = (__half*)((uint8_t*)workspace + flag_bytes);
  p.hx_lo = p.hx_hi + (size_t)2 * p.nbp * H;
  // only the counters are cleared: exchange rows >= Nb of the last tile are never written and hold whatever the workspace
  // held, but a row of the product depends on its own row of the operand only, and those rows' results are discarded
  cudaError_t me = cudaMemsetAsync(workspace, 0, flag_bytes, s);
  CUSRL_REQUIRE(me == cudaSuccess, (int)me, "lstm_seq_fwd: cudaMemsetAsync: %s", cudaGetErrorString(me));
  p.xp = xp, p.ldxp = ldxp, p.b_hh = b_hh, p.h0 = h0, p.c0 = c0, p.done = done;
  p.gates = gates, p.cseq = cseq, p.out = out, p.hin = hin, p.cin = cin, p.c_last = c_last, p.h_last = h_last, p.wstats = w_stats;
  p.ld0 = ld0, p.ld_last = ld_last;
  p.T = (int)T, p.Nb = (int)Nb, p.H = (int)H, p.debug = g_lstm_debug;
  CUtensorMap tWh, tWl, tAh, tAl, tX;
  if (int e = encode_tmap_2d_f32(&tX, xp, (uint64_t)(4 * H), (uint64_t)(T * Nb), (uint64_t)ldxp, LS_HS, BM, TMAP_SW64)) return e;
  if (int e = encode_tmap_2d_f16(&tWh, Whi, (uint64_t)H, (uint64_t)(4 * H), (uint64_t)ldw, LS_KB, LS_HS, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tWl, Wlo, (uint64_t)H, (uint64_t)(4 * H), (uint64_t)ldw, LS_KB, LS_HS, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tAh, p.hx_hi, (uint64_t)H, (uint64_t)(2 * p.nbp), (uint64_t)H, LS_KB, BM, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tAl, p.hx_lo, (uint64_t)H, (uint64_t)(2 * p.nbp), (uint64_t)H, LS_KB, BM, TMAP_SW128)) return e;
  lstm_seq_fwd_kernel<<<p.groups * p.slices, LS_THREADS, LS_SMEM_BYTES, s>>>(tWh, tWl, tAh, tAl, tX, p);
  return check_launch("lstm_seq_fwd_kernel");
}

}  // extern "C"

namespace cusrl_b200 {

// =================================================================================================
// Backward through time of one layer in ONE launch.
//
//   dgates_t = cell'(dh_above_t + mask_t * dh_rec_t, mask_t * dc_rec_t)          elementwise (lstm_cell_bwd_kernel's arithmetic)
//   dh_rec_{t-1} = dgates_t @ W_hh                                               [Nb, 4H] x [4H, H]
//
// The reduction runs over the GATE axis, which the elementwise step produces, and the result is needed along the HIDDEN axis,
// which the next elementwise step consumes: a CTA owning 32 hidden units (all 128 of their gate columns, 128 batch rows)
// holds a K-slice of the product.  It multiplies its dgates slice [128 x 128] (split into an fp16 pair with a scale taken from
// the slice's own maximum, written to shared memory in the K-major SWIZZLE_128B layout the MMA reads) by its rows of W_hh
// (resident, [H x 128] pair, gate-interleaved copy made by lstm_wt_perm_kernel) into a [128 x H] fp32 partial product, and
// hands it to the H/32 CTAs of its row tile through L2 (double-buffered by step parity, one counter per (tile, step) as in
// the forward kernel).  Each consumer sums the H/32 partials of its 32 columns in a FIXED order: deterministic.
// dc is carried in registers; dgates_t is written to HBM after the hand-over (the weight gradients and the gradient
// w.r.t. the layer input are large GEMMs over all T * Nb rows afterwards).
constexpr int LB_HS = 32;
constexpr int LB_KS = 4 * LB_HS;           // 128 gate columns per CTA = K of its partial product
constexpr int LB_NKB = LB_KS / LS_KB;      // 2 k-blocks
constexpr int LB_EPI_WARPS = 16;
constexpr int LB_THREADS = 128 + 32 * LB_EPI_WARPS;
constexpr int LB_MAX_H = 256;
constexpr int LB_B_KB_BYTES = LB_MAX_H * LS_KB * 2;   // 32 KB per half per k-block
constexpr int LB_SMEM_BYTES = LB_NKB * 2 * (LB_B_KB_BYTES + LS_A_KB_BYTES) + 512 + 1024;

struct LstmSeqBwdParams {
  const float* dout;      // [T * Nb, H] gradient w.r.t. h_t from the layer above / the head, row pitch lddo
  int64_t lddo;
  const float* gates;     // [T, Nb, 4H] activated gates (forward)
  const float* cseq;      // [T, Nb, H]
  const float* cin;       // [T, Nb, H]
  const uint8_t* done;    // [T, Nb] or null
  float* dgates;          // [T, Nb, 4H] pre-activation gate gradients
  float* part;            // [2, tiles, slices, 128, H] partial products
  const float* wstats;
  unsigned int* flags;    // [tiles, T + 1]
  int T, Nb, H, tiles, slices, groups;
  int nsplit;             // CTAs per slice: 2 = each computes HALF of the partial product's H columns (H >= 128), else 1
  int debug;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void lb_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * LB_EPI_WARPS) : "memory"); }

__global__ void __launch_bounds__(LB_THREADS, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const LstmSeqBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                     // [2 kb][hi | lo][H rows x 128 B]
  uint8_t* sA = smem + LB_NKB * 2 * LB_B_KB_BYTES;        // [2 kb][hi | lo][128 rows x 128 B]
  uint8_t* tail = sA + LB_NKB * 2 * LS_A_KB_BYTES;
  unsigned int* s_amax = reinterpret_cast<unsigned int*>(tail);   // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 64);
  uint64_t* b_full = bars;
  uint64_t* a_ready = bars + 1;
  uint64_t* tfull = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a slice's elementwise work (its 128 gate gradients per row) is replicated in its `nsplit` CTAs; each forms the partial
  // product for its own Hn = H / nsplit output columns: twice the CTAs, half the MMA time and half the partial-product store
  // per CTA on the critical path
  const int per_tile = p.slices * p.nsplit;
  const int slice = (blockIdx.x % per_tile) / p.nsplit, nhalf = blockIdx.x % p.nsplit, group = blockIdx.x / per_tile;
  const int T = p.T, H = p.H, Hn = H / p.nsplit;

  if (threadIdx.x == 0) s_amax[0] = 0u, s_amax[1] = 0u;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmBhi);
    tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(b_full, 1);
    mbar_init(a_ready, 1);
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // the CTA's rows of W_hh (as columns of the gate-interleaved transposed pair), once
      mbar_expect_tx(b_full, (uint32_t)(LB_NKB * 2 * Hn * LS_KB * 2));
      for (int kb = 0; kb < LB_NKB; ++kb) {
        tma_load_2d(sB + (kb * 2 + 0) * LB_B_KB_BYTES, &tmBhi, slice * LB_KS + kb * LS_KB, nhalf * Hn, b_full);
        tma_load_2d(sB + (kb * 2 + 1) * LB_B_KB_BYTES, &tmBlo, slice * LB_KS + kb * LS_KB, nhalf * Hn, b_full);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BM, Hn, 0, 0);
      mbar_wait(b_full, 0);
      uint32_t it = 0;
      for (int tile = group; tile < p.tiles; tile += p.groups)
        for (int t = T - 1; t >= 1; --t, ++it) {
          mbar_wait(a_ready, it & 1);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < LB_NKB; ++kb) {
            const uint32_t ahi = smem_u32(sA + (kb * 2 + 0) * LS_A_KB_BYTES), alo = smem_u32(sA + (kb * 2 + 1) * LS_A_KB_BYTES);
            const uint32_t bhi = smem_u32(sB + (kb * 2 + 0) * LB_B_KB_BYTES), blo = smem_u32(sB + (kb * 2 + 1) * LB_B_KB_BYTES);
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
    // ===================================== elementwise + hand-over ==========================
    const int ew = warp & 3;                  // TMEM lane quarter
    const int part = (warp - 4) >> 2;         // 0..3: which 8 of the CTA's 32 hidden units / which quarter of the H columns
    const int u0 = slice * LB_HS + part * 8;
    const int quad0 = u0 / 4;
    const int rloc = ew * 32 + lane;
    const int etid = threadIdx.x - 128;
    const float s_w = f16x3_scale(__ldg(p.wstats + WSTAT_AMAX));
    const int cols_per_part = Hn / 4;         // 16, 32, 48 or 64 columns of this CTA's Hn per epilogue-warp group
    const bool leader = threadIdx.x == 128;
    uint32_t it = 0;
    for (int tile = group; tile < p.tiles; tile += p.groups) {
      unsigned int* flag = p.flags + (int64_t)tile * (T + 1);
      const int row = tile * BM + rloc;
      const bool valid = row < p.Nb;
      float dc_carry[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dc_carry[j] = 0.f;
      for (int t = T - 1; t >= 0; --t) {
        float dh[8], g[4][8], cv[8], cpv[8];
        float m = 1.f;
        const int64_t qb = ((int64_t)t * p.tiles + tile) * (H / 4) + quad0;
        {
          // the forward kernel's tensors of this step (private layout: the 32 rows of a warp are contiguous) are pulled into
          // L2 now and LOADED after the partial products: holding them in registers across the hand-over wait spilled
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const float* gp = p.gates + (((qb + hq) * 4) * BM + rloc) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) prefetch_l2(gp + q * BM * 4);
            prefetch_l2(p.cseq + ((qb + hq) * BM + rloc) * 4);
            prefetch_l2(p.cin + ((qb + hq) * BM + rloc) * 4);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dh[j] = 0.f;
        if (valid) {
          const int64_t rt = (int64_t)t * p.Nb + row;
          const float* dp = p.dout + rt * p.lddo + u0;
          const float4 d0 = __ldg(reinterpret_cast<const float4*>(dp)), d1 = __ldg(reinterpret_cast<const float4*>(dp + 4));
          dh[0] = d0.x, dh[1] = d0.y, dh[2] = d0.z, dh[3] = d0.w, dh[4] = d1.x, dh[5] = d1.y, dh[6] = d1.z, dh[7] = d1.w;
          if (p.done && t < T - 1 && p.done[rt]) m = 0.f;
        }
        float dc[8];
        if (t < T - 1) {
          // the partial products of step t + 1 from every slice of this row tile
          if (leader)
            while (ld_acquire_u32(flag + t + 1) < (unsigned int)per_tile) {
            }
          lb_bar_sync();
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          {
            // [parity][tile][source slice][H/4 quads][128 rows][4]: a warp reads 512 contiguous bytes per quad
            const float* base = p.part + ((((int64_t)((t + 1) & 1) * p.tiles + tile) * p.slices) * (H / 4) + quad0) * BM * 4 + rloc * 4;
            // fixed order: deterministic.  Four sources per round trip: a rolled loop issued one pair of loads per L2 latency
            // (ncu: the adds of this loop were the top long-scoreboard stall after the counter poll)
            for (int s0 = 0; s0 < p.slices; s0 += 4) {
              float4 a[4], b[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (s0 + u < p.slices) {
                  a[u] = ld_cg4(base + (int64_t)(s0 + u) * H * BM);
                  b[u] = ld_cg4(base + (int64_t)(s0 + u) * H * BM + BM * 4);
                }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (s0 + u < p.slices) {
                  acc[0] += a[u].x, acc[1] += a[u].y, acc[2] += a[u].z, acc[3] += a[u].w;
                  acc[4] += b[u].x, acc[5] += b[u].y, acc[6] += b[u].z, acc[7] += b[u].w;
                }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dh[j] += m * acc[j], dc[j] = m * dc_carry[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dc[j] = 0.f;
        }
        {
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const float* gp = p.gates + (((qb + hq) * 4) * BM + rloc) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(gp + q * BM * 4));
              g[q][4 * hq] = a.x, g[q][4 * hq + 1] = a.y, g[q][4 * hq + 2] = a.z, g[q][4 * hq + 3] = a.w;
            }
            const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cseq + ((qb + hq) * BM + rloc) * 4));
            cv[4 * hq] = cc.x, cv[4 * hq + 1] = cc.y, cv[4 * hq + 2] = cc.z, cv[4 * hq + 3] = cc.w;
            const float4 ce = __ldg(reinterpret_cast<const float4*>(p.cin + ((qb + hq) * BM + rloc) * 4));
            cpv[4 * hq] = ce.x, cpv[4 * hq + 1] = ce.y, cpv[4 * hq + 2] = ce.z, cpv[4 * hq + 3] = ce.w;
          }
        }
        float dg[4][8];
        float amax = 0.f;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float iv = g[0][j], fv = g[1][j], gv = g[2][j], ov = g[3][j];
            const float tc = tanhf(cv[j]);
            const float dct = dc[j] + dh[j] * ov * (1.f - tc * tc);
            dg[3][j] = dh[j] * tc * ov * (1.f - ov);
            dg[0][j] = dct * gv * iv * (1.f - iv);
            dg[1][j] = dct * cpv[j] * fv * (1.f - fv);
            dg[2][j] = dct * iv * (1.f - gv * gv);
            dc_carry[j] = dct * fv;
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(dg[0][j]), fabsf(dg[1][j]))), fmaxf(fabsf(dg[2][j]), fabsf(dg[3][j])));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) dg[q][j] = 0.f;
        }
        if (t >= 1) {
          // scale of this step's operand pair: the exact maximum of the CTA's slice
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
          unsigned int* slot = s_amax + (it & 1);
          if (lane == 0) atomicMax(slot, __float_as_uint(amax));
          lb_bar_sync();
          const float s_a = f16x3_scale(__uint_as_float(*slot));
          if (leader) s_amax[(it & 1) ^ 1] = 0u;   // the other slot is idle until the next step's barrier
          // operand tile: k = gate * 32 + unit -> k-block gate / 2, 16-byte chunk (gate % 2) * 4 + part of the 128-byte row
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 hv, lv;
            __half2* h2 = reinterpret_cast<__half2*>(&hv);
            __half2* l2 = reinterpret_cast<__half2*>(&lv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = dg[q][2 * k] * s_a, b = dg[q][2 * k + 1] * s_a;
              h2[k] = __floats2half2_rn(a, b);
              const float2 back = __half22float2(h2[k]);
              l2[k] = __floats2half2_rn(a - back.x, b - back.y);
            }
            const int kb = q >> 1, chunk = ((q & 1) * 4 + part) ^ (rloc & 7);
            const int off = (rloc >> 3) * 1024 + (rloc & 7) * 128 + chunk * 16;
            *reinterpret_cast<uint4*>(sA + (kb * 2 + 0) * LS_A_KB_BYTES + off) = hv;
            *reinterpret_cast<uint4*>(sA + (kb * 2 + 1) * LS_A_KB_BYTES + off) = lv;
          }
          fence_proxy_async_smem();
          lb_bar_sync();
          if (leader) mbar_arrive(a_ready);
          // partial product -> L2
          mbar_wait(tfull, it & 1);
          tc_fence_after();
          const float inv = 1.f / (s_a * s_w);
          float* dst = p.part + (((((int64_t)(t & 1) * p.tiles + tile) * p.slices + slice) * (H / 4)) + (nhalf * Hn + part * cols_per_part) / 4) * BM * 4 + rloc * 4;
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(part * cols_per_part);
          for (int c0 = 0; c0 < cols_per_part; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(taddr + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 4; ++v)
              *reinterpret_cast<float4*>(dst + (c0 / 4 + v) * BM * 4) =
                  make_float4(__uint_as_float(r[4 * v]) * inv, __uint_as_float(r[4 * v + 1]) * inv,
                              __uint_as_float(r[4 * v + 2]) * inv, __uint_as_float(r[4 * v + 3]) * inv);
          }
          tc_fence_before();
          lb_bar_sync();
          if (leader) red_release_add(flag + t, 1u);
          ++it;
        }
        if (!(p.debug & 2) && nhalf == 0) {   // (uniform per CTA: the barrier counts stay consistent)
          // off the critical path: dgates_t is consumed row-major by the weight-gradient / input-gradient GEMMs.  It is
          // transposed through the operand tile (idle: this step's MMAs are complete, the next step's operand is written two
          // barriers from here): [128 rows][4 gates][32 units] fp32, 16-byte units XOR-swizzled by the row within each
          // 128-byte gate segment, then written 16 bytes per thread along rows (a warp covers one row's four segments)
          float* stg = reinterpret_cast<float*>(sA);
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int hq = 0; hq < 2; ++hq)
              *reinterpret_cast<float4*>(stg + rloc * 128 + q * 32 + (((part * 2 + hq) ^ (rloc & 7)) * 4)) =
                  make_float4(dg[q][4 * hq], dg[q][4 * hq + 1], dg[q][4 * hq + 2], dg[q][4 * hq + 3]);
          lb_bar_sync();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int idx = k * (32 * LB_EPI_WARPS) + etid;
            const int r = idx >> 5, unit = idx & 31, q = unit >> 3, ch = unit & 7;
            const int grow = tile * BM + r;
            if (grow < p.Nb)
              *reinterpret_cast<float4*>(p.dgates + ((int64_t)t * p.Nb + grow) * 4 * H + q * H + slice * LB_HS + ch * 4) =
                  *reinterpret_cast<const float4*>(stg + r * 128 + q * 32 + ((ch ^ (r & 7)) * 4));
          }
          lb_bar_sync();   // the staging tile is free again before anyone writes the next operand into it
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// W_hh^T pair [H, 4H] (weight_prep_f16's transposed pair) -> gate-interleaved copy: out[n][s * 128 + q * 32 + u] =
// in[n][q * H + s * 32 + u], so that the 128 reduction indices a backward CTA owns are 128 consecutive columns.
__global__ void __launch_bounds__(256) lstm_wt_perm_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int64_t ldi,
                                                           __half* __restrict__ out_hi, __half* __restrict__ out_lo, int H) {
  const int chunks_per_row = 4 * H / 8;   // 16-byte chunks of 8 halves
  const int total = H * chunks_per_row;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / chunks_per_row, c = i - n * chunks_per_row;
    const int col = c * 8;                       // destination column
    const int s = col / LB_KS, q = (col % LB_KS) / LB_HS, u = col % LB_HS;
    const int64_t src = (int64_t)n * ldi + q * H + s * LB_HS + u;
    const int64_t dst = (int64_t)n * 4 * H + col;
    *reinterpret_cast<uint4*>(out_hi + dst) = *reinterpret_cast<const uint4*>(in_hi + src);
    *reinterpret_cast<uint4*>(out_lo + dst) = *reinterpret_cast<const uint4*>(in_lo + src);
  }
}

}  // namespace cusrl_b200

extern "C" {

size_t cusrl_b200_lstm_seq_bwd_workspace_bytes(int64_t T, int64_t Nb, int64_t H) {
  if (T <= 0 || Nb <= 0 || !lstm_seq_shape_ok(H)) return 0;
  const int64_t tiles = (Nb + BM - 1) / BM, slices = H / LB_HS;
  const size_t flags = (size_t)((tiles * (T + 1) * 4 + 255) / 256 * 256);
  const size_t perm = (size_t)(2 * H * 4 * H) * sizeof(uint16_t);                 // gate-interleaved W_hh^T pair
  const size_t part = (size_t)(2 * tiles * slices * BM * H) * sizeof(float);
  return flags + perm + part;
}

int cusrl_b200_lstm_seq_bwd_f32(const float* dout, int64_t lddo, const float* gates, const float* cseq, const float* cin,
                                const uint8_t* done, const uint16_t* WThi, const uint16_t* WTlo, int64_t ldwt, const float* w_stats,
                                float* dgates, int64_t T, int64_t Nb, int64_t H, void* workspace, size_t workspace_bytes,
                                void* stream) {
  CUSRL_REQUIRE(dout && gates && cseq && cin && WThi && WTlo && w_stats && dgates && workspace, CUSRL_B200_EINVAL,
                "lstm_seq_bwd: null pointer");
  CUSRL_REQUIRE(T > 0 && Nb > 0 && T < (1 << 20) && Nb < (1ll << 30), CUSRL_B200_EINVAL, "lstm_seq_bwd: bad sizes");
  CUSRL_REQUIRE(lstm_seq_shape_ok(H), CUSRL_B200_EUNSUPPORTED, "lstm_seq_bwd: H must be a multiple of 64, at most 256 (got %lld)",
                (long long)H);
  CUSRL_REQUIRE((lddo % 4) == 0 && lddo >= H && ldwt >= 4 * H && (ldwt % 8) == 0, CUSRL_B200_EALIGN, "lstm_seq_bwd: leading dimensions");
  CUSRL_REQUIRE(aligned_to(dout, 16) && aligned_to(gates, 16) && aligned_to(cseq, 16) && aligned_to(cin, 16) && aligned_to(WThi, 16) &&
                    aligned_to(WTlo, 16) && aligned_to(dgates, 16) && aligned_to(workspace, 256),
                CUSRL_B200_EALIGN, "lstm_seq_bwd: pointers must be 16-byte aligned (workspace: 256)");
  const size_t need = cusrl_b200_lstm_seq_bwd_workspace_bytes(T, Nb, H);
  CUSRL_REQUIRE(workspace_bytes >= need, CUSRL_B200_ESCRATCH, "lstm_seq_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t s = (cudaStream_t)stream;
  LstmSeqBwdParams p{};
  p.tiles = (int)((Nb + BM - 1) / BM);
  p.slices = (int)(H / LB_HS);
  // Two CTAs per slice cut a tile-step from 13.9 to 11.1 us (measured, H = 256) but halve the number of resident row-tile
  // groups: worth it while the tiles still fit in as many rounds (the training minibatch: 8 tiles, one round either way).
  // H / nsplit / 4 columns per epilogue-warp group must be a multiple of 16 -> only for H = 128, 256.
  p.nsplit = 1;
  if ((H % 128) == 0 && sm_count() / (p.slices * 2) >= 1) {
    const int g1 = sm_count() / p.slices, g2 = sm_count() / (p.slices * 2);
    const int rounds1 = (p.tiles + g1 - 1) / g1, rounds2 = (p.tiles + g2 - 1) / g2;
    if (rounds2 * 4 <= rounds1 * 5) p.nsplit = 2;
  }
  const int max_groups = sm_count() / (p.slices * p.nsplit);
  CUSRL_REQUIRE(max_groups >= 1, CUSRL_B200_EUNSUPPORTED, "lstm_seq_bwd: not enough SMs for one row tile");
  p.groups = p.tiles < max_groups ? p.tiles : max_groups;
  const size_t flag_bytes = (size_t)(((int64_t)p.tiles * (T + 1) * 4 + 255) / 256 * 256);
  p.flags = (unsigned int*)workspace;
  __half* perm_hi = (__half*)((uint8_t*)workspace + flag_bytes);
  __half* perm_lo = perm_hi + (size_t)H * 4 * H;
  p.part = (float*)(perm_lo + (size_t)H * 4 * H);
  cudaError_t me = cudaMemsetAsync(workspace, 0, flag_bytes, s);
  CUSRL_REQUIRE(me == cudaSuccess, (int)me, "lstm_seq_bwd: cudaMemsetAsync: %s", cudaGetErrorString(me));
  lstm_wt_perm_kernel<<<(int)((H * 4 * H / 8 + 255) / 256), 256, 0, s>>>((const __half*)WThi, (const __half*)WTlo, ldwt, perm_hi, perm_lo, (int)H);
  if (int e = check_launch("lstm_wt_perm_kernel")) return e;
  p.dout = dout, p.lddo = lddo, p.gates = gates, p.cseq = cseq, p.cin = cin, p.done = done, p.dgates = dgates, p.wstats = w_stats;
  p.T = (int)T, p.Nb = (int)Nb, p.H = (int)H, p.debug = g_lstm_debug;
  CUtensorMap tBh, tBl;
  const uint32_t box_rows = (uint32_t)(H / p.nsplit);
  if (int e = encode_tmap_2d_f16(&tBh, perm_hi, (uint64_t)(4 * H), (uint64_t)H, (uint64_t)(4 * H), LS_KB, box_rows, TMAP_SW128)) return e;
  if (int e = encode_tmap_2d_f16(&tBl, perm_lo, (uint64_t)(4 * H), (uint64_t)H, (uint64_t)(4 * H), LS_KB, box_rows, TMAP_SW128)) return e;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM_BYTES);
    CUSRL_REQUIRE(e == cudaSuccess, (int)e, "lstm_seq_bwd: cudaFuncSetAttribute(%d bytes): %s", LB_SMEM_BYTES, cudaGetErrorString(e));
    configured = true;
  }
  lstm_seq_bwd_kernel<<<p.groups * p.slices * p.nsplit, LB_THREADS, LB_SMEM_BYTES, s>>>(tBh, tBl, p);
  return check_launch("lstm_seq_bwd_kernel");
}

}  // extern "C"
