// K1, TMA-staged variant: hook/on_policy/gae.py:8-20,85-110 on [T,N,1] leaves.
//
// The register-resident kernel (rollout_kernels.cu) issues every load of a column through the LSU; at 65536 x 24 it
// stops at ~55 % of the HBM copy rate.  This variant moves the bytes through the TMA unit instead.  MEASURED (round 1,
// profiles/r01_kbench_gae_variants.jsonl): 9.5-10.4 us per launch against 9.3-9.5 us for the register kernel and 7.7 us
// for a device-to-device copy of the same 33 MB -- the request path is not what limits a launch this short, so the
// register kernel stays the default and this one is selectable (cusrl_b200_gae_set_variant) for larger rollouts.
//   * a CTA of W warps owns column tiles of C = 32 W environments; one tile = the boxes {C columns x T rows} of
//     reward / value / next_value (fp32) and done (u8), i.e. the WHOLE rollout of C environments (13 T C bytes);
//   * an elected thread issues the four bulk-tensor loads of a tile into one shared-memory stage (mbarrier
//     complete_tx); up to `stages` tiles are in flight per CTA before the first dependent instruction;
//   * thread c walks column c of the stage backwards in time with the reference's exact op order (no FMA contraction),
//     overwriting the reward tile with the advantage and the value tile with the return;
//   * two bulk-tensor stores write the tiles back (rows of C*4 contiguous bytes, clipped at N), asynchronously, while the
//     CTA is already working on its next stage.
// Lanes run along the contiguous env axis, so shared-memory accesses are conflict-free without swizzling.
#include "gae_common.cuh"
#include "tc_common.cuh"

namespace cusrl_b200 {

using namespace tc;

struct GaeTmaParams {
  int T, num_tiles, stages;
  uint32_t stage_bytes;
  float gamma, c_adv, c_ret;
  int two_lambda, has_ret;
};

constexpr int kGaeTmaU = 8;            // time steps whose shared-memory loads are issued together
constexpr int kGaeTmaMaxStages = 8;
constexpr int kGaeTmaHeader = 128;     // mbarriers live in front of the (128-byte aligned) stages

__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(256) gae_tma_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmV,
                                                      const __grid_constant__ CUtensorMap tmNV, const __grid_constant__ CUtensorMap tmD,
                                                      const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmRet,
                                                      const GaeTmaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* stage0 = smem + kGaeTmaHeader;
  const int C = blockDim.x, T = p.T, c = threadIdx.x;
  const int TC = T * C;
  const uint32_t fbytes = (uint32_t)TC * 4u;
  const int my_tiles = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  auto issue_load = [&](int j) {  // elected thread only
    const int s = j % p.stages;
    uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
    const int col0 = ((int)blockIdx.x + j * (int)gridDim.x) * C;
    mbar_expect_tx(&full[s], 3u * fbytes + (uint32_t)TC);
    tma_load_2d(st, &tmR, col0, 0, &full[s]);
    tma_load_2d(st + fbytes, &tmV, col0, 0, &full[s]);
    tma_load_2d(st + 2 * fbytes, &tmNV, col0, 0, &full[s]);
    tma_load_2d(st + 3 * fbytes, &tmD, col0, 0, &full[s]);
  };

  if (c == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
    const int first = my_tiles < p.stages ? my_tiles : p.stages;
    for (int j = 0; j < first; ++j) issue_load(j);
  }
  __syncthreads();

  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % p.stages;
    uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
    float* R = reinterpret_cast<float*>(st);
    float* V = R + TC;
    const float* NV = V + TC;
    const uint8_t* D = st + 3 * fbytes;
    mbar_wait(&full[s], (uint32_t)(i / p.stages) & 1u);

    float adv_next = 0.f, adv2_next = 0.f;
    for (int t_hi = T; t_hi > 0; t_hi -= kGaeTmaU) {
      float r[kGaeTmaU], v[kGaeTmaU], nv[kGaeTmaU];
      uint8_t d[kGaeTmaU];
#pragma unroll
      for (int u = 0; u < kGaeTmaU; ++u) {
        const int t = t_hi - 1 - u;
        if (t >= 0) {
          const int o = t * C + c;
          r[u] = R[o], v[u] = V[o], nv[u] = NV[o], d[u] = D[o];
        }
      }
#pragma unroll
      for (int u = 0; u < kGaeTmaU; ++u) {
        const int t = t_hi - 1 - u;
        if (t >= 0) {
          // gae.py:17  advantage = reward + next_value * gamma - value
          const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(nv[u], p.gamma)), v[u]);
          float a = delta, a2 = delta;
          if (t != T - 1) {
            // gae.py:19  advantage[t] += not_done[t] * (gamma*lamda) * advantage[t+1]
            a = __fadd_rn(delta, __fmul_rn(d[u] ? 0.f : p.c_adv, adv_next));
            if (p.two_lambda) a2 = __fadd_rn(delta, __fmul_rn(d[u] ? 0.f : p.c_ret, adv2_next));
          }
          adv_next = a, adv2_next = a2;
          const int o = t * C + c;
          R[o] = a;
          // gae.py:99-110  return = value + advantage (or the lamda_value scan)
          V[o] = __fadd_rn(v[u], p.two_lambda ? a2 : a);
        }
      }
    }
    fence_proxy_async_smem();  // the generic-proxy writes above must be visible to the bulk stores
    __syncthreads();
    if (c == 0) {
      const int col0 = ((int)blockIdx.x + i * (int)gridDim.x) * C;
      tma_store_2d(&tmA, R, col0, 0);
      if (p.has_ret) tma_store_2d(&tmRet, V, col0, 0);
      tma_store_commit();
      if (p.stages == 1) {
        // a single stage: the next tile can only be requested once the stores above have finished reading it
        if (i + 1 < my_tiles) {
          tma_store_wait_read();
          issue_load(i + 1);
        }
      } else if (i >= 1 && i - 1 + p.stages < my_tiles) {
        // refill the stage of the PREVIOUS tile once its stores have finished reading it (one store group may stay in
        // flight, so this never waits for the group committed just above); tile i + 1 is already on its way
        bulk_wait_read_le1();
        issue_load(i - 1 + p.stages);
      }
    }
  }
  if (c == 0) tma_store_wait_read();  // shared memory must outlive the reads of the last stores
}

static inline uint32_t gae_tma_stage_bytes(int T, int C) {
  const uint32_t b = (uint32_t)T * (uint32_t)C * 13u;
  return (b + 127u) & ~127u;
}

int launch_gae_tma(const GaeParams& p, const GaeTmaConfig& cfg, cudaStream_t s) {
  // layout requirements of the tensor maps: 16-byte aligned bases and row pitches (N*4 and N bytes), Dv == 1
  if (p.Dv != 1 || (p.N % 16) != 0 || p.T > 256 || p.N >= (1ll << 31)) return CUSRL_B200_EUNSUPPORTED;
  if (!aligned_to(p.reward, 16) || !aligned_to(p.value, 16) || !aligned_to(p.next_value, 16) || !aligned_to(p.done, 16) ||
      !aligned_to(p.advantage, 16) || (p.ret && !aligned_to(p.ret, 16)))
    return CUSRL_B200_EUNSUPPORTED;
  const int T = (int)p.T;
  const int64_t N = p.N;
  const int k = cfg.ctas_per_sm < 1 ? 1 : (cfg.ctas_per_sm > 8 ? 8 : cfg.ctas_per_sm);
  const int64_t ctas_max = (int64_t)sm_count() * k;
  const int64_t budget = (228 * 1024) / k - 1024 - kGaeTmaHeader;  // shared memory per CTA for k resident CTAs
  const int want_stages = cfg.stages < 1 ? 1 : (cfg.stages > kGaeTmaMaxStages ? kGaeTmaMaxStages : cfg.stages);

  // tile width: the columns an SM walks through one after the other are waves * W (per resident CTA); take the W that
  // minimises it (N = 65536 on 148 SMs: W = 7 -> 293 tiles, 99 % of the grid busy; W = 8 -> 256 tiles, 86 %)
  int W = 0, stages = 0;
  int64_t best_cost = 0;
  for (int w = (cfg.warps > 0 ? cfg.warps : 8); w >= (cfg.warps > 0 ? cfg.warps : 1); --w) {
    if (w > 8) continue;  // box inner dimension <= 256 elements
    const int64_t tiles = (N + 32 * w - 1) / (32 * w);
    const int64_t waves = (tiles + ctas_max - 1) / ctas_max;
    const uint32_t sb = gae_tma_stage_bytes(T, 32 * w);
    int st = (int)(waves < want_stages ? waves : want_stages);
    while (st > 0 && (int64_t)sb * st > budget) --st;
    if (st == 0) continue;
    const int64_t cost = waves * w;
    if (W == 0 || cost < best_cost) W = w, stages = st, best_cost = cost;
  }
  if (W == 0) return CUSRL_B200_EUNSUPPORTED;
  const int C = 32 * W;
  const int64_t tiles = (N + C - 1) / C;

  CUtensorMap tR, tV, tNV, tD, tA, tRet;
  const uint64_t uN = (uint64_t)N, uT = (uint64_t)T;
  if (int e = encode_tmap_2d_plain(&tR, p.reward, 4, uN, uT, uN * 4, (uint32_t)C, (uint32_t)T)) return e;
  if (int e = encode_tmap_2d_plain(&tV, p.value, 4, uN, uT, uN * 4, (uint32_t)C, (uint32_t)T)) return e;
  if (int e = encode_tmap_2d_plain(&tNV, p.next_value, 4, uN, uT, uN * 4, (uint32_t)C, (uint32_t)T)) return e;
  if (int e = encode_tmap_2d_plain(&tD, p.done, 1, uN, uT, uN, (uint32_t)C, (uint32_t)T)) return e;
  if (int e = encode_tmap_2d_plain(&tA, p.advantage, 4, uN, uT, uN * 4, (uint32_t)C, (uint32_t)T)) return e;
  if (int e = encode_tmap_2d_plain(&tRet, p.ret ? p.ret : p.advantage, 4, uN, uT, uN * 4, (uint32_t)C, (uint32_t)T)) return e;

  GaeTmaParams q{};
  q.T = T, q.num_tiles = (int)tiles, q.stages = stages, q.stage_bytes = gae_tma_stage_bytes(T, C);
  q.gamma = p.gamma, q.c_adv = p.c_adv, q.c_ret = p.c_ret, q.two_lambda = p.two_lambda, q.has_ret = p.ret != nullptr;
  const size_t smem_bytes = kGaeTmaHeader + (size_t)q.stage_bytes * stages;

  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gae_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gae_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_last_error("gae_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const unsigned grid = (unsigned)(tiles < ctas_max ? tiles : ctas_max);
  gae_tma_kernel<<<grid, C, smem_bytes, s>>>(tR, tV, tNV, tD, tA, tRet, q);
  return check_launch("gae_tma_kernel");
}

}  // namespace cusrl_b200
