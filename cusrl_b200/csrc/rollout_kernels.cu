// K3 next_value, K1 GAE scan (+fused K3), K2 advantage statistics / normalisation.
// All HBM-bound: lanes run along the contiguous env axis N of the time-major [T,N,Dv] leaves
// (reference layout: template/buffer.py:144), the T-step recurrence lives in registers.
#include "common.cuh"
#include "gae_common.cuh"

namespace cusrl_b200 {

// ------------------------------------------------------------------------------------------------
// K3: hook/on_policy/value.py:68-82
// ------------------------------------------------------------------------------------------------
__global__ void next_value_kernel(const float* __restrict__ value, const uint8_t* __restrict__ terminated,
                                  const uint8_t* __restrict__ truncated, const float* __restrict__ boot,
                                  const float* __restrict__ trunc_value, float* __restrict__ next_value,
                                  int64_t T, int64_t N, int64_t Dv, float termination_value) {
  const int64_t C = N * Dv;
  const int64_t total = T * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / C, c = i - t * C, n = c / Dv;
    float nv = (t + 1 < T) ? value[i + C] : boot[c];
    if (terminated[t * N + n]) nv = termination_value;
    if (truncated[t * N + n]) nv = trunc_value ? trunc_value[i] : value[i];
    next_value[i] = nv;
  }
}

// ------------------------------------------------------------------------------------------------
// K1: hook/on_policy/gae.py:8-20,85-110 (+ value.py:68-82 when FUSED)
// One thread owns VEC adjacent columns; the backward-in-time loop is processed in chunks of U steps
// whose loads are all issued before the dependent arithmetic (memory-level parallelism), the
// recurrence itself is evaluated strictly in the reference's order with contraction disabled.
// ------------------------------------------------------------------------------------------------
template <int VEC>
struct VecLoad;
template <>
struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[1]) { o[0] = ldg_stream(p); }
  static __device__ __forceinline__ void ldf(const uint8_t* p, uint8_t (&o)[1]) { o[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { p[0] = v[0]; }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[2]) {
    float2 v = ldg_stream2(p);
    o[0] = v.x, o[1] = v.y;
  }
  static __device__ __forceinline__ void ldf(const uint8_t* p, uint8_t (&o)[2]) {
    uchar2 v = __ldg(reinterpret_cast<const uchar2*>(p));
    o[0] = v.x, o[1] = v.y;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    float4 v = ldg_stream4(p);
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  }
  static __device__ __forceinline__ void ldf(const uint8_t* p, uint8_t (&o)[4]) {
    uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <int VEC, int U, bool FUSED, typename idx_t>
__global__ void __launch_bounds__(128, (U * VEC >= 40 ? 2 : (U * VEC >= 16 ? 4 : 6))) gae_kernel(const GaeParams p) {
  const idx_t C = (idx_t)(p.N * p.Dv);
  const idx_t N = (idx_t)p.N;
  const idx_t c0 = (idx_t)(blockIdx.x * (idx_t)blockDim.x + threadIdx.x) * VEC;
  if (c0 >= C) return;
  // flag column: done broadcasts over Dv (gae.py:19); VEC > 1 is only launched with Dv == 1
  const idx_t n0 = (VEC == 1) ? (idx_t)(c0 / (idx_t)p.Dv) : c0;
  const int T = (int)p.T;

  float adv_next[VEC], adv2_next[VEC], v_next[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) adv_next[k] = 0.f, adv2_next[k] = 0.f, v_next[k] = 0.f;
  if (FUSED) VecLoad<VEC>::ld(p.boot + c0, v_next);

  for (int t_hi = T; t_hi > 0; t_hi -= U) {
    float r[U][VEC], v[U][VEC], nv[U][VEC];
    uint8_t f0[U][VEC], f1[U][VEC];
    // ---- issue every load of the chunk first
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t_hi - 1 - u;
      if (t >= 0) {
        const idx_t off = (idx_t)t * C + c0;
        const idx_t foff = (idx_t)t * N + n0;
        VecLoad<VEC>::ld(p.reward + off, r[u]);
        VecLoad<VEC>::ld(p.value + off, v[u]);
        if (!FUSED) VecLoad<VEC>::ld(p.next_value + off, nv[u]);
        VecLoad<VEC>::ldf(p.done + foff, f0[u]);
        if (FUSED) VecLoad<VEC>::ldf(p.truncated + foff, f1[u]);
      }
    }
    // ---- then the sequential recurrence
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t_hi - 1 - u;
      if (t >= 0) {
        float adv[VEC], rt[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          bool done;
          float nvk;
          if (FUSED) {
            // value.py:68-82: shift / bootstrap, then terminated, then truncated (IsaacLab branch)
            nvk = v_next[k];
            if (f0[u][k]) nvk = p.termination_value;
            if (f1[u][k]) nvk = v[u][k];
            done = (f0[u][k] | f1[u][k]) != 0;  // actor_critic.py:277
            nv[u][k] = nvk;
          } else {
            nvk = nv[u][k];
            done = f0[u][k] != 0;
          }
          // gae.py:17  advantage = reward + next_value * gamma - value
          const float delta = __fsub_rn(__fadd_rn(r[u][k], __fmul_rn(nvk, p.gamma)), v[u][k]);
          float a = delta, a2 = delta;
          if (t != T - 1) {
            // gae.py:19  advantage[t] += not_done[t] * (gamma*lamda) * advantage[t+1]
            const float m = done ? 0.f : p.c_adv;
            a = __fadd_rn(delta, __fmul_rn(m, adv_next[k]));
            if (p.two_lambda) {
              const float m2 = done ? 0.f : p.c_ret;
              a2 = __fadd_rn(delta, __fmul_rn(m2, adv2_next[k]));
            }
          }
          adv_next[k] = a;
          adv2_next[k] = a2;
          v_next[k] = v[u][k];
          adv[k] = a;
          // gae.py:99-110  return = value + advantage (or the lamda_value scan)
          rt[k] = __fadd_rn(v[u][k], p.two_lambda ? a2 : a);
        }
        const idx_t off = (idx_t)t * C + c0;
        VecLoad<VEC>::st(p.advantage + off, adv);
        if (p.ret) VecLoad<VEC>::st(p.ret + off, rt);
        if (FUSED && p.next_value_out) VecLoad<VEC>::st(p.next_value_out + off, nv[u]);
      }
    }
  }
}

// Exact-length variant (schedule 1): T is a template parameter, so no step is predicated, the addresses are a base
// pointer plus t * pitch, and the loads are issued in ASCENDING time order while the recurrence starts at t = T-1 --
// its first instruction depends on the loads issued LAST, so the compiler cannot hoist arithmetic into the load
// sequence.  (ncu on the chunked kernel above: ptxas placed the first FMUL of step T-1 after 24 of the 96 loads; the warp
// then sat on the long scoreboard with three quarters of its requests not yet issued.)  MEASURED: 9.5 us per launch at
// 65536 x 24 against 9.2-9.3 us for the chunked kernel -- with 14 warps per SM the other warps cover that stall, and the
// launch is bounded by its fixed ramp / drain cost (profiles/r01_gae_ncu.md).  Kept selectable, not the default.
// 64 threads x 7 blocks = 448 threads per SM hold all 65536 columns of the BASELINE rollout in one wave on 148 SMs
// and leave 144 registers per thread (the 96 loaded values + addresses fit without spilling).
template <int T_>
__global__ void __launch_bounds__(64, 7) gae_exact_kernel(const GaeParams p) {
  const int64_t C = p.N * p.Dv;
  const int64_t c0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c0 >= C) return;
  const int64_t n0 = p.Dv == 1 ? c0 : c0 / p.Dv;  // done broadcasts over Dv (gae.py:19)
  const float* __restrict__ rp = p.reward + c0;
  const float* __restrict__ vp = p.value + c0;
  const float* __restrict__ np = p.next_value + c0;
  const uint8_t* __restrict__ dp = p.done + n0;
  float r[T_], v[T_], nv[T_];
  uint8_t d[T_];
#pragma unroll
  for (int t = 0; t < T_; ++t) {
    r[t] = ldg_stream(rp + t * C);
    v[t] = ldg_stream(vp + t * C);
    nv[t] = ldg_stream(np + t * C);
  }
#pragma unroll
  for (int t = 0; t < T_; ++t) d[t] = __ldg(dp + t * p.N);
  // the flags are the last requests issued; folding them into one mask frees T-1 registers for the recurrence
  uint32_t done_mask = 0;
#pragma unroll
  for (int t = 0; t < T_; ++t) done_mask |= (d[t] ? 1u : 0u) << t;
  float* __restrict__ ap = p.advantage + c0;
  float* __restrict__ tp = p.ret ? p.ret + c0 : nullptr;
  float adv_next = 0.f, adv2_next = 0.f;
#pragma unroll
  for (int t = T_ - 1; t >= 0; --t) {
    // gae.py:17  advantage = reward + next_value * gamma - value
    const float delta = __fsub_rn(__fadd_rn(r[t], __fmul_rn(nv[t], p.gamma)), v[t]);
    float a = delta, a2 = delta;
    if (t != T_ - 1) {
      // gae.py:19  advantage[t] += not_done[t] * (gamma*lamda) * advantage[t+1]
      const bool done = (done_mask >> t) & 1u;
      a = __fadd_rn(delta, __fmul_rn(done ? 0.f : p.c_adv, adv_next));
      if (p.two_lambda) a2 = __fadd_rn(delta, __fmul_rn(done ? 0.f : p.c_ret, adv2_next));
    }
    adv_next = a, adv2_next = a2;
    ap[t * C] = a;
    // gae.py:99-110  return = value + advantage (or the lamda_value scan)
    if (tp) tp[t * C] = __fadd_rn(v[t], p.two_lambda ? a2 : a);
  }
}

static int g_gae_schedule = 0;  // 0: chunked kernel above, 1: exact-length kernel when T has an instantiation

template <int T_>
static void launch_gae_exact(const GaeParams& p, int threads, cudaStream_t s) {
  const int64_t C = p.N * p.Dv;
  if (threads > 64) threads = 64;  // the kernel's launch bound
  gae_exact_kernel<T_><<<(unsigned)((C + threads - 1) / threads), threads, 0, s>>>(p);
}

// returns false when T has no exact-length instantiation
static bool try_launch_gae_exact(const GaeParams& p, int threads, cudaStream_t s) {
  switch (p.T) {
    case 8: launch_gae_exact<8>(p, threads, s); return true;
    case 12: launch_gae_exact<12>(p, threads, s); return true;
    case 16: launch_gae_exact<16>(p, threads, s); return true;
    case 24: launch_gae_exact<24>(p, threads, s); return true;
    case 32: launch_gae_exact<32>(p, threads, s); return true;
    default: return false;
  }
}

static int g_gae_vec = 1, g_gae_threads = 64;  // tuning knobs, see cusrl_b200_gae_set_config
static int g_gae_variant = 0;                  // 0: register-resident LDG kernel, 1: TMA-staged kernel (gae_tma.cu)
static GaeTmaConfig g_gae_tma_cfg = {0, 2, 2};  // see cusrl_b200_gae_set_variant

template <int VEC, bool FUSED, typename idx_t>
static void launch_gae_u(const GaeParams& p, unsigned grid, int threads, cudaStream_t s) {
  // U >= T puts the whole rollout of a column in registers: every load of the kernel is issued
  // before the first dependent instruction (one DRAM round trip instead of T/U of them).
  if constexpr (VEC == 1) {
    if (p.T <= 8)
      gae_kernel<1, 8, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
    else if (p.T <= 16)
      gae_kernel<1, 16, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
    else if (p.T <= 24)
      gae_kernel<1, 24, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
    else if (p.T <= 32)
      gae_kernel<1, 32, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
    else
      gae_kernel<1, 16, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
  } else if constexpr (VEC == 2) {
    if (p.T > 12 && p.T <= 24 && !FUSED)
      gae_kernel<2, 24, FUSED, idx_t><<<grid, threads, 0, s>>>(p);  // whole rollout of two columns in registers
    else if (p.T <= 12 || p.T == 24)
      gae_kernel<2, 12, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
    else
      gae_kernel<2, 8, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
  } else {
    gae_kernel<4, 6, FUSED, idx_t><<<grid, threads, 0, s>>>(p);
  }
}

template <bool FUSED>
static int launch_gae(const GaeParams& p, cudaStream_t s) {
  const int64_t C = p.N * p.Dv;
  if (!FUSED && g_gae_variant == 1) {
    const int rc = launch_gae_tma(p, g_gae_tma_cfg, s);
    if (rc != CUSRL_B200_EUNSUPPORTED) return rc;  // otherwise: layout not TMA-able, use the LDG kernel below
  }
  if (!FUSED && g_gae_schedule == 1 && g_gae_vec == 1 && try_launch_gae_exact(p, g_gae_threads, s))
    return check_launch("gae_exact_kernel");
  auto ok = [&](int vec) {
    if (p.Dv != 1 || (p.N % vec) != 0) return false;
    const size_t a = 4 * vec;
    bool al = aligned_to(p.reward, a) && aligned_to(p.value, a) && aligned_to(p.advantage, a) &&
              aligned_to(p.done, vec) && (!p.ret || aligned_to(p.ret, a));
    if (FUSED)
      al = al && aligned_to(p.boot, a) && aligned_to(p.truncated, vec) &&
           (!p.next_value_out || aligned_to(p.next_value_out, a));
    else
      al = al && aligned_to(p.next_value, a);
    return al;
  };
  int vec = g_gae_vec;
  while (vec > 1 && !ok(vec)) vec >>= 1;
  const int threads = g_gae_threads;
  const int64_t nthreads = (C + vec - 1) / vec;
  const unsigned grid = (unsigned)((nthreads + threads - 1) / threads);
  // 64-bit indexing everywhere: ptxas strength-reduces it to pointer increments (fewer live
  // registers than 32-bit offsets, checked with -Xptxas -v)
#define CUSRL_GAE_DISPATCH(V) launch_gae_u<V, FUSED, int64_t>(p, grid, threads, s);
  if (vec == 4) {
    CUSRL_GAE_DISPATCH(4)
  } else if (vec == 2) {
    CUSRL_GAE_DISPATCH(2)
  } else {
    CUSRL_GAE_DISPATCH(1)
  }
#undef CUSRL_GAE_DISPATCH
  return check_launch("gae_kernel");
}

// ------------------------------------------------------------------------------------------------
// The whole pre-update chain of the PPO preset in ONE launch (Dv == 1): K3 next_value (value.py:68-82) formed on the fly
// and published, K1 advantage / return scan (gae.py:8-20,85-110), and the block partial sums K2 needs for the advantage
// mean / variance (advantage.py:110-111) -- the three reference stages re-read the same [T, N] leaves three times.
// One thread per env column, the whole rollout of the column in registers (exact-length schedule: loads ascending in
// time, recurrence descending), THREADS a compile-time block size so that 65536 columns are one wave of one CTA per SM.
// Arithmetic is the FUSED branch of gae_kernel above, same op order, contraction disabled: bit-identical results.
// partials: [gridDim.x][2] doubles (sum | sum of squares of the advantages of the block's columns).
// ------------------------------------------------------------------------------------------------
template <int T_, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS >= 384 ? 1 : (THREADS >= 192 ? 2 : 4))) gae_chain_kernel(const GaeParams p, double* __restrict__ partials) {
  __shared__ double smem[2 * 32];
  const int64_t N = p.N;
  const int64_t c0 = blockIdx.x * (int64_t)THREADS + threadIdx.x;
  const bool live = c0 < N;
  const int64_t c = live ? c0 : 0;  // dead threads of the last block read column 0 and store nothing
  float r[T_], v[T_];
  uint8_t te[T_], tr[T_];
#pragma unroll
  for (int t = 0; t < T_; ++t) {
    r[t] = ldg_stream(p.reward + t * N + c);
    v[t] = ldg_stream(p.value + t * N + c);
  }
#pragma unroll
  for (int t = 0; t < T_; ++t) {
    te[t] = __ldg(p.done + t * N + c);
    tr[t] = __ldg(p.truncated + t * N + c);
  }
  float v_next = ldg_stream(p.boot + c);
  uint32_t term_mask = 0, trunc_mask = 0;
#pragma unroll
  for (int t = 0; t < T_; ++t) term_mask |= (te[t] ? 1u : 0u) << t, trunc_mask |= (tr[t] ? 1u : 0u) << t;
  float adv_next = 0.f, adv2_next = 0.f, s = 0.f, q = 0.f;
#pragma unroll
  for (int t = T_ - 1; t >= 0; --t) {
    // value.py:68-82: shift / bootstrap, then terminated, then truncated (final_state_is_missing branch)
    float nv = v_next;
    const bool is_term = (term_mask >> t) & 1u, is_trunc = (trunc_mask >> t) & 1u;
    if (is_term) nv = p.termination_value;
    if (is_trunc) nv = v[t];
    const bool done = is_term || is_trunc;  // actor_critic.py:277
    // gae.py:17  advantage = reward + next_value * gamma - value
    const float delta = __fsub_rn(__fadd_rn(r[t], __fmul_rn(nv, p.gamma)), v[t]);
    float a = delta, a2 = delta;
    if (t != T_ - 1) {
      // gae.py:19  advantage[t] += not_done[t] * (gamma*lamda) * advantage[t+1]
      a = __fadd_rn(delta, __fmul_rn(done ? 0.f : p.c_adv, adv_next));
      if (p.two_lambda) a2 = __fadd_rn(delta, __fmul_rn(done ? 0.f : p.c_ret, adv2_next));
    }
    adv_next = a, adv2_next = a2, v_next = v[t];
    s += a, q += a * a;
    if (live) {
      p.advantage[t * N + c0] = a;
      if (p.ret) p.ret[t * N + c0] = __fadd_rn(v[t], p.two_lambda ? a2 : a);  // gae.py:99-110
      if (p.next_value_out) p.next_value_out[t * N + c0] = nv;
    }
  }
  double acc[2] = {live ? (double)s : 0.0, live ? (double)q : 0.0};
  block_sum<2>(acc, smem);
  if (threadIdx.x == 0) partials[2 * blockIdx.x + 0] = acc[0], partials[2 * blockIdx.x + 1] = acc[1];
}

static int g_gae_chain_threads = 448;  // 65536 columns = 147 CTAs of 448 threads: one wave, one CTA per SM (gae_set_chain_threads)

template <int T_>
static unsigned launch_gae_chain_t(const GaeParams& p, double* partials, cudaStream_t s) {
  const int64_t N = p.N;
  switch (g_gae_chain_threads) {
    case 128: { const unsigned g = (unsigned)((N + 127) / 128); gae_chain_kernel<T_, 128><<<g, 128, 0, s>>>(p, partials); return g; }
    case 256: { const unsigned g = (unsigned)((N + 255) / 256); gae_chain_kernel<T_, 256><<<g, 256, 0, s>>>(p, partials); return g; }
    default: { const unsigned g = (unsigned)((N + 447) / 448); gae_chain_kernel<T_, 448><<<g, 448, 0, s>>>(p, partials); return g; }
  }
}

// returns the number of blocks launched (= partial pairs written), 0 when T has no instantiation
static unsigned launch_gae_chain(const GaeParams& p, double* partials, cudaStream_t s) {
  switch (p.T) {
    case 8: return launch_gae_chain_t<8>(p, partials, s);
    case 12: return launch_gae_chain_t<12>(p, partials, s);
    case 16: return launch_gae_chain_t<16>(p, partials, s);
    case 24: return launch_gae_chain_t<24>(p, partials, s);
    case 32: return launch_gae_chain_t<32>(p, partials, s);
    default: return 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: hook/on_policy/advantage.py:108-115
// ------------------------------------------------------------------------------------------------
constexpr int kStatsThreads = 256;
constexpr int kStatsMaxBlocks = 592;  // 4 x 148 SMs
constexpr int kStatsMaxDv = 64;

// partials layout: [block][2*Dv] doubles (sum | sumsq)
__global__ void __launch_bounds__(kStatsThreads) adv_stats_partial_dv1(const float* __restrict__ adv, int64_t E,
                                                                       double* __restrict__ partials) {
  __shared__ double smem[2 * 32];
  double acc[2] = {0.0, 0.0};
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t E4 = aligned_to(adv, 16) ? (E >> 2) : 0;
  float s = 0.f, q = 0.f;
  int cnt = 0;
  for (int64_t i = tid; i < E4; i += stride) {
    const float4 v = ldg_stream4(adv + 4 * i);
    s += (v.x + v.y) + (v.z + v.w);
    q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (++cnt == 8) {  // spill the fp32 running sums into double every 32 elements
      acc[0] += s, acc[1] += q, s = 0.f, q = 0.f, cnt = 0;
    }
  }
  for (int64_t i = 4 * E4 + tid; i < E; i += stride) {
    const float v = adv[i];
    s += v, q += v * v;
  }
  acc[0] += s, acc[1] += q;
  block_sum<2>(acc, smem);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x + 0] = acc[0];
    partials[2 * blockIdx.x + 1] = acc[1];
  }
}

__global__ void __launch_bounds__(kStatsThreads) adv_stats_partial_generic(const float* __restrict__ adv, int64_t E,
                                                                           int Dv, double* __restrict__ partials) {
  __shared__ double sh[2 * kStatsMaxDv];
  for (int i = threadIdx.x; i < 2 * Dv; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int64_t total = E * Dv;
  // stride is a multiple of Dv so a thread always sees one channel
  const int64_t raw = (int64_t)gridDim.x * blockDim.x;
  const int64_t stride = ((raw + Dv - 1) / Dv) * Dv;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int64_t i = tid; i < total; i += stride) {
    const double v = adv[i];
    s += v, q += v * v;
  }
  if (tid < total) {
    const int ch = (int)(tid % Dv);
    atomicAdd(&sh[ch], s);
    atomicAdd(&sh[Dv + ch], q);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Dv; i += blockDim.x) partials[(int64_t)blockIdx.x * 2 * Dv + i] = sh[i];
}

// Note: the generic path uses shared-memory double atomics, so its summation order within a block is
// not fixed; results agree to ~1e-16 relative, far inside the fp32 output precision.
__global__ void adv_stats_finalize(const double* __restrict__ partials, int nblocks, int Dv, int64_t E,
                                   float* __restrict__ mean_var) {
  const int ch = blockIdx.x;
  __shared__ double smem[2 * 32];
  double acc[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    acc[0] += partials[(int64_t)b * 2 * Dv + ch];
    acc[1] += partials[(int64_t)b * 2 * Dv + Dv + ch];
  }
  block_sum<2>(acc, smem);
  if (threadIdx.x == 0) {
    const double n = (double)E;
    const double mean = acc[0] / n;
    // torch.var_mean(correction=1): sum((x-mean)^2)/(n-1); n == 1 gives nan like torch
    const double var = (acc[1] - acc[0] * mean) / (n - 1.0);
    mean_var[ch] = (float)mean;
    mean_var[Dv + ch] = (float)(var < 0.0 ? 0.0 : var);
  }
}

__global__ void adv_normalize_kernel(float* __restrict__ adv, int64_t E, int Dv, const float* __restrict__ mean_var,
                                     float eps) {
  const int64_t total = E * Dv;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (Dv == 1) {
    const float mean = mean_var[0];
    const float sd = __fsqrt_rn(__fadd_rn(mean_var[1], eps));  // advantage.py:114
    const int64_t n4 = aligned_to(adv, 16) ? (total >> 2) : 0;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 v = *reinterpret_cast<float4*>(adv + 4 * i);
      v.x = __fdiv_rn(__fsub_rn(v.x, mean), sd);  // advantage.py:115 sub_ then div_
      v.y = __fdiv_rn(__fsub_rn(v.y, mean), sd);
      v.z = __fdiv_rn(__fsub_rn(v.z, mean), sd);
      v.w = __fdiv_rn(__fsub_rn(v.w, mean), sd);
      *reinterpret_cast<float4*>(adv + 4 * i) = v;
    }
    for (int64_t i = 4 * n4 + tid; i < total; i += stride) adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean), sd);
  } else {
    for (int64_t i = tid; i < total; i += stride) {
      const int ch = (int)(i % Dv);
      const float sd = __fsqrt_rn(__fadd_rn(mean_var[Dv + ch], eps));
      adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean_var[ch]), sd);
    }
  }
}

// utils/distributed.py:175-183
__global__ void merge_mean_var_kernel(const float* __restrict__ gathered, int W, int Dv, float* __restrict__ mean_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= Dv) return;
  float m = 0.f;
  for (int r = 0; r < W; ++r) m += gathered[(int64_t)r * 2 * Dv + ch];
  m = m / (float)W;
  float v = 0.f;
  for (int r = 0; r < W; ++r) {
    const float d = gathered[(int64_t)r * 2 * Dv + ch] - m;
    v += gathered[(int64_t)r * 2 * Dv + Dv + ch] + d * d;
  }
  mean_var[ch] = m;
  mean_var[Dv + ch] = v / (float)W;
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_gae_set_config(int vec, int threads) {
  if ((vec != 1 && vec != 2 && vec != 4) || threads < 32 || threads > 128 || (threads % 32)) return CUSRL_B200_EINVAL;
  g_gae_vec = vec;
  g_gae_threads = threads;
  return 0;
}

int cusrl_b200_gae_set_schedule(int schedule) {
  if (schedule != 0 && schedule != 1) return CUSRL_B200_EINVAL;
  g_gae_schedule = schedule;
  return 0;
}

int cusrl_b200_gae_set_variant(int variant, int warps, int stages, int ctas_per_sm) {
  if ((variant != 0 && variant != 1) || warps < 0 || warps > 8 || stages < 1 || stages > 8 || ctas_per_sm < 1 || ctas_per_sm > 8)
    return CUSRL_B200_EINVAL;
  g_gae_variant = variant;
  g_gae_tma_cfg = GaeTmaConfig{warps, stages, ctas_per_sm};
  return 0;
}

int cusrl_b200_next_value_f32(const float* value, const uint8_t* terminated, const uint8_t* truncated,
                              const float* boot_value, const float* trunc_value, float* next_value, int64_t T,
                              int64_t N, int64_t Dv, float termination_value, void* stream) {
  CUSRL_REQUIRE(value && terminated && truncated && boot_value && next_value, CUSRL_B200_EINVAL,
                "next_value: null pointer");
  CUSRL_REQUIRE(T > 0 && N > 0 && Dv > 0, CUSRL_B200_EINVAL, "next_value: T, N, Dv must be positive");
  const int64_t total = T * N * Dv;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  next_value_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      value, terminated, truncated, boot_value, trunc_value, next_value, T, N, Dv, termination_value);
  return check_launch("next_value_kernel");
}

static int gae_common_checks(int64_t T, int64_t N, int64_t Dv, double gamma, double lamda, double lamda_value) {
  CUSRL_REQUIRE(T > 0 && N > 0 && Dv > 0, CUSRL_B200_EINVAL, "gae: T, N, Dv must be positive");
  // same domain checks as GeneralizedAdvantageEstimation.__init__ (gae.py:59-64)
  CUSRL_REQUIRE(gamma >= 0 && gamma < 1, CUSRL_B200_EINVAL, "gae: 'gamma' must be in [0, 1); got %g", gamma);
  CUSRL_REQUIRE(lamda >= 0 && lamda <= 1, CUSRL_B200_EINVAL, "gae: 'lamda' must be in [0, 1]; got %g", lamda);
  CUSRL_REQUIRE(lamda_value <= 1, CUSRL_B200_EINVAL, "gae: 'lamda_value' must be in [0, 1]; got %g", lamda_value);
  return 0;
}

int cusrl_b200_gae_f32(const float* reward, const uint8_t* done, const float* value, const float* next_value,
                       float* advantage, float* ret, int64_t T, int64_t N, int64_t Dv, double gamma, double lamda,
                       double lamda_value, void* stream) {
  CUSRL_REQUIRE(reward && done && value && next_value && advantage, CUSRL_B200_EINVAL, "gae: null pointer");
  if (int e = gae_common_checks(T, N, Dv, gamma, lamda, lamda_value)) return e;
  GaeParams p{};
  p.reward = reward, p.done = done, p.value = value, p.next_value = next_value;
  p.advantage = advantage, p.ret = ret, p.T = T, p.N = N, p.Dv = Dv;
  p.gamma = (float)gamma;
  p.c_adv = (float)(gamma * lamda);  // python double product rounded once (gae.py:19)
  p.two_lambda = lamda_value >= 0;
  p.c_ret = p.two_lambda ? (float)(gamma * lamda_value) : p.c_adv;
  return launch_gae<false>(p, (cudaStream_t)stream);
}

int cusrl_b200_gae_fused_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated,
                             const float* value, const float* boot_value, float termination_value,
                             float* next_value_out, float* advantage, float* ret, int64_t T, int64_t N, int64_t Dv,
                             double gamma, double lamda, double lamda_value, void* stream) {
  CUSRL_REQUIRE(reward && terminated && truncated && value && boot_value && advantage, CUSRL_B200_EINVAL,
                "gae_fused: null pointer");
  if (int e = gae_common_checks(T, N, Dv, gamma, lamda, lamda_value)) return e;
  GaeParams p{};
  p.reward = reward, p.done = terminated, p.truncated = truncated, p.value = value, p.boot = boot_value;
  p.next_value_out = next_value_out, p.advantage = advantage, p.ret = ret, p.T = T, p.N = N, p.Dv = Dv;
  p.gamma = (float)gamma;
  p.c_adv = (float)(gamma * lamda);
  p.two_lambda = lamda_value >= 0;
  p.c_ret = p.two_lambda ? (float)(gamma * lamda_value) : p.c_adv;
  p.termination_value = termination_value;
  return launch_gae<true>(p, (cudaStream_t)stream);
}

int cusrl_b200_gae_set_chain_threads(int threads) {
  if (threads != 128 && threads != 256 && threads != 448) return CUSRL_B200_EINVAL;
  g_gae_chain_threads = threads;
  return 0;
}

size_t cusrl_b200_gae_chain_scratch_bytes(int64_t N) {
  if (N <= 0) return 0;
  return (size_t)((N + 127) / 128) * 2 * sizeof(double);
}

int cusrl_b200_gae_chain_supported(int64_t T, int64_t Dv) {
  return Dv == 1 && (T == 8 || T == 12 || T == 16 || T == 24 || T == 32);
}

int cusrl_b200_gae_chain_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated, const float* value,
                             const float* boot_value, float termination_value, float* next_value_out, float* advantage,
                             float* ret, int64_t T, int64_t N, double gamma, double lamda, double lamda_value,
                             float* mean_var, void* scratch, size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(reward && terminated && truncated && value && boot_value && advantage && mean_var && scratch,
                CUSRL_B200_EINVAL, "gae_chain: null pointer");
  if (int e = gae_common_checks(T, N, 1, gamma, lamda, lamda_value)) return e;
  CUSRL_REQUIRE(cusrl_b200_gae_chain_supported(T, 1), CUSRL_B200_EUNSUPPORTED,
                "gae_chain: rollout length %lld has no instantiation (8, 12, 16, 24, 32)", (long long)T);
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_gae_chain_scratch_bytes(N) && aligned_to(scratch, 8), CUSRL_B200_ESCRATCH,
                "gae_chain: scratch too small or misaligned");
  GaeParams p{};
  p.reward = reward, p.done = terminated, p.truncated = truncated, p.value = value, p.boot = boot_value;
  p.next_value_out = next_value_out, p.advantage = advantage, p.ret = ret, p.T = T, p.N = N, p.Dv = 1;
  p.gamma = (float)gamma;
  p.c_adv = (float)(gamma * lamda);  // python double product rounded once (gae.py:19)
  p.two_lambda = lamda_value >= 0;
  p.c_ret = p.two_lambda ? (float)(gamma * lamda_value) : p.c_adv;
  p.termination_value = termination_value;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = launch_gae_chain(p, (double*)scratch, s);
  if (int e = check_launch("gae_chain_kernel")) return e;
  adv_stats_finalize<<<1, 256, 0, s>>>((const double*)scratch, (int)blocks, 1, T * N, mean_var);
  return check_launch("adv_stats_finalize");
}

size_t cusrl_b200_advantage_stats_scratch_bytes(int64_t Dv) {
  if (Dv <= 0) return 0;
  return (size_t)kStatsMaxBlocks * 2 * (size_t)Dv * sizeof(double);
}

int cusrl_b200_advantage_stats_f32(const float* advantage, int64_t E, int64_t Dv, float* mean_var, void* scratch,
                                   size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(advantage && mean_var && scratch, CUSRL_B200_EINVAL, "advantage_stats: null pointer");
  CUSRL_REQUIRE(E > 0 && Dv > 0, CUSRL_B200_EINVAL, "advantage_stats: E, Dv must be positive");
  CUSRL_REQUIRE(Dv <= kStatsMaxDv, CUSRL_B200_EUNSUPPORTED, "advantage_stats: Dv > %d", kStatsMaxDv);
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_advantage_stats_scratch_bytes(Dv), CUSRL_B200_ESCRATCH,
                "advantage_stats: scratch too small");
  CUSRL_REQUIRE(aligned_to(scratch, 8), CUSRL_B200_EALIGN, "advantage_stats: scratch must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  double* partials = (double*)scratch;
  const int64_t work = (Dv == 1) ? (E + 3) / 4 : E * Dv;
  int64_t blocks = (work + kStatsThreads - 1) / kStatsThreads;
  int64_t cap = (int64_t)sm_count() * 4;
  if (cap > kStatsMaxBlocks) cap = kStatsMaxBlocks;
  if (blocks > cap) blocks = cap;
  if (Dv == 1)
    adv_stats_partial_dv1<<<(unsigned)blocks, kStatsThreads, 0, s>>>(advantage, E, partials);
  else
    adv_stats_partial_generic<<<(unsigned)blocks, kStatsThreads, 0, s>>>(advantage, E, (int)Dv, partials);
  if (int e = check_launch("adv_stats_partial")) return e;
  adv_stats_finalize<<<(unsigned)Dv, 256, 0, s>>>(partials, (int)blocks, (int)Dv, E, mean_var);
  return check_launch("adv_stats_finalize");
}

int cusrl_b200_advantage_normalize_f32(float* advantage, int64_t E, int64_t Dv, const float* mean_var, float eps,
                                       void* stream) {
  CUSRL_REQUIRE(advantage && mean_var, CUSRL_B200_EINVAL, "advantage_normalize: null pointer");
  CUSRL_REQUIRE(E > 0 && Dv > 0, CUSRL_B200_EINVAL, "advantage_normalize: E, Dv must be positive");
  const int threads = 256;
  const int64_t work = (Dv == 1) ? (E + 3) / 4 : E * Dv;
  int64_t blocks = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adv_normalize_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(advantage, E, (int)Dv, mean_var, eps);
  return check_launch("adv_normalize_kernel");
}

int cusrl_b200_merge_mean_var_f32(const float* gathered, int64_t W, int64_t Dv, float* mean_var, void* stream) {
  CUSRL_REQUIRE(gathered && mean_var, CUSRL_B200_EINVAL, "merge_mean_var: null pointer");
  CUSRL_REQUIRE(W > 0 && Dv > 0, CUSRL_B200_EINVAL, "merge_mean_var: W, Dv must be positive");
  merge_mean_var_kernel<<<(unsigned)((Dv + 63) / 64), 64, 0, (cudaStream_t)stream>>>(gathered, (int)W, (int)Dv,
                                                                                     mean_var);
  return check_launch("merge_mean_var_kernel");
}

}  // extern "C"
