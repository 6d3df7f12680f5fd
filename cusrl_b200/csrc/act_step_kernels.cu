// Rollout-side kernels: what `ActorCritic.act` / `ActorCritic.step` do around the network
// forward passes, written straight into the time-major rollout buffer slots instead of through ~40 elementwise
// launches and 14 indexed copies per environment step.
//   reference: cusrl/template/actor_critic.py:227-291 (act / step), cusrl/nn/module/distribution.py:195-213
//              (Normal.rsample + log_prob), cusrl/template/buffer.py:124-151 (push: one indexed copy per leaf).
// All HBM-bound streaming kernels; rows are env instances, lanes run along the contiguous feature axis.
#include <math.h>

#include "common.cuh"

namespace cusrl_b200 {

// ------------------------------------------------------------------------------------------------
// rows of `width` floats, source pitch lds (any 4-byte aligned layout, e.g. a dense [N, 235] observation), destination
// pitch ldd >= width (the buffer slot, rows padded to 16 bytes): one warp per row, lanes stride the row, so both the
// 940-byte source rows and the 944-byte destination rows are accessed as contiguous 128-byte segments.  The padding
// columns of the destination are (re)written as zeros so the slot is a legal zero-padded TMA source.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_row_padded(const float* __restrict__ src, float* __restrict__ dst, int width, int ldd,
                                                int lane) {
  for (int c = lane; c < ldd; c += 32) dst[c] = c < width ? ldg_stream(src + c) : 0.f;
}

struct StoreStepParams {
  // wide leaves (observation-like): up to 2 per launch (next_observation, next_state)
  const float* wide_src[2];
  float* wide_dst[2];
  int64_t wide_lds[2], wide_ldd[2];
  int wide_width[2];
  int n_wide;
  // narrow leaves
  const float* reward_src;
  float* reward_dst;
  int reward_dim;
  const uint8_t *terminated_src, *truncated_src;
  uint8_t *terminated_dst, *truncated_dst, *done_dst;
  int64_t N;
};

// step(): next_observation / next_state rows, reward, terminated, truncated, done = terminated | truncated
// (actor_critic.py:255-291 + buffer.py:124-151) in ONE launch.
__global__ void __launch_bounds__(256) rollout_store_step_kernel(const __grid_constant__ StoreStepParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < p.N; row += nwarps) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (k < p.n_wide)
        copy_row_padded(p.wide_src[k] + row * p.wide_lds[k], p.wide_dst[k] + row * p.wide_ldd[k], p.wide_width[k],
                        (int)p.wide_ldd[k], lane);
    if (lane < p.reward_dim && p.reward_src) p.reward_dst[row * p.reward_dim + lane] = p.reward_src[row * p.reward_dim + lane];
    if (lane == 0 && p.terminated_src) {
      const uint8_t te = p.terminated_src[row] != 0, tr = p.truncated_src[row] != 0;
      p.terminated_dst[row] = te;
      p.truncated_dst[row] = tr;
      p.done_dst[row] = te | tr;
    }
  }
}

__global__ void __launch_bounds__(256) copy_rows_padded_kernel(const float* __restrict__ src, int64_t lds,
                                                               float* __restrict__ dst, int64_t ldd, int64_t rows, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < rows; row += nwarps) copy_row_padded(src + row * lds, dst + row * ldd, width, (int)ldd, lane);
}

// ------------------------------------------------------------------------------------------------
// act(): given the mean the head kernel already wrote into its buffer slot, the state-independent std vector and the
// standard-normal draw eps (torch's Philox stream, so the generator advances exactly as Normal.rsample's does):
//   std[n, :]    = sigma                                   (StddevVector.forward: param.repeat(N, 1))
//   action[n, :] = mean + eps * sigma                      (Normal.rsample: loc + eps * scale; deterministic: mean)
//   logp[n]      = sum_d( -((a - mu)^2) / (2 sigma^2) - log(sigma) - log(sqrt(2 pi)) )     (Normal.log_prob, summed)
// evaluated in torch's operation order with contraction disabled.  One thread per row; rows are <= 16 floats.
// ------------------------------------------------------------------------------------------------
template <int A_MAX>
__global__ void __launch_bounds__(256) sample_logp_kernel(const float* __restrict__ mean, const float* __restrict__ sigma,
                                                          const float* __restrict__ eps, int64_t N, int A, int deterministic,
                                                          float* __restrict__ std_out, float* __restrict__ action_out,
                                                          float* __restrict__ logp_out) {
  __shared__ float s_sigma[A_MAX], s_two_var[A_MAX], s_log[A_MAX];
  if (threadIdx.x < A) {
    const float s = sigma[threadIdx.x];
    s_sigma[threadIdx.x] = s;
    s_two_var[threadIdx.x] = __fmul_rn(2.f, __fmul_rn(s, s));
    s_log[threadIdx.x] = logf(s);
  }
  __syncthreads();
  const float log_sqrt_2pi = 0.91893853320467274178f;  // math.log(math.sqrt(2 * math.pi))
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += stride) {
    float lp = 0.f;
    for (int d = 0; d < A; ++d) {
      const float mu = mean[n * A + d];
      const float a = deterministic ? mu : __fadd_rn(mu, __fmul_rn(eps[n * A + d], s_sigma[d]));
      const float diff = __fsub_rn(a, mu);
      float t = __fdiv_rn(-__fmul_rn(diff, diff), s_two_var[d]);
      t = __fsub_rn(__fsub_rn(t, s_log[d]), log_sqrt_2pi);
      lp = __fadd_rn(lp, t);
      std_out[n * A + d] = s_sigma[d];
      action_out[n * A + d] = a;
    }
    logp_out[n] = lp;
  }
}


// ------------------------------------------------------------------------------------------------
// step() of a recurrent agent: `memory[done] = 0` for up to four [N, W] memory tensors (Module.reset_memory,
// nn/module/module.py + hook/on_policy/value.py:56 for the critic's) and, in the same pass, the copies of the reset memory
// the rollout buffer needs (actor_memory / critic_memory of the NEXT step, next_critic_memory of this one): up to two extra
// destinations per tensor.  Replaces four masked fills and up to six copies per environment step.
// ------------------------------------------------------------------------------------------------
struct MemoryResetParams {
  float* mem[4];
  float* dst_a[4];
  float* dst_b[4];
  const uint8_t* done;
  int64_t N;
  int width, count;   // W (a multiple of 4), number of tensors
};

__global__ void __launch_bounds__(256) memory_reset_store_kernel(const MemoryResetParams p) {
  const int w4 = p.width >> 2;
  const int64_t total = p.N * w4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / w4;
    const bool reset = p.done[n] != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= p.count) break;
      float4* src = reinterpret_cast<float4*>(p.mem[k]) + i;
      float4 v = *src;
      if (reset) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        *src = v;
      }
      if (p.dst_a[k]) reinterpret_cast<float4*>(p.dst_a[k])[i] = v;
      if (p.dst_b[k]) reinterpret_cast<float4*>(p.dst_b[k])[i] = v;
    }
  }
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_copy_rows_padded_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t width,
                                    void* stream) {
  CUSRL_REQUIRE(src && dst, CUSRL_B200_EINVAL, "copy_rows_padded: null pointer");
  CUSRL_REQUIRE(rows > 0 && width > 0 && lds >= width && ldd >= width && ldd < (1ll << 30), CUSRL_B200_EINVAL,
                "copy_rows_padded: bad sizes");
  int64_t blocks = (rows + 7) / 8;  // 8 warps per block, one row per warp iteration
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  copy_rows_padded_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, (int)width);
  return check_launch("copy_rows_padded_kernel");
}

int cusrl_b200_rollout_store_step_f32(const float* next_obs, int64_t ld_next_obs, float* next_obs_slot, int64_t ld_next_obs_slot,
                                      int64_t obs_dim, const float* next_state, int64_t ld_next_state, float* next_state_slot,
                                      int64_t ld_next_state_slot, int64_t state_dim, const float* reward, float* reward_slot,
                                      int64_t reward_dim, const uint8_t* terminated, const uint8_t* truncated,
                                      uint8_t* terminated_slot, uint8_t* truncated_slot, uint8_t* done_slot, int64_t N,
                                      void* stream) {
  CUSRL_REQUIRE(N > 0, CUSRL_B200_EINVAL, "rollout_store_step: N must be positive");
  CUSRL_REQUIRE(!next_obs || (next_obs_slot && obs_dim > 0 && ld_next_obs >= obs_dim && ld_next_obs_slot >= obs_dim),
                CUSRL_B200_EINVAL, "rollout_store_step: bad next_observation arguments");
  CUSRL_REQUIRE(!next_state || (next_state_slot && state_dim > 0 && ld_next_state >= state_dim && ld_next_state_slot >= state_dim),
                CUSRL_B200_EINVAL, "rollout_store_step: bad next_state arguments");
  CUSRL_REQUIRE(!reward || (reward_slot && reward_dim > 0 && reward_dim <= 32), CUSRL_B200_EINVAL,
                "rollout_store_step: reward_dim must be in 1..32");
  CUSRL_REQUIRE(!terminated || (truncated && terminated_slot && truncated_slot && done_slot), CUSRL_B200_EINVAL,
                "rollout_store_step: terminated / truncated / done come together");
  StoreStepParams p{};
  if (next_obs) {
    p.wide_src[p.n_wide] = next_obs, p.wide_dst[p.n_wide] = next_obs_slot, p.wide_lds[p.n_wide] = ld_next_obs;
    p.wide_ldd[p.n_wide] = ld_next_obs_slot, p.wide_width[p.n_wide] = (int)obs_dim, ++p.n_wide;
  }
  if (next_state) {
    p.wide_src[p.n_wide] = next_state, p.wide_dst[p.n_wide] = next_state_slot, p.wide_lds[p.n_wide] = ld_next_state;
    p.wide_ldd[p.n_wide] = ld_next_state_slot, p.wide_width[p.n_wide] = (int)state_dim, ++p.n_wide;
  }
  p.reward_src = reward, p.reward_dst = reward_slot, p.reward_dim = (int)reward_dim;
  p.terminated_src = terminated, p.truncated_src = truncated;
  p.terminated_dst = terminated_slot, p.truncated_dst = truncated_slot, p.done_dst = done_slot;
  p.N = N;
  int64_t blocks = (N + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rollout_store_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("rollout_store_step_kernel");
}

int cusrl_b200_sample_logp_f32(const float* mean, const float* sigma, const float* eps, int64_t N, int64_t A,
                               int deterministic, float* std_out, float* action_out, float* logp_out, void* stream) {
  CUSRL_REQUIRE(mean && sigma && std_out && action_out && logp_out && (eps || deterministic), CUSRL_B200_EINVAL,
                "sample_logp: null pointer");
  CUSRL_REQUIRE(N > 0 && A > 0 && A <= 64, CUSRL_B200_EUNSUPPORTED, "sample_logp: action dim must be in 1..64");
  int64_t blocks = (N + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sample_logp_kernel<64><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mean, sigma, eps, N, (int)A, deterministic, std_out,
                                                                             action_out, logp_out);
  return check_launch("sample_logp_kernel");
}

int cusrl_b200_memory_reset_store_f32(float* const* mem, float* const* dst_a, float* const* dst_b, int64_t count, const uint8_t* done,
                                      int64_t N, int64_t width, void* stream) {
  CUSRL_REQUIRE(mem && done && count >= 1 && count <= 4, CUSRL_B200_EINVAL, "memory_reset_store: 1..4 tensors");
  CUSRL_REQUIRE(N > 0 && width > 0 && (width % 4) == 0 && width < (1 << 20), CUSRL_B200_EINVAL,
                "memory_reset_store: width must be a positive multiple of 4");
  MemoryResetParams p{};
  for (int k = 0; k < (int)count; ++k) {
    p.mem[k] = mem[k];
    p.dst_a[k] = dst_a ? dst_a[k] : nullptr;
    p.dst_b[k] = dst_b ? dst_b[k] : nullptr;
    CUSRL_REQUIRE(p.mem[k] && aligned_to(p.mem[k], 16) && (!p.dst_a[k] || aligned_to(p.dst_a[k], 16)) &&
                      (!p.dst_b[k] || aligned_to(p.dst_b[k], 16)),
                  CUSRL_B200_EALIGN, "memory_reset_store: tensors must be dense, 16-byte aligned");
  }
  p.done = done, p.N = N, p.width = (int)width, p.count = (int)count;
  int64_t blocks = (N * (width / 4) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  memory_reset_store_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("memory_reset_store_kernel");
}

}  // extern "C"
