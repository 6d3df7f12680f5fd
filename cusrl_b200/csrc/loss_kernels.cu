// K4: fused PPO objective (policy log-prob / entropy / ratio, clipped surrogate, entropy bonus, value
// loss) forward + unit gradients in one pass over the minibatch, and the policy statistics pass used
// by OnPolicyStatistics.  HBM-bound: 112 B read + 52 B written per sample at A = 12.
#include <math.h>

#include "common.cuh"

namespace cusrl_b200 {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1184;  // 8 x 148
constexpr int kMaxA = 32;
// per-block partial layout (doubles): [0] sum surrogate  [1] sum |logp_ratio|  [2] sum value-loss terms
//                                     [3] sum_B sum_Dv curr_value            [4 .. 4+A) sum_i g_i*dlogp/dstd
constexpr int kLossHead = 4;

template <int APAD, bool VEC4>
__device__ __forceinline__ void load_row(const float* __restrict__ base, int64_t row, int A, float (&out)[APAD]) {
  const float* p = base + row * A;
  if (VEC4) {
#pragma unroll
    for (int q = 0; q < APAD / 4; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
      out[4 * q + 0] = v.x, out[4 * q + 1] = v.y, out[4 * q + 2] = v.z, out[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < APAD; ++d) out[d] = d < A ? __ldg(p + d) : 0.f;
  }
}

struct LossParams {
  const float *mean, *std, *action, *logp_old, *advantage, *ret, *value_old, *curr_value;
  int64_t B;
  int A, Dv, has_value;
  float clip_ratio, w_s, w_e, w_v, value_clip;
  float *logp, *entropy, *logp_ratio, *prob_ratio;
  float *d_mean, *d_value;
  double* partials;
};

template <int APAD, bool VEC4>
__global__ void __launch_bounds__(kLossThreads) ppo_loss_kernel(const LossParams p) {
  __shared__ double smem[(kLossHead + kMaxA) * 32];
  const int A = p.A;
  // state-independent std (distribution.py:232-245): per-dimension constants, computed once per thread
  float sd[APAD], inv_var[APAD], inv_sd[APAD];
  float log_norm = 0.f;  // sum_d ( log(sd) + log(sqrt(2*pi)) )
  float ent = 0.f;       // sum_d ( 0.5 + 0.5*log(2*pi) + log(sd) )   (torch Normal.entropy)
  const float kLogSqrt2Pi = 0.91893853320467274178f;
#pragma unroll
  for (int d = 0; d < APAD; ++d) {
    sd[d] = d < A ? __ldg(p.std + d) : 1.f;
    const float var = sd[d] * sd[d];
    inv_var[d] = 1.f / var;
    inv_sd[d] = 1.f / sd[d];
    if (d < A) {
      const float ls = logf(sd[d]);
      log_norm += ls + kLogSqrt2Pi;
      ent += (0.5f + kLogSqrt2Pi) + ls;
    }
  }
  const float lo = 1.f - p.clip_ratio, hi = 1.f + p.clip_ratio;
  const float inv_B = 1.f / (float)p.B;
  const float gs_scale = -p.w_s * inv_B;                               // d(w_s * L_s)/d(min term)
  const float gv_scale = 2.f * p.w_v / ((float)p.B * (float)p.Dv);     // d(w_v * L_v)/d(value) factor

  float acc_s = 0.f, acc_lr = 0.f, acc_v = 0.f, acc_val = 0.f;
  float acc_dstd[APAD];
#pragma unroll
  for (int d = 0; d < APAD; ++d) acc_dstd[d] = 0.f;

  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.B; i += stride) {
    float mu[APAD], ac[APAD];
    load_row<APAD, VEC4>(p.mean, i, A, mu);
    load_row<APAD, VEC4>(p.action, i, A, ac);
    const float lp_old = ldg_stream(p.logp_old + i);
    const float adv = ldg_stream(p.advantage + i);
    // log_prob (distribution.py:207-209 -> torch Normal.log_prob): -(a-mu)^2/(2 var) - log sd - log sqrt(2 pi)
    float quad = 0.f;
    float diff[APAD];
#pragma unroll
    for (int d = 0; d < APAD; ++d) {
      diff[d] = ac[d] - mu[d];
      if (d < A) quad += (diff[d] * diff[d]) * (0.5f * inv_var[d]);
    }
    const float logp = -quad - log_norm;
    const float lr = logp - lp_old;     // common.py:35
    const float r = expf(lr);           // common.py:41
    // ppo.py:15-18
    const float s1 = adv * r;
    const float rc = fminf(fmaxf(r, lo), hi);
    const float s2 = adv * rc;
    acc_s += fminf(s1, s2);
    acc_lr += fabsf(lr);
    // autograd of -mean(min(s1, s2)): torch.min splits ties 1/2-1/2, clamp passes on the closed interval
    const float w1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
    const float w2 = 1.f - w1;
    const float in_range = (r >= lo && r <= hi) ? 1.f : 0.f;
    const float ds_dr = adv * (w1 + w2 * in_range);
    const float g = gs_scale * ds_dr * r;  // d(w_s L_s)/dlogp_i

    if (p.logp) p.logp[i] = logp;
    if (p.entropy) p.entropy[i] = ent;
    if (p.logp_ratio) p.logp_ratio[i] = lr;
    if (p.prob_ratio) p.prob_ratio[i] = r;

    float dmu[APAD];
#pragma unroll
    for (int d = 0; d < APAD; ++d) {
      const float z = diff[d] * inv_var[d];          // dlogp/dmu_d = (a-mu)/var
      dmu[d] = g * z;
      // dlogp/dsd_d = (a-mu)^2/sd^3 - 1/sd
      acc_dstd[d] += g * (diff[d] * z * inv_sd[d] - inv_sd[d]);
    }
    if (p.d_mean) {
      float* o = p.d_mean + i * A;
      if (VEC4) {
#pragma unroll
        for (int q = 0; q < APAD / 4; ++q)
          reinterpret_cast<float4*>(o)[q] = make_float4(dmu[4 * q], dmu[4 * q + 1], dmu[4 * q + 2], dmu[4 * q + 3]);
      } else {
#pragma unroll
        for (int d = 0; d < APAD; ++d)
          if (d < A) o[d] = dmu[d];
      }
    }

    if (p.has_value) {
      for (int k = 0; k < p.Dv; ++k) {
        const int64_t j = i * p.Dv + k;
        const float v = ldg_stream(p.curr_value + j);
        const float rt = ldg_stream(p.ret + j);
        acc_val += v;
        float dv;
        if (p.value_clip > 0.f) {
          // value.py:85-89
          const float vo = ldg_stream(p.value_old + j);
          const float dlt = v - vo;
          const float cl = vo + fminf(fmaxf(dlt, -p.value_clip), p.value_clip);
          const float e1 = v - rt, e2 = cl - rt;
          const float l1 = e1 * e1, l2 = e2 * e2;
          acc_v += fmaxf(l1, l2);
          const float m1 = l1 > l2 ? 1.f : (l1 == l2 ? 0.5f : 0.f);
          const float pass = (dlt >= -p.value_clip && dlt <= p.value_clip) ? 1.f : 0.f;
          dv = gv_scale * (m1 * e1 + (1.f - m1) * pass * e2);
        } else {
          // value.py:131-133  mse_loss(return_, curr_value)
          const float e = rt - v;
          acc_v += e * e;
          dv = -gv_scale * e;
        }
        if (p.d_value) p.d_value[j] = dv;
      }
    }
  }

  // ---- block reduction -> partials
  double head[kLossHead] = {(double)acc_s, (double)acc_lr, (double)acc_v, (double)acc_val};
  block_sum<kLossHead>(head, smem);
  double* out = p.partials + (int64_t)blockIdx.x * (kLossHead + kMaxA);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kLossHead; ++k) out[k] = head[k];
  }
#pragma unroll
  for (int d0 = 0; d0 < APAD; d0 += 4) {
    double v4[4] = {(double)acc_dstd[d0], (double)acc_dstd[d0 + 1], (double)acc_dstd[d0 + 2], (double)acc_dstd[d0 + 3]};
    block_sum<4>(v4, smem);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out[kLossHead + d0 + k] = v4[k];
    }
  }
}

__global__ void ppo_loss_finalize(const double* __restrict__ partials, int nblocks, const float* __restrict__ std,
                                  int64_t B, int A, int Dv, int has_value, float w_s, float w_e, float w_v,
                                  float* __restrict__ losses, float* __restrict__ metrics,
                                  float* __restrict__ d_std_surr, float* __restrict__ d_std_ent) {
  // one warp per output slot, fixed summation order -> deterministic
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = kLossHead + A;
  __shared__ double res[kLossHead + kMaxA];
  for (int s = slot; s < nslots; s += (blockDim.x >> 5)) {
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) acc += partials[(int64_t)b * (kLossHead + kMaxA) + s];
    acc = warp_sum(acc);
    if (lane == 0) res[s] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = (double)B;
    float ent = 0.f;
    for (int d = 0; d < A; ++d) ent += (0.5f + 0.91893853320467274178f) + logf(std[d]);
    if (losses) {
      losses[0] = has_value ? (float)(res[2] / (n * Dv)) * w_v : 0.f;  // value.py:137
      losses[1] = (float)(-res[0] / n) * w_s;                          // ppo.py:55
      losses[2] = (-ent) * w_e;                                        // ppo.py:83-84
    }
    if (metrics) {
      metrics[0] = (float)(res[1] / n);
      metrics[1] = ent;
      metrics[2] = has_value ? (float)(res[3] / n) : 0.f;
    }
  }
  if (threadIdx.x < A) {
    const int d = threadIdx.x;
    if (d_std_surr) d_std_surr[d] = (float)res[kLossHead + d];
    if (d_std_ent) d_std_ent[d] = -w_e / std[d];  // d(-w_e * mean_i sum_d log sd_d)/dsd_d
  }
}

template <typename F>
static int dispatch_apad(int A, bool vec4, F&& f) {
  const int apad = (A + 3) & ~3;
  switch (apad) {
#define CASE(P)                      \
  case P:                            \
    return vec4 ? f.template run<P, true>() : f.template run<P, false>();
    CASE(4) CASE(8) CASE(12) CASE(16) CASE(20) CASE(24) CASE(28) CASE(32)
#undef CASE
  }
  return CUSRL_B200_EUNSUPPORTED;
}

struct LossLauncher {
  LossParams p;
  unsigned grid;
  cudaStream_t s;
  template <int APAD, bool VEC4>
  int run() {
    ppo_loss_kernel<APAD, VEC4><<<grid, kLossThreads, 0, s>>>(p);
    return check_launch("ppo_loss_kernel");
  }
};

__global__ void scale_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ scale) {
  const float s = *scale;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// dz = dy * act'(z) expressed through the stored post-activation y (ELU: y > 0 ? 1 : y + 1; ReLU: y > 0)
__global__ void act_grad_mul_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                                    int64_t n, int act) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float yi = y[i];
    const float g = act == 1 ? (yi > 0.f ? 1.f : yi + 1.f) : (act == 2 ? (yi > 0.f ? 1.f : 0.f) : 1.f);
    dz[i] = dy[i] * g;
  }
}

// ------------------------------------------------------------------------------------------------
// OnPolicyStatistics (hook/on_policy/stats.py:29-40): KL(old || new) of diagonal normals
// (torch kl_divergence(Normal, Normal): 0.5*(var_ratio + t1 - 1 - log var_ratio), var_ratio=(s_p/s_q)^2,
// t1 = ((mu_p-mu_q)/s_q)^2), importance-weighted advantage and mean std.
// ------------------------------------------------------------------------------------------------
constexpr int kStatSlots = 2;

__global__ void __launch_bounds__(kLossThreads) policy_stats_kernel(
    const float* __restrict__ mean_old, const float* __restrict__ std_old, const float* __restrict__ mean_new,
    const float* __restrict__ std_new, const float* __restrict__ action, const float* __restrict__ logp_old,
    const float* __restrict__ advantage, int64_t E, int A, double* __restrict__ partials) {
  __shared__ double smem[kStatSlots * 32];
  __shared__ float s_sd[kMaxA], s_lognorm;
  if (threadIdx.x < A) s_sd[threadIdx.x] = std_new[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    float ln = 0.f;
    for (int d = 0; d < A; ++d) ln += logf(s_sd[d]) + 0.91893853320467274178f;
    s_lognorm = ln;
  }
  __syncthreads();
  float acc_kl = 0.f, acc_iwa = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += stride) {
    float kl = 0.f, quad = 0.f;
    for (int d = 0; d < A; ++d) {
      const float mp = __ldg(mean_old + i * A + d), sp = __ldg(std_old + i * A + d);
      const float mq = __ldg(mean_new + i * A + d), sq = s_sd[d];
      const float ratio = sp / sq;
      const float vr = ratio * ratio;
      const float t = (mp - mq) / sq;
      kl += 0.5f * (vr + t * t - 1.f - logf(vr));
      const float df = __ldg(action + i * A + d) - mq;
      quad += df * df / (2.f * sq * sq);
    }
    acc_kl += kl;
    const float logp = -quad - s_lognorm;
    acc_iwa += advantage[i] * expf(logp - logp_old[i]);
  }
  double v[kStatSlots] = {(double)acc_kl, (double)acc_iwa};
  block_sum<kStatSlots>(v, smem);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x + 0] = v[0];
    partials[2 * blockIdx.x + 1] = v[1];
  }
}

__global__ void policy_stats_finalize(const double* __restrict__ partials, int nblocks, const float* __restrict__ std_new,
                                      int64_t E, int A, float* __restrict__ out) {
  __shared__ double smem[kStatSlots * 32];
  double v[kStatSlots] = {0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v[0] += partials[2 * b], v[1] += partials[2 * b + 1];
  block_sum<kStatSlots>(v, smem);
  if (threadIdx.x == 0) {
    out[0] = (float)(v[0] / (double)E);
    out[1] = (float)(v[1] / (double)E);
    float s = 0.f;
    for (int d = 0; d < A; ++d) s += std_new[d];
    out[2] = s / (float)A;
  }
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

size_t cusrl_b200_ppo_loss_scratch_bytes(int64_t A) {
  (void)A;
  return (size_t)kLossMaxBlocks * (kLossHead + kMaxA) * sizeof(double);
}

int cusrl_b200_ppo_loss_f32(const float* mean, const float* std, const float* action, const float* logp_old,
                            const float* advantage, const float* ret, const float* value_old,
                            const float* curr_value, int64_t B, int64_t A, int64_t Dv, int has_value,
                            float clip_ratio, float w_surrogate, float w_entropy, float w_value, float value_clip,
                            float* logp, float* entropy, float* logp_ratio, float* prob_ratio, float* losses,
                            float* metrics, float* d_mean, float* d_std_surr, float* d_std_ent, float* d_value,
                            void* scratch, size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(mean && std && action && logp_old && advantage && scratch, CUSRL_B200_EINVAL,
                "ppo_loss: null pointer");
  CUSRL_REQUIRE(B > 0 && A > 0 && Dv > 0, CUSRL_B200_EINVAL, "ppo_loss: B, A, Dv must be positive");
  CUSRL_REQUIRE(A <= kMaxA, CUSRL_B200_EUNSUPPORTED, "ppo_loss: A > %d", kMaxA);
  CUSRL_REQUIRE(!has_value || (curr_value && ret), CUSRL_B200_EINVAL, "ppo_loss: value loss needs curr_value and ret");
  CUSRL_REQUIRE(!(has_value && value_clip > 0.f) || value_old, CUSRL_B200_EINVAL,
                "ppo_loss: clipped value loss needs value_old");
  // same domain checks as PpoSurrogateLoss / EntropyLoss / ValueLoss constructors (ppo.py:38-41,73-74, value.py:108-111)
  CUSRL_REQUIRE(clip_ratio > 0.f, CUSRL_B200_EINVAL, "ppo_loss: 'clip_ratio' must be positive");
  CUSRL_REQUIRE(w_surrogate >= 0.f && w_entropy >= 0.f, CUSRL_B200_EINVAL, "ppo_loss: 'weight' must be non-negative");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_ppo_loss_scratch_bytes(A), CUSRL_B200_ESCRATCH, "ppo_loss: scratch too small");
  CUSRL_REQUIRE(aligned_to(scratch, 8), CUSRL_B200_EALIGN, "ppo_loss: scratch must be 8-byte aligned");

  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (B + kLossThreads - 1) / kLossThreads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kLossMaxBlocks) cap = kLossMaxBlocks;
  if (blocks > cap) blocks = cap;

  LossLauncher L;
  L.p = LossParams{mean, std, action, logp_old, advantage, ret, value_old, curr_value, B, (int)A, (int)Dv, has_value,
                   clip_ratio, w_surrogate, w_entropy, w_value, value_clip, logp, entropy, logp_ratio, prob_ratio,
                   d_mean, d_value, (double*)scratch};
  L.grid = (unsigned)blocks;
  L.s = s;
  const bool vec4 = (A % 4 == 0) && aligned_to(mean, 16) && aligned_to(action, 16) && (!d_mean || aligned_to(d_mean, 16));
  if (int e = dispatch_apad((int)A, vec4, L)) return e;
  ppo_loss_finalize<<<1, 512, 0, s>>>((const double*)scratch, (int)blocks, std, B, (int)A, (int)Dv, has_value,
                                      w_surrogate, w_entropy, w_value, losses, metrics, d_std_surr, d_std_ent);
  return check_launch("ppo_loss_finalize");
}

int cusrl_b200_scale_f32(float* x, int64_t n, const float* scale_dev, void* stream) {
  CUSRL_REQUIRE(x && scale_dev, CUSRL_B200_EINVAL, "scale: null pointer");
  CUSRL_REQUIRE(n >= 0, CUSRL_B200_EINVAL, "scale: negative size");
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, scale_dev);
  return check_launch("scale_kernel");
}

int cusrl_b200_act_grad_mul_f32(const float* dy, const float* y, float* dz, int64_t n, int act, void* stream) {
  CUSRL_REQUIRE(dy && y && dz, CUSRL_B200_EINVAL, "act_grad_mul: null pointer");
  CUSRL_REQUIRE(n >= 0 && act >= 0 && act <= 2, CUSRL_B200_EINVAL, "act_grad_mul: bad arguments");
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  act_grad_mul_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, dz, n, act);
  return check_launch("act_grad_mul_kernel");
}

size_t cusrl_b200_policy_stats_scratch_bytes(void) { return (size_t)kLossMaxBlocks * kStatSlots * sizeof(double); }

int cusrl_b200_policy_stats_f32(const float* mean_old, const float* std_old, const float* mean_new,
                                const float* std_new, const float* action, const float* logp_old,
                                const float* advantage, int64_t E, int64_t A, float* out, void* scratch,
                                size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(mean_old && std_old && mean_new && std_new && action && logp_old && advantage && out && scratch,
                CUSRL_B200_EINVAL, "policy_stats: null pointer");
  CUSRL_REQUIRE(E > 0 && A > 0, CUSRL_B200_EINVAL, "policy_stats: E, A must be positive");
  CUSRL_REQUIRE(A <= kMaxA, CUSRL_B200_EUNSUPPORTED, "policy_stats: A > %d", kMaxA);
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_policy_stats_scratch_bytes(), CUSRL_B200_ESCRATCH,
                "policy_stats: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (E + kLossThreads - 1) / kLossThreads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kLossMaxBlocks) cap = kLossMaxBlocks;
  if (blocks > cap) blocks = cap;
  policy_stats_kernel<<<(unsigned)blocks, kLossThreads, 0, s>>>(mean_old, std_old, mean_new, std_new, action, logp_old,
                                                                advantage, E, (int)A, (double*)scratch);
  if (int e = check_launch("policy_stats_kernel")) return e;
  policy_stats_finalize<<<1, 256, 0, s>>>((const double*)scratch, (int)blocks, std_new, E, (int)A, out);
  return check_launch("policy_stats_finalize");
}

}  // extern "C"

// ================================================================================================
// K5: Random Network Distillation arithmetic (hook/auxiliary/rnd.py:68-81).  The two small MLPs run on the K6 GEMMs;
// these kernels fuse everything after them.
// ================================================================================================
namespace cusrl_b200 {

// rnd_reward[m] = scale * mean_d (target[m,d] - pred[m,d])^2 ;  reward[m, 0..Dr) += rnd_reward[m]   (rnd.py:72-74)
__global__ void __launch_bounds__(256) rnd_reward_kernel(const float* __restrict__ target, const float* __restrict__ pred,
                                                         int64_t M, int D, float scale, float* __restrict__ reward, int Dr,
                                                         float* __restrict__ rnd_reward, double* __restrict__ partials) {
  __shared__ double smem[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += stride) {
    float s = 0.f;
    if ((D & 3) == 0) {
      for (int d = 0; d < D; d += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(target + m * D + d));
        const float4 q = __ldg(reinterpret_cast<const float4*>(pred + m * D + d));
        const float a = t.x - q.x, b = t.y - q.y, c = t.z - q.z, e = t.w - q.w;
        s += (a * a + b * b) + (c * c + e * e);
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const float a = target[m * D + d] - pred[m * D + d];
        s += a * a;
      }
    }
    const float r = scale * (s / (float)D);
    if (rnd_reward) rnd_reward[m] = r;
    for (int k = 0; k < Dr; ++k) reward[m * Dr + k] += r;  // broadcast add over the reward dim (torch add_ broadcasting)
    acc += r;
  }
  double v[1] = {(double)acc};
  block_sum<1>(v, smem);
  if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

// loss = mean over M*D of (pred - target)^2 ;  d_pred = 2 (pred - target) / (M*D)     (nn.MSELoss, rnd.py:80)
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                  float inv_n, float* __restrict__ d_pred, double* __restrict__ partials) {
  __shared__ double smem[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float e = pred[i] - target[i];
    acc += e * e;
    if (d_pred) d_pred[i] = 2.f * e * inv_n;
  }
  double v[1] = {(double)acc};
  block_sum<1>(v, smem);
  if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double scale, float* __restrict__ out) {
  __shared__ double smem[32];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += partials[i];
  block_sum<1>(v, smem);
  if (threadIdx.x == 0) *out = (float)(v[0] * scale);
}

}  // namespace cusrl_b200

extern "C" {

size_t cusrl_b200_rnd_scratch_bytes(void) { return (size_t)kLossMaxBlocks * sizeof(double); }

int cusrl_b200_rnd_reward_f32(const float* target, const float* pred, int64_t M, int64_t D, float reward_scale, float* reward,
                              int64_t Dr, float* rnd_reward, float* mean_out, void* scratch, size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(target && pred && reward && scratch, CUSRL_B200_EINVAL, "rnd_reward: null pointer");
  CUSRL_REQUIRE(M > 0 && D > 0 && Dr > 0, CUSRL_B200_EINVAL, "rnd_reward: M, D, Dr must be positive");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_rnd_scratch_bytes(), CUSRL_B200_ESCRATCH, "rnd_reward: scratch too small");
  CUSRL_REQUIRE((D & 3) || (aligned_to(target, 16) && aligned_to(pred, 16)), CUSRL_B200_EALIGN, "rnd_reward: alignment");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (M + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kLossMaxBlocks) cap = kLossMaxBlocks;
  if (blocks > cap) blocks = cap;
  rnd_reward_kernel<<<(unsigned)blocks, 256, 0, s>>>(target, pred, M, (int)D, reward_scale, reward, (int)Dr, rnd_reward,
                                                     (double*)scratch);
  if (int e = check_launch("rnd_reward_kernel")) return e;
  if (mean_out) {
    sum_partials_kernel<<<1, 256, 0, s>>>((const double*)scratch, (int)blocks, 1.0 / (double)M, mean_out);
    return check_launch("sum_partials_kernel");
  }
  return 0;
}

int cusrl_b200_mse_f32(const float* pred, const float* target, int64_t M, int64_t D, float* loss, float* d_pred, void* scratch,
                       size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(pred && target && loss && scratch, CUSRL_B200_EINVAL, "mse: null pointer");
  CUSRL_REQUIRE(M > 0 && D > 0, CUSRL_B200_EINVAL, "mse: M, D must be positive");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_rnd_scratch_bytes(), CUSRL_B200_ESCRATCH, "mse: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = M * D;
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kLossMaxBlocks) cap = kLossMaxBlocks;
  if (blocks > cap) blocks = cap;
  mse_kernel<<<(unsigned)blocks, 256, 0, s>>>(pred, target, n, 1.f / (float)n, d_pred, (double*)scratch);
  if (int e = check_launch("mse_kernel")) return e;
  sum_partials_kernel<<<1, 256, 0, s>>>((const double*)scratch, (int)blocks, 1.0 / (double)n, loss);
  return check_launch("sum_partials_kernel");
}

}  // extern "C"
