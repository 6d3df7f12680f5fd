// K8 multi-field row gather (minibatch sampling) and K9 grad-norm clip + Adam on a flat arena.
#include <math.h>

#include "common.cuh"
#include "f16x3_common.cuh"

namespace cusrl_b200 {

// ------------------------------------------------------------------------------------------------
// K8: sampler/mini_batch_sampler.py:76,89 -- `leaf.flatten(0,1)[indices]` for every consumed leaf,
// one launch: blockIdx.y selects the field, consecutive threads move consecutive vectors of a row.
// ------------------------------------------------------------------------------------------------
struct GatherField {
  const uint8_t* src;
  uint8_t* dst;
  int64_t src_stride, dst_stride;
  int32_t payload_vecs;  // vectors carrying source data
  int32_t row_vecs;      // vectors per destination row (payload + zero padding)
  int32_t vec_bytes;     // 16, 8, 4, 2 or 1
};
struct GatherArgs {
  GatherField f[CUSRL_B200_MAX_GATHER_FIELDS];
  const int64_t* index;
  int64_t n_index, n_src_rows;
};

template <typename V>
__device__ __forceinline__ void gather_field(const GatherField& f, const int64_t* __restrict__ index, int64_t n_index,
                                             int64_t n_src_rows) {
  const int64_t total = n_index * f.row_vecs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / f.row_vecs;
    const int32_t v = (int32_t)(i - row * f.row_vecs);
    V val;
    memset(&val, 0, sizeof(V));
    if (v < f.payload_vecs) {
      int64_t src_row = __ldg(index + row);
      // out-of-range indices are a caller bug; clamp instead of faulting (torch would raise)
      src_row = src_row < 0 ? 0 : (src_row >= n_src_rows ? n_src_rows - 1 : src_row);
      val = *reinterpret_cast<const V*>(f.src + src_row * f.src_stride + (int64_t)v * sizeof(V));
    }
    *reinterpret_cast<V*>(f.dst + row * f.dst_stride + (int64_t)v * sizeof(V)) = val;
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherArgs a) {
  const GatherField& f = a.f[blockIdx.y];
  switch (f.vec_bytes) {
    case 16: gather_field<uint4>(f, a.index, a.n_index, a.n_src_rows); break;
    case 8: gather_field<uint2>(f, a.index, a.n_index, a.n_src_rows); break;
    case 4: gather_field<uint32_t>(f, a.index, a.n_index, a.n_src_rows); break;
    case 2: gather_field<uint16_t>(f, a.index, a.n_index, a.n_src_rows); break;
    default: gather_field<uint8_t>(f, a.index, a.n_index, a.n_src_rows); break;
  }
}

// K8 for the f16x3 dense layers: dst pair[i] = split(src[index[i]]) -- the gathered minibatch rows of a wide fp32 leaf
// (observation / state) emitted directly as the fp16 hi / lo pair the first trunk layer consumes, so the minibatch is not
// re-read by a separate split pass.  One thread per 8 consecutive columns (two 16-byte loads, one 16-byte store per half).
__global__ void __launch_bounds__(256) gather_split_f16_kernel(const float* __restrict__ src, int64_t lds, const int64_t* __restrict__ index,
                                                               int64_t n, int64_t n_src_rows, int width, const float* __restrict__ bound,
                                                               __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldh) {
  const float s = f16x3_scale(__ldg(bound));
  const int units = (int)(ldh >> 3);
  const int64_t total = n * units;
  const bool vec = (lds & 3) == 0 && aligned_to(src, 16);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / units;
    const int c = (int)(i - r * units) * 8;
    int64_t row = __ldg(index + r);
    row = row < 0 ? 0 : (row >= n_src_rows ? n_src_rows - 1 : row);
    const float* p = src + row * lds + c;
    float v[8];
    if (vec && c + 8 <= width) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c + k < width ? __ldg(p + k) : 0.f;
    }
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
    *reinterpret_cast<uint4*>(hi + r * ldh + c) = h;
    *reinterpret_cast<uint4*>(lo + r * ldh + c) = l;
  }
}

// ------------------------------------------------------------------------------------------------
// K9: hook/on_policy/gradient_clipping.py:56-75 (torch clip_grad_norm_) + torch.optim.Adam.step
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  __shared__ double smem[32];
  double acc[1] = {0.0};
  float s = 0.f;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = aligned_to(g, 16) ? (n >> 2) : 0;
  int cnt = 0;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (++cnt == 8) acc[0] += s, s = 0.f, cnt = 0;
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += stride) s += g[i] * g[i];
  acc[0] += s;
  block_sum<1>(acc, smem);
  if (threadIdx.x == 0) atomicAdd(out, acc[0]);
}

__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float max_norm, float* __restrict__ norm_out,
                                 float* __restrict__ coef_out) {
  const float norm = (float)sqrt(*sumsq);
  if (norm_out) *norm_out = norm;
  // torch clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1.0
  if (coef_out) *coef_out = fminf(max_norm / (norm + 1e-6f), 1.0f);
}

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ coef_dev, float lr, float beta1,
                                                        float beta2, float eps, float weight_decay, float bc1,
                                                        float bc2_sqrt) {
  const float coef = coef_dev ? *coef_dev : 1.f;
  const float step_size = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i] * coef;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = gi + weight_decay * pi;
    // torch/optim/adam.py _single_tensor_adam: exp_avg.lerp_(grad, 1-beta1);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2);
    // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps); param.addcdiv_(exp_avg, denom, value=-step_size)
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / denom);
  }
}

// Same update with the learning rate and the step count read from DEVICE memory, so that a captured CUDA graph of the
// optimizer step stays valid while both change between replays.  The bias corrections 1 - beta^step are evaluated in
// double like torch does on the host (adam.py), then rounded to float exactly as the host path rounds them.
__global__ void __launch_bounds__(256) adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                            const float* __restrict__ coef_dev, const float* __restrict__ lr_dev,
                                                            const int64_t* __restrict__ step_dev, float beta1, float beta2,
                                                            float eps, float weight_decay) {
  const float coef = coef_dev ? *coef_dev : 1.f;
  const double step = (double)*step_dev;
  const float bc1 = (float)(1.0 - pow((double)beta1, step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  const float step_size = *lr_dev / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i] * coef;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = gi + weight_decay * pi;
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_gather_rows(const cusrl_b200_gather_field* fields_host, int n_fields, const int64_t* index,
                           int64_t n_index, int64_t n_src_rows, void* stream) {
  CUSRL_REQUIRE(fields_host && index, CUSRL_B200_EINVAL, "gather_rows: null pointer");
  CUSRL_REQUIRE(n_fields > 0 && n_fields <= CUSRL_B200_MAX_GATHER_FIELDS, CUSRL_B200_EINVAL,
                "gather_rows: n_fields must be in [1, %d]", CUSRL_B200_MAX_GATHER_FIELDS);
  CUSRL_REQUIRE(n_index >= 0 && n_src_rows > 0, CUSRL_B200_EINVAL, "gather_rows: bad sizes");
  if (n_index == 0) return 0;
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.index = index, a.n_index = n_index, a.n_src_rows = n_src_rows;
  int64_t max_vecs = 0;
  for (int k = 0; k < n_fields; ++k) {
    const cusrl_b200_gather_field& in = fields_host[k];
    CUSRL_REQUIRE(in.src && in.dst, CUSRL_B200_EINVAL, "gather_rows: field %d has a null pointer", k);
    CUSRL_REQUIRE(in.row_bytes > 0 && in.src_stride >= in.row_bytes && in.dst_stride >= in.row_bytes,
                  CUSRL_B200_EINVAL, "gather_rows: field %d has inconsistent strides", k);
    // widest vector that divides the payload, both strides and both base addresses; bytes
    // row_bytes..dst_stride of every destination row are written as zeros
    int vb = 16;
    while (vb > 1 && ((in.row_bytes % vb) || (in.src_stride % vb) || (in.dst_stride % vb) || !aligned_to(in.src, vb) ||
                      !aligned_to(in.dst, vb)))
      vb >>= 1;
    GatherField& f = a.f[k];
    f.src = (const uint8_t*)in.src, f.dst = (uint8_t*)in.dst;
    f.src_stride = in.src_stride, f.dst_stride = in.dst_stride;
    f.vec_bytes = vb;
    f.payload_vecs = (int32_t)(in.row_bytes / vb);
    f.row_vecs = (int32_t)(in.dst_stride / vb);
    const int64_t vecs = n_index * f.row_vecs;
    if (vecs > max_vecs) max_vecs = vecs;
  }
  int64_t blocks = (max_vecs + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)n_fields);
  gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("gather_rows_kernel");
}

int cusrl_b200_gather_split_f16(const float* src, int64_t lds, const int64_t* index, int64_t n, int64_t n_src_rows, int64_t width,
                                const float* bound, uint16_t* hi, uint16_t* lo, int64_t ldh, void* stream) {
  CUSRL_REQUIRE(src && index && bound && hi && lo, CUSRL_B200_EINVAL, "gather_split_f16: null pointer");
  CUSRL_REQUIRE(n >= 0 && n_src_rows > 0 && width > 0 && lds >= width && ldh >= width && (ldh % 8) == 0, CUSRL_B200_EINVAL,
                "gather_split_f16: bad sizes (ldh must be a multiple of 8 halves covering the row)");
  CUSRL_REQUIRE(aligned_to(hi, 16) && aligned_to(lo, 16), CUSRL_B200_EALIGN, "gather_split_f16: outputs must be 16-byte aligned");
  if (n == 0) return 0;
  int64_t blocks = (n * (ldh / 8) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gather_split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, index, n, n_src_rows, (int)width, bound,
                                                                            (__half*)hi, (__half*)lo, ldh);
  return check_launch("gather_split_f16_kernel");
}

int cusrl_b200_grad_sumsq_f32(const float* grad, int64_t n, double* sumsq_dev, void* stream) {
  CUSRL_REQUIRE(grad && sumsq_dev, CUSRL_B200_EINVAL, "grad_sumsq: null pointer");
  CUSRL_REQUIRE(n >= 0, CUSRL_B200_EINVAL, "grad_sumsq: negative size");
  if (n == 0) return 0;
  int64_t blocks = ((n + 3) / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  grad_sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad, n, sumsq_dev);
  return check_launch("grad_sumsq_kernel");
}

int cusrl_b200_clip_coef_f32(const double* sumsq_dev, float max_norm, float* norm_dev, float* coef_dev, void* stream) {
  CUSRL_REQUIRE(sumsq_dev, CUSRL_B200_EINVAL, "clip_coef: null pointer");
  CUSRL_REQUIRE(max_norm >= 0.f, CUSRL_B200_EINVAL, "clip_coef: 'max_grad_norm' must be non-negative");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq_dev, max_norm, norm_dev, coef_dev);
  return check_launch("clip_coef_kernel");
}

int cusrl_b200_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* coef_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int64_t step, void* stream) {
  CUSRL_REQUIRE(param && grad && exp_avg && exp_avg_sq, CUSRL_B200_EINVAL, "adam_step: null pointer");
  CUSRL_REQUIRE(n >= 0 && step >= 1, CUSRL_B200_EINVAL, "adam_step: n >= 0 and step >= 1 required");
  CUSRL_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, CUSRL_B200_EINVAL,
                "adam_step: invalid betas / eps");
  if (n == 0) return 0;
  // bias corrections as torch does on the host in double (adam.py: 1 - beta ** step)
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, coef_dev, lr,
                                                                       beta1, beta2, eps, weight_decay, (float)bc1,
                                                                       (float)sqrt(bc2));
  return check_launch("adam_step_kernel");
}

int cusrl_b200_adam_step_dev_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 const float* coef_dev, const float* lr_dev, const int64_t* step_dev, float beta1, float beta2,
                                 float eps, float weight_decay, void* stream) {
  CUSRL_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev, CUSRL_B200_EINVAL, "adam_step_dev: null pointer");
  CUSRL_REQUIRE(n >= 0, CUSRL_B200_EINVAL, "adam_step_dev: negative size");
  CUSRL_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, CUSRL_B200_EINVAL,
                "adam_step_dev: invalid betas / eps");
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_step_dev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, coef_dev, lr_dev,
                                                                           step_dev, beta1, beta2, eps, weight_decay);
  return check_launch("adam_step_dev_kernel");
}

}  // extern "C"
