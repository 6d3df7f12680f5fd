// Running mean / std observation normalisation: the per-step arithmetic of
//   cusrl/nn/utils/normalization.py:15-49 (mean_var_count: torch.var_mean(input, dim=0, correction=0)),
//   cusrl/nn/utils/normalization.py:78-93 (merge_mean_var_: Chan et al. parallel update) + cusrl/nn/layer/rms.py:157-167,
//   cusrl/nn/layer/rms.py:202-214 (normalize / normalize_: (x - mean) / std, clamp)
// as three launches per environment step instead of ~15 ATen kernels: a streaming column-statistics pass over the [N, C]
// observation (HBM-bound, one read), a C-element merge into the running statistics, and the normalisation (one read, one
// write, rows may be pitched: it writes straight into padded buffer rows).
#include "common.cuh"

namespace cusrl_b200 {

constexpr int kRmsThreads = 256;
constexpr int kRmsMaxBlocks = 592;

// partials: [blocks][2][C] doubles (sum | sum of squares); block b handles rows b, b + grid, ...; thread t columns t, t + 256, ...
// (lanes run along the contiguous feature axis: coalesced).
__global__ void __launch_bounds__(kRmsThreads) column_stats_partial_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int C,
                                                                          double* __restrict__ partials) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    float fs = 0.f, fq = 0.f;
    int cnt = 0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
      const float v = ldg_stream(x + r * ld + c);
      fs += v, fq += v * v;
      if (++cnt == 32) s += fs, q += fq, fs = 0.f, fq = 0.f, cnt = 0;  // spill the fp32 running sums into double
    }
    s += fs, q += fq;
    partials[((int64_t)blockIdx.x * 2 + 0) * C + c] = s;
    partials[((int64_t)blockIdx.x * 2 + 1) * C + c] = q;
  }
}

// mean_var[0:C] = mean, mean_var[C:2C] = population variance (correction = 0), fixed-order finalisation in double
__global__ void column_stats_final_kernel(const double* __restrict__ partials, int nblocks, int C, int64_t rows,
                                          float* __restrict__ mean_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partials[((int64_t)b * 2 + 0) * C + c], q += partials[((int64_t)b * 2 + 1) * C + c];
  const double n = (double)rows, mean = s / n;
  double var = q / n - mean * mean;
  mean_var[c] = (float)mean;
  mean_var[C + c] = (float)(var < 0.0 ? 0.0 : var);
}

// merge_mean_var_(mean, var, w_old, batch_mean, batch_var, w_new) followed by std = sqrt(var + eps)   (rms.py:157-167)
__global__ void rms_merge_kernel(float* __restrict__ mean, float* __restrict__ var, float* __restrict__ std_,
                                 const float* __restrict__ batch_mean, const float* __restrict__ batch_var, int C, double w_old,
                                 double w_new, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // python-float (double) weights rounded when they meet the fp32 tensors, as in the reference
  const double w_sum = w_old + w_new;
  const float wo = (float)(w_old / w_sum), wn = (float)(w_new / w_sum), wp = (float)((w_old / w_sum) * (w_new / w_sum));
  const float m = mean[c], v = var[c];
  const float delta = batch_mean[c] - m;
  const float m2 = m + delta * wn;
  const float v2 = v + ((batch_var[c] - v) * wn + (delta * delta) * wp);
  mean[c] = m2;
  var[c] = v2;
  std_[c] = sqrtf(v2 + eps);
  (void)wo;
}

// out[r, c] = clamp((x[r, c] - mean[c]) / std[c], -clamp, clamp); padding columns C..ldo-1 of `out` are zeroed (pad != 0)
__global__ void __launch_bounds__(kRmsThreads) rms_normalize_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                                   int64_t ldo, int64_t rows, int C, const float* __restrict__ mean,
                                                                   const float* __restrict__ std_, float clamp, int pad) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int width = pad ? (int)ldo : C;
  for (int64_t r = warp; r < rows; r += nwarps) {
    for (int c = lane; c < width; c += 32) {
      float v = 0.f;
      if (c < C) {
        v = __fdiv_rn(__fsub_rn(x[r * ldx + c], __ldg(mean + c)), __ldg(std_ + c));
        if (clamp > 0.f) v = fminf(fmaxf(v, -clamp), clamp);
      }
      out[r * ldo + c] = v;
    }
  }
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

size_t cusrl_b200_column_stats_scratch_bytes(int64_t C) { return C > 0 ? (size_t)kRmsMaxBlocks * 2 * (size_t)C * sizeof(double) : 0; }

int cusrl_b200_column_stats_f32(const float* x, int64_t ld, int64_t rows, int64_t C, float* mean_var, void* scratch,
                                size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(x && mean_var && scratch, CUSRL_B200_EINVAL, "column_stats: null pointer");
  CUSRL_REQUIRE(rows > 0 && C > 0 && ld >= C && C < (1 << 24), CUSRL_B200_EINVAL, "column_stats: bad sizes");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_column_stats_scratch_bytes(C) && aligned_to(scratch, 8), CUSRL_B200_ESCRATCH,
                "column_stats: scratch too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (int64_t)sm_count() * 4;
  if (blocks > kRmsMaxBlocks) blocks = kRmsMaxBlocks;
  if (blocks > rows) blocks = rows;
  column_stats_partial_kernel<<<(unsigned)blocks, kRmsThreads, 0, s>>>(x, ld, rows, (int)C, (double*)scratch);
  if (int e = check_launch("column_stats_partial_kernel")) return e;
  column_stats_final_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s>>>((const double*)scratch, (int)blocks, (int)C, rows, mean_var);
  return check_launch("column_stats_final_kernel");
}

int cusrl_b200_rms_merge_f32(float* mean, float* var, float* std_, const float* batch_mean, const float* batch_var, int64_t C,
                             double w_old, double w_new, float eps, void* stream) {
  CUSRL_REQUIRE(mean && var && std_ && batch_mean && batch_var, CUSRL_B200_EINVAL, "rms_merge: null pointer");
  CUSRL_REQUIRE(C > 0, CUSRL_B200_EINVAL, "rms_merge: C must be positive");
  CUSRL_REQUIRE(w_old + w_new > 0, CUSRL_B200_EINVAL, "rms_merge: Weight sum must be positive; got %g", w_old + w_new);
  rms_merge_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(mean, var, std_, batch_mean, batch_var, (int)C,
                                                                                  w_old, w_new, eps);
  return check_launch("rms_merge_kernel");
}

int cusrl_b200_rms_normalize_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int64_t C, const float* mean,
                                 const float* std_, float clamp, int zero_padding, void* stream) {
  CUSRL_REQUIRE(x && out && mean && std_, CUSRL_B200_EINVAL, "rms_normalize: null pointer");
  CUSRL_REQUIRE(rows > 0 && C > 0 && ldx >= C && ldo >= C, CUSRL_B200_EINVAL, "rms_normalize: bad sizes");
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rms_normalize_kernel<<<(unsigned)blocks, kRmsThreads, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, rows, (int)C, mean, std_, clamp,
                                                                                  zero_padding);
  return check_launch("rms_normalize_kernel");
}

}  // extern "C"
