// Small-N output heads of the actor / critic (mean_head 128 -> 12, value_head 128 -> 1; reference LinearFp32 /
// nn.Linear in nn/module/distribution.py:56, nn/module/critic.py:87-88).  N <= 16 is far below a tensor-core tile,
// and the layers are HBM-bound on the [M, K] latent read, so they run as warp-per-row SIMT kernels in exact fp32:
//   forward : Y[M,No] = H[M,K] @ W[No,K]^T + b
//   backward: dH[M,K] = (dY[M,No] @ W) * act'(H)      (act' of the trunk's last layer, from its output H)
//             dW[No,K] (+)= dY^T @ H ;  db[No] (+)= column sums of dY
//             dbH[K]   (+)= column sums of dH        (optional: the bias gradient of the trunk's last layer)
#include "common.cuh"

namespace cusrl_b200 {

constexpr int kHeadMaxNo = 16;
constexpr int kHeadThreads = 128;  // 4 warps: the backward's shared-memory reduction buffer stays <= 32 KB
constexpr int kHeadMaxBlocks = 592;

__device__ __forceinline__ float head_act_grad(float y, int act) {
  if (act == 1) return y > 0.f ? 1.f : y + 1.f;
  if (act == 2) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}

// K = 128 * KV: lane l owns columns {128*v + 4*l .. +3}
template <int NO, int KV>
__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(const float* __restrict__ H, int64_t ldh,
                                                                const float* __restrict__ W, const float* __restrict__ b,
                                                                float* __restrict__ Y, int M) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 w[NO][KV];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int v = 0; v < KV; ++v) w[o][v] = __ldg(reinterpret_cast<const float4*>(W + o * (128 * KV) + 128 * v + 4 * lane));
  const float bias = (b && lane < NO) ? __ldg(b + lane) : 0.f;
  for (int row = warp; row < M; row += nwarps) {
    float4 h[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) h[v] = ldg_stream4(H + (int64_t)row * ldh + 128 * v + 4 * lane);
    float acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < KV; ++v) a += h[v].x * w[o][v].x + h[v].y * w[o][v].y + h[v].z * w[o][v].z + h[v].w * w[o][v].w;
      acc[o] = a;
    }
    // butterfly: afterwards every lane holds every sum; lane o writes output o
    float mine = 0.f;
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      const float s = warp_sum(acc[o]);
      if (lane == o) mine = s;
    }
    if (lane < NO) Y[(int64_t)row * NO + lane] = mine + bias;
  }
}

template <int NO, int KV>
__global__ void __launch_bounds__(kHeadThreads) head_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ H,
                                                                int64_t ldh, const float* __restrict__ W, int act,
                                                                float* __restrict__ dH, int64_t lddh, int M,
                                                                float* __restrict__ partial /*[blocks][NO*K + NO + K]*/) {
  constexpr int K = 128 * KV;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 w[NO][KV], gw[NO][KV], gd[KV];
  float gb[NO];
#pragma unroll
  for (int v = 0; v < KV; ++v) gd[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    gb[o] = 0.f;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      w[o][v] = __ldg(reinterpret_cast<const float4*>(W + o * K + 128 * v + 4 * lane));
      gw[o][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int row = warp; row < M; row += nwarps) {
    float g[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) g[o] = __ldg(dY + (int64_t)row * NO + o);  // same address in every lane: broadcast
    float4 h[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) h[v] = ldg_stream4(H + (int64_t)row * ldh + 128 * v + 4 * lane);
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        d.x += g[o] * w[o][v].x, d.y += g[o] * w[o][v].y, d.z += g[o] * w[o][v].z, d.w += g[o] * w[o][v].w;
        gw[o][v].x += g[o] * h[v].x, gw[o][v].y += g[o] * h[v].y, gw[o][v].z += g[o] * h[v].z, gw[o][v].w += g[o] * h[v].w;
      }
      if (dH) {
        d.x *= head_act_grad(h[v].x, act), d.y *= head_act_grad(h[v].y, act);
        d.z *= head_act_grad(h[v].z, act), d.w *= head_act_grad(h[v].w, act);
        *reinterpret_cast<float4*>(dH + (int64_t)row * lddh + 128 * v + 4 * lane) = d;
        gd[v].x += d.x, gd[v].y += d.y, gd[v].z += d.z, gd[v].w += d.w;
      }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) gb[o] += g[o];
  }
  // block reduction of the weight-gradient partials through shared memory (fixed order over warps)
  __shared__ float red[kHeadThreads / 32][NO * K > 2048 ? 1 : NO * K];  // only instantiated with NO*K <= 2048
  float* out = partial + (int64_t)blockIdx.x * (NO * K + NO + K);
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int v = 0; v < KV; ++v) *reinterpret_cast<float4*>(&red[wib][o * K + 128 * v + 4 * lane]) = gw[o][v];
  __syncthreads();
  for (int i = threadIdx.x; i < NO * K; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kHeadThreads / 32; ++q) s += red[q][i];
    out[i] = s;
  }
  // bias gradient: every lane of a warp accumulated the same rows -> take lane 0 of each warp
  __shared__ float redb[kHeadThreads / 32][NO];
  if (lane == 0)
#pragma unroll
    for (int o = 0; o < NO; ++o) redb[wib][o] = gb[o];
  __syncthreads();
  if (threadIdx.x < NO) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kHeadThreads / 32; ++q) s += redb[q][threadIdx.x];
    out[NO * K + threadIdx.x] = s;
  }
  // column sums of dH (re-using the reduction buffer)
  __syncthreads();
#pragma unroll
  for (int v = 0; v < KV; ++v) *reinterpret_cast<float4*>(&red[wib][128 * v + 4 * lane]) = gd[v];
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kHeadThreads / 32; ++q) s += red[q][i];
    out[NO * K + NO + i] = s;
  }
}

__global__ void head_bwd_final_kernel(const float* __restrict__ partial, int nblocks, int NO, int K, float* __restrict__ dW,
                                      float* __restrict__ db, int accumulate, float* __restrict__ dbH, int accumulate_dbh) {
  const int total = NO * K + NO + K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const bool trunk = i >= NO * K + NO;
  float* dst = i < NO * K ? dW + i : (trunk ? (dbH ? dbH + (i - NO * K - NO) : nullptr) : (db ? db + (i - NO * K) : nullptr));
  if (!dst) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += partial[(int64_t)b * total + i];
  const int add = trunk ? accumulate_dbh : accumulate;
  *dst = add ? *dst + (float)acc : (float)acc;
}

template <typename F>
static int head_dispatch(int No, int K, F&& f) {
  const int kv = K / 128;
#define CUSRL_HEAD_CASE(NO_, KV_) \
  if (No == NO_ && kv == KV_) return f.template run<NO_, KV_>();
  CUSRL_HEAD_CASE(1, 1) CUSRL_HEAD_CASE(2, 1) CUSRL_HEAD_CASE(3, 1) CUSRL_HEAD_CASE(4, 1) CUSRL_HEAD_CASE(5, 1)
  CUSRL_HEAD_CASE(6, 1) CUSRL_HEAD_CASE(7, 1) CUSRL_HEAD_CASE(8, 1) CUSRL_HEAD_CASE(9, 1) CUSRL_HEAD_CASE(10, 1)
  CUSRL_HEAD_CASE(11, 1) CUSRL_HEAD_CASE(12, 1) CUSRL_HEAD_CASE(13, 1) CUSRL_HEAD_CASE(14, 1) CUSRL_HEAD_CASE(15, 1)
  CUSRL_HEAD_CASE(16, 1) CUSRL_HEAD_CASE(1, 2) CUSRL_HEAD_CASE(2, 2) CUSRL_HEAD_CASE(4, 2) CUSRL_HEAD_CASE(6, 2)
  CUSRL_HEAD_CASE(8, 2)
#undef CUSRL_HEAD_CASE
  set_last_error("head: unsupported (outputs=%d, latent=%d); supported: latent 128 with 1..16 outputs, latent 256 with 1/2/4/6/8", No, K);
  return CUSRL_B200_EUNSUPPORTED;
}

struct HeadFwd {
  const float *H, *W, *b;
  int64_t ldh;
  float* Y;
  int M;
  unsigned grid;
  cudaStream_t s;
  template <int NO, int KV>
  int run() {
    head_fwd_kernel<NO, KV><<<grid, kHeadThreads, 0, s>>>(H, ldh, W, b, Y, M);
    return check_launch("head_fwd_kernel");
  }
};
struct HeadBwd {
  const float *dY, *H, *W;
  int64_t ldh, lddh;
  int act, M;
  float *dH, *partial;
  unsigned grid;
  cudaStream_t s;
  template <int NO, int KV>
  int run() {
    head_bwd_kernel<NO, KV><<<grid, kHeadThreads, 0, s>>>(dY, H, ldh, W, act, dH, lddh, M, partial);
    return check_launch("head_bwd_kernel");
  }
};

static unsigned head_grid(int64_t M) {
  int64_t blocks = (M + (kHeadThreads / 32) - 1) / (kHeadThreads / 32);
  int64_t cap = (int64_t)sm_count() * 4;
  if (cap > kHeadMaxBlocks) cap = kHeadMaxBlocks;
  return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_head_fwd_f32(const float* H, int64_t ldh, const float* W, const float* bias, float* Y, int64_t M,
                            int64_t K, int64_t No, void* stream) {
  CUSRL_REQUIRE(H && W && Y, CUSRL_B200_EINVAL, "head_fwd: null pointer");
  CUSRL_REQUIRE(M > 0 && K > 0 && No > 0 && M < (1ll << 31), CUSRL_B200_EINVAL, "head_fwd: bad sizes");
  CUSRL_REQUIRE((K % 128) == 0 && No <= kHeadMaxNo, CUSRL_B200_EUNSUPPORTED,
                "head_fwd: latent dim must be a multiple of 128 and outputs <= %d", kHeadMaxNo);
  CUSRL_REQUIRE((ldh % 4) == 0 && ldh >= K && aligned_to(H, 16) && aligned_to(W, 16), CUSRL_B200_EALIGN,
                "head_fwd: H and W must be 16-byte aligned with ldh a multiple of 4");
  HeadFwd f{H, W, bias, ldh, Y, (int)M, head_grid(M), (cudaStream_t)stream};
  return head_dispatch((int)No, (int)K, f);
}

size_t cusrl_b200_head_bwd_scratch_bytes(int64_t K, int64_t No) {
  return (size_t)kHeadMaxBlocks * (size_t)(No * K + No + K) * sizeof(float);
}

int cusrl_b200_head_bwd_f32(const float* dY, const float* H, int64_t ldh, const float* W, int act, float* dH,
                            int64_t lddh, float* dW, float* db, int64_t M, int64_t K, int64_t No, int accumulate,
                            float* dbH, int accumulate_dbh, void* scratch, size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(dY && H && W && dW && scratch, CUSRL_B200_EINVAL, "head_bwd: null pointer");
  CUSRL_REQUIRE(M > 0 && K > 0 && No > 0 && M < (1ll << 31), CUSRL_B200_EINVAL, "head_bwd: bad sizes");
  CUSRL_REQUIRE((K % 128) == 0 && No <= kHeadMaxNo && No * K <= 2048, CUSRL_B200_EUNSUPPORTED,
                "head_bwd: unsupported head shape");
  CUSRL_REQUIRE((ldh % 4) == 0 && ldh >= K && (!dH || ((lddh % 4) == 0 && lddh >= K)) && aligned_to(H, 16) &&
                    aligned_to(W, 16) && (!dH || aligned_to(dH, 16)),
                CUSRL_B200_EALIGN, "head_bwd: alignment");
  CUSRL_REQUIRE(act >= 0 && act <= 2, CUSRL_B200_EINVAL, "head_bwd: unknown activation code");
  CUSRL_REQUIRE(!dbH || dH, CUSRL_B200_EINVAL, "head_bwd: dbH (column sums of dH) needs dH");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_head_bwd_scratch_bytes(K, No), CUSRL_B200_ESCRATCH, "head_bwd: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  HeadBwd f{dY, H, W, ldh, lddh, act, (int)M, dH, (float*)scratch, head_grid(M), s};
  if (int e = head_dispatch((int)No, (int)K, f)) return e;
  const int total = (int)(No * K + No + K);
  head_bwd_final_kernel<<<(total + 255) / 256, 256, 0, s>>>((const float*)scratch, (int)f.grid, (int)No, (int)K, dW, db,
                                                            accumulate, dbH, accumulate_dbh);
  return check_launch("head_bwd_final_kernel");
}

}  // extern "C"
