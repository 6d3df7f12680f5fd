// Small-N output heads of the actor / critic (mean_head 128 -> 12, value_head 128 -> 1; reference LinearFp32 /
// nn.Linear in nn/module/distribution.py:56, nn/module/critic.py:87-88).  N <= 16 is far below a tensor-core tile,
// and the layers are HBM-bound on the [M, K] latent read, so they run as warp-per-row SIMT kernels in exact fp32:
//   forward : Y[M,No] = H[M,K] @ W[No,K]^T + b
//   backward: dH[M,K] = (dY[M,No] @ W) * act'(H)      (act' of the trunk's last layer, from its output H)
//             dW[No,K] (+)= dY^T @ H ;  db[No] (+)= column sums of dY
//             dbH[K]   (+)= column sums of dH        (optional: the bias gradient of the trunk's last layer)
#include "common.cuh"
#include "f16x3_common.cuh"

namespace cusrl_b200 {

constexpr int kHeadMaxNo = 16;
constexpr int kHeadThreads = 128;  // 4 warps: the backward's shared-memory reduction buffer stays <= 32 KB
constexpr int kHeadMaxBlocks = 592;

__device__ __forceinline__ float head_act_grad(float y, int act) {
  if (act == 1) return y > 0.f ? 1.f : y + 1.f;
  if (act == 2) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}

// Both kernels stream [M, K] rows once and are HBM-bound only if enough loads are in flight: a warp works on 8 rows per
// iteration and issues all their loads before the arithmetic (one row per iteration left the SM with ~8 KB in flight
// and the kernels at ~1.5 TB/s).
constexpr int kHeadRows = 8;  // rows per warp iteration

// forward: 8 lanes per row (lane `sub` owns the float4 columns {q*8 + sub}), 4 rows side by side, 2 row batches per
// iteration; W lives in shared memory (the 8 sub-lanes read one 128-byte line, broadcast over the 4 row groups), so
// the cross-lane reduction is 3 shuffle steps per output instead of 5.
template <int NO, int KV>
__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(const float* __restrict__ H, int64_t ldh,
                                                                const float* __restrict__ W, const float* __restrict__ b,
                                                                float* __restrict__ Y, int M) {
  constexpr int K = 128 * KV, Q = 4 * KV;
  __shared__ float4 sW[NO][K / 4];
  for (int i = threadIdx.x; i < NO * (K / 4); i += blockDim.x)
    sW[i / (K / 4)][i % (K / 4)] = __ldg(reinterpret_cast<const float4*>(W) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 7, rg = lane >> 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float b_lo = (b && sub < NO) ? __ldg(b + sub) : 0.f;
  const float b_hi = (b && sub + 8 < NO) ? __ldg(b + sub + 8) : 0.f;
  for (int64_t base = (int64_t)warp * kHeadRows; base < M; base += (int64_t)nwarps * kHeadRows) {
    float4 h[2][Q];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const int64_t row = base + rb * 4 + rg;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        h[rb][q] = row < M ? ldg_stream4(H + row * ldh + (q * 8 + sub) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc[2][NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float4 w = sW[o][q * 8 + sub];
        a0 += h[0][q].x * w.x + h[0][q].y * w.y + h[0][q].z * w.z + h[0][q].w * w.w;
        a1 += h[1][q].x * w.x + h[1][q].y * w.y + h[1][q].z * w.z + h[1][q].w * w.w;
      }
      acc[0][o] = a0, acc[1][o] = a1;
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        acc[0][o] += __shfl_xor_sync(0xffffffffu, acc[0][o], off);
        acc[1][o] += __shfl_xor_sync(0xffffffffu, acc[1][o], off);
      }
    // every sub-lane now holds every sum of its row; sub-lane s writes outputs s and s + 8
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const int64_t row = base + rb * 4 + rg;
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int o = 0; o < NO; ++o)
        if ((o & 7) == sub) (o < 8 ? lo : hi) = acc[rb][o];
      if (row < M) {
        if (sub < NO) Y[row * NO + sub] = lo + b_lo;
        if (sub + 8 < NO) Y[row * NO + sub + 8] = hi + b_hi;
      }
    }
  }
}

// backward: lane l owns columns {128*v + 4*l .. +3} of every row its warp handles (W and the dW accumulators stay in
// registers); the NO upstream gradients of a row are loaded by lanes 0..NO-1 and broadcast by shuffle.
// PAIR: dH is written as the fp16 hi / lo pair the f16x3 dense-layer kernels consume (f16x3_common.cuh) instead of fp32.  Its
// scale comes from the ANALYTIC bound |dH[m,k]| <= max|dY| * max_k sum_o |W[o,k]| (act' <= 1), formed here from the device
// scalar *dy_amax and the head weights every lane already holds, and published through *bound_out for the consumers.
struct HeadPairOut {
  __half *hi, *lo;
  int64_t ld;              // halves
  const float* dy_amax;    // device scalar: max |dY|
  float* bound_out;        // device scalar written by block 0
};

template <int NO, int KV, bool PAIR>
__global__ void __launch_bounds__(kHeadThreads) head_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ H,
                                                                int64_t ldh, const float* __restrict__ W, int act,
                                                                float* __restrict__ dH, int64_t lddh, int M,
                                                                float* __restrict__ partial /*[blocks][NO*K + NO + K]*/,
                                                                const HeadPairOut po) {
  constexpr int K = 128 * KV;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 w[NO][KV], gw[NO][KV], gd[KV];
  float gb = 0.f;  // lane o: column sum of dY[:, o]
#pragma unroll
  for (int v = 0; v < KV; ++v) gd[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int o = 0; o < NO; ++o) {
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      w[o][v] = __ldg(reinterpret_cast<const float4*>(W + o * K + 128 * v + 4 * lane));
      gw[o][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float s_out = 1.f;
  if (PAIR) {
    float col = 0.f;  // max over this lane's columns of sum_o |W[o,k]|
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NO; ++o) a.x += fabsf(w[o][v].x), a.y += fabsf(w[o][v].y), a.z += fabsf(w[o][v].z), a.w += fabsf(w[o][v].w);
      col = fmaxf(col, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) col = fmaxf(col, __shfl_xor_sync(0xffffffffu, col, o));
    const float bound = __ldg(po.dy_amax) * col * 1.001f;
    s_out = f16x3_scale(bound);
    if (blockIdx.x == 0 && threadIdx.x == 0) *po.bound_out = bound;
  }
  // software pipeline: the loads of the next 8 rows are issued before the arithmetic on the current 8
  float4 h[kHeadRows][KV], hn[kHeadRows][KV];
  float gl[kHeadRows], gn[kHeadRows];
  auto load_rows = [&](int64_t base, float4 (&hh)[kHeadRows][KV], float (&gg)[kHeadRows]) {
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const int64_t row = base + r;  // rows >= M contribute zeros and are not stored
#pragma unroll
      for (int v = 0; v < KV; ++v)
        hh[r][v] = row < M ? ldg_stream4(H + row * ldh + 128 * v + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      gg[r] = (row < M && lane < NO) ? __ldg(dY + row * NO + lane) : 0.f;
    }
  };
  const int64_t stride = (int64_t)nwarps * kHeadRows;
  load_rows((int64_t)warp * kHeadRows, h, gl);
  for (int64_t base = (int64_t)warp * kHeadRows; base < M; base += stride) {
    load_rows(base + stride, hn, gn);
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const int64_t row = base + r;
      float4 d[KV];
#pragma unroll
      for (int v = 0; v < KV; ++v) d[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const float g = __shfl_sync(0xffffffffu, gl[r], o);
#pragma unroll
        for (int v = 0; v < KV; ++v) {
          d[v].x += g * w[o][v].x, d[v].y += g * w[o][v].y, d[v].z += g * w[o][v].z, d[v].w += g * w[o][v].w;
          gw[o][v].x += g * h[r][v].x, gw[o][v].y += g * h[r][v].y, gw[o][v].z += g * h[r][v].z, gw[o][v].w += g * h[r][v].w;
        }
      }
      if ((PAIR || dH) && row < M) {
#pragma unroll
        for (int v = 0; v < KV; ++v) {
          d[v].x *= head_act_grad(h[r][v].x, act), d[v].y *= head_act_grad(h[r][v].y, act);
          d[v].z *= head_act_grad(h[r][v].z, act), d[v].w *= head_act_grad(h[r][v].w, act);
          if (PAIR) {
            const float a0 = d[v].x * s_out, a1 = d[v].y * s_out, a2 = d[v].z * s_out, a3 = d[v].w * s_out;
            const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
            const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(a0 - b01.x, a1 - b01.y), l23 = __floats2half2_rn(a2 - b23.x, a3 - b23.y);
            uint2 uh, ul;
            uh.x = *reinterpret_cast<const uint32_t*>(&h01), uh.y = *reinterpret_cast<const uint32_t*>(&h23);
            ul.x = *reinterpret_cast<const uint32_t*>(&l01), ul.y = *reinterpret_cast<const uint32_t*>(&l23);
            *reinterpret_cast<uint2*>(po.hi + row * po.ld + 128 * v + 4 * lane) = uh;
            *reinterpret_cast<uint2*>(po.lo + row * po.ld + 128 * v + 4 * lane) = ul;
          } else {
            *reinterpret_cast<float4*>(dH + row * lddh + 128 * v + 4 * lane) = d[v];
          }
          gd[v].x += d[v].x, gd[v].y += d[v].y, gd[v].z += d[v].z, gd[v].w += d[v].w;
        }
      }
      gb += gl[r];
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      gl[r] = gn[r];
#pragma unroll
      for (int v = 0; v < KV; ++v) h[r][v] = hn[r][v];
    }
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
  // bias gradient: lane o of every warp holds its warp's share of column o
  __shared__ float redb[kHeadThreads / 32][NO];
  if (lane < NO) redb[wib][lane] = gb;
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

// one WARP per output element: lanes stride over the block partials (fixed order -> deterministic), then a shuffle tree;
// one thread per element walked ~600 partials serially (16.5 us for 1 676 elements)
__global__ void head_bwd_final_kernel(const float* __restrict__ partial, int nblocks, int NO, int K, float* __restrict__ dW,
                                      float* __restrict__ db, int accumulate, float* __restrict__ dbH, int accumulate_dbh) {
  const int total = NO * K + NO + K;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= total) return;
  const bool trunk = i >= NO * K + NO;
  float* dst = i < NO * K ? dW + i : (trunk ? (dbH ? dbH + (i - NO * K - NO) : nullptr) : (db ? db + (i - NO * K) : nullptr));
  if (!dst) return;
  double acc = 0.0;
  for (int b = lane; b < nblocks; b += 32) acc += partial[(int64_t)b * total + i];
  acc = warp_sum(acc);
  if (lane == 0) {
    const int add = trunk ? accumulate_dbh : accumulate;
    *dst = add ? *dst + (float)acc : (float)acc;
  }
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
  HeadPairOut po;
  template <int NO, int KV>
  int run() {
    if (po.hi)
      head_bwd_kernel<NO, KV, true><<<grid, kHeadThreads, 0, s>>>(dY, H, ldh, W, act, dH, lddh, M, partial, po);
    else
      head_bwd_kernel<NO, KV, false><<<grid, kHeadThreads, 0, s>>>(dY, H, ldh, W, act, dH, lddh, M, partial, po);
    return check_launch("head_bwd_kernel");
  }
};

static unsigned head_grid(int64_t M, int blocks_per_sm, int64_t max_blocks) {
  const int64_t rows_per_block = (kHeadThreads / 32) * kHeadRows;
  int64_t blocks = (M + rows_per_block - 1) / rows_per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (cap > max_blocks) cap = max_blocks;
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
  HeadFwd f{H, W, bias, ldh, Y, (int)M, head_grid(M, 6, 1 << 20), (cudaStream_t)stream};
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
  HeadBwd f{dY, H, W, ldh, lddh, act, (int)M, dH, (float*)scratch, head_grid(M, 2, kHeadMaxBlocks), s, HeadPairOut{}};  // 176 registers: 2 blocks per SM
  if (int e = head_dispatch((int)No, (int)K, f)) return e;
  const int total = (int)(No * K + No + K);
  head_bwd_final_kernel<<<(total * 32 + 255) / 256, 256, 0, s>>>((const float*)scratch, (int)f.grid, (int)No, (int)K, dW, db,
                                                            accumulate, dbH, accumulate_dbh);
  return check_launch("head_bwd_final_kernel");
}

int cusrl_b200_head_bwd_f16pair(const float* dY, const float* dy_amax, const float* H, int64_t ldh, const float* W, int act,
                                uint16_t* dH_hi, uint16_t* dH_lo, int64_t lddh, float* dh_bound, float* dW, float* db, int64_t M,
                                int64_t K, int64_t No, int accumulate, float* dbH, int accumulate_dbh, void* scratch,
                                size_t scratch_bytes, void* stream) {
  CUSRL_REQUIRE(dY && dy_amax && H && W && dW && scratch && dH_hi && dH_lo && dh_bound, CUSRL_B200_EINVAL, "head_bwd_f16pair: null pointer");
  CUSRL_REQUIRE(M > 0 && K > 0 && No > 0 && M < (1ll << 31), CUSRL_B200_EINVAL, "head_bwd_f16pair: bad sizes");
  CUSRL_REQUIRE((K % 128) == 0 && No <= kHeadMaxNo && No * K <= 2048, CUSRL_B200_EUNSUPPORTED, "head_bwd_f16pair: unsupported head shape");
  CUSRL_REQUIRE((ldh % 4) == 0 && ldh >= K && (lddh % 8) == 0 && lddh >= K && aligned_to(H, 16) && aligned_to(W, 16) &&
                    aligned_to(dH_hi, 16) && aligned_to(dH_lo, 16),
                CUSRL_B200_EALIGN, "head_bwd_f16pair: alignment");
  CUSRL_REQUIRE(act >= 0 && act <= 2, CUSRL_B200_EINVAL, "head_bwd_f16pair: unknown activation code");
  CUSRL_REQUIRE(scratch_bytes >= cusrl_b200_head_bwd_scratch_bytes(K, No), CUSRL_B200_ESCRATCH, "head_bwd_f16pair: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  HeadBwd f{dY, H, W, ldh, 0, act, (int)M, nullptr, (float*)scratch, head_grid(M, 2, kHeadMaxBlocks), s,
            HeadPairOut{(__half*)dH_hi, (__half*)dH_lo, lddh, dy_amax, dh_bound}};
  if (int e = head_dispatch((int)No, (int)K, f)) return e;
  const int total = (int)(No * K + No + K);
  head_bwd_final_kernel<<<(total * 32 + 255) / 256, 256, 0, s>>>((const float*)scratch, (int)f.grid, (int)No, (int)K, dW, db,
                                                            accumulate, dbH, accumulate_dbh);
  return check_launch("head_bwd_final_kernel");
}

}  // extern "C"
