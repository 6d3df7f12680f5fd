// Operand preparation for the f16x3 dense-layer kernels (f16x3_common.cuh): exact range (amax) of an fp32 tensor, its split
// into the fp16 hi / lo pair, and the per-optimizer-step preparation of a weight matrix (pair, transposed pair, and the
// norms the analytic activation bounds are built from).  All HBM-bound streaming kernels.
#include "f16x3_common.cuh"

namespace cusrl_b200 {

// ---- amax ------------------------------------------------------------------------------------------------------------
// bound[0] = max |x| over a [rows, width] fp32 array with row pitch ld.  Non-negative floats order like their bit
// patterns, so the cross-block combine is an integer atomicMax; `bound` is zeroed by the launcher first.
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int width,
                                                   float* __restrict__ bound) {
  float m = 0.f;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (ld == width && aligned_to(x, 16)) {
    // dense rows: one flat array whatever the width (a [M, 1] value gradient is M consecutive floats, not M one-element rows)
    const int64_t n = rows * width, n4 = n / 4;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 v = ldg_stream4(x + 4 * i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nthreads) m = fmaxf(m, fabsf(ldg_stream(x + i)));
  } else {
    for (int64_t r = warp; r < rows; r += nwarps)
      for (int c = lane; c < width; c += 32) m = fmaxf(m, fabsf(ldg_stream(x + r * ld + c)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if (lane == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    // a NaN anywhere makes the comparison chain drop it; surface it explicitly so the consumer's scale guard trips
    atomicMax(reinterpret_cast<unsigned int*>(bound), __float_as_uint(m));
  }
}

// ---- split -----------------------------------------------------------------------------------------------------------
// hi / lo [rows, ldh] halves from x [rows, width] fp32 (pitch ld); columns width..ldh-1 are written as zeros so the rows are
// zero-padded TMA sources.  One thread per 8 consecutive columns: two 16-byte loads in, one 16-byte store per half out.
__device__ __forceinline__ void split8(const float (&x)[8], float s, uint4& h, uint4& l) {
  __half2* h2 = reinterpret_cast<__half2*>(&h);
  __half2* l2 = reinterpret_cast<__half2*>(&l);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = x[2 * k] * s, b = x[2 * k + 1] * s;
    h2[k] = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h2[k]);
    l2[k] = __floats2half2_rn(a - back.x, b - back.y);
  }
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int width,
                                                        const float* __restrict__ bound, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t ldh) {
  const float s = f16x3_scale(__ldg(bound));
  const int units = (int)(ldh >> 3);  // 8-column units per row
  const int64_t total = rows * units;
  const bool vec = (ld & 3) == 0 && aligned_to(x, 16);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / units;
    const int c = (int)(i - r * units) * 8;
    const float* src = x + r * ld + c;
    float v[8];
    if (vec && c + 8 <= width) {
      const float4 a = ldg_stream4(src), b = ldg_stream4(src + 4);
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c + k < width ? ldg_stream(src + k) : 0.f;
    }
    uint4 h, l;
    split8(v, s, h, l);
    *reinterpret_cast<uint4*>(hi + r * ldh + c) = h;
    *reinterpret_cast<uint4*>(lo + r * ldh + c) = l;
  }
}

// ---- weights ---------------------------------------------------------------------------------------------------------
// stats[4] = { max|W|, max_n sum_k |W[n,k]|, max_k sum_n |W[n,k]|, max|bias| }, zeroed by the launcher; combined across
// blocks with integer atomicMax (non-negative floats order like their bit patterns).  Blocks 0 .. row_blocks-1 take 8 rows
// each (one warp per row, lanes along the contiguous K axis); the remaining blocks take 32 columns each (8 row groups per
// column).  The L1 norms are rounded sums: consumers widen the bounds they build from them.
__device__ __forceinline__ void weight_stats_block(const float* __restrict__ w, int N, int K, const float* __restrict__ bias,
                                                   float* __restrict__ stats, int row_blocks, int block) {
  unsigned int* out = reinterpret_cast<unsigned int*>(stats);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (block < row_blocks) {
    const int n = block * 8 + warp;
    if (n >= N) return;
    float s = 0.f, amax = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float a = fabsf(__ldg(w + (int64_t)n * K + k));
      s += a, amax = fmaxf(amax, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (lane == 0) {
      atomicMax(out + WSTAT_AMAX, __float_as_uint(amax));
      atomicMax(out + WSTAT_ROW_L1, __float_as_uint(s));
      if (bias) atomicMax(out + WSTAT_BIAS_MAX, __float_as_uint(fabsf(__ldg(bias + n))));
    }
  } else {
    // 32 columns x 8 row groups per block: lanes along the contiguous K axis (128-byte segments), each thread sums every
    // 8th row of its column, the 8 partial sums meet in shared memory
    __shared__ float part[8][33];
    const int k = (block - row_blocks) * 32 + lane;
    float s0 = 0.f, s1 = 0.f;
    if (k < K) {
      int n = warp;
      for (; n + 8 < N; n += 16) s0 += fabsf(__ldg(w + (int64_t)n * K + k)), s1 += fabsf(__ldg(w + (int64_t)(n + 8) * K + k));
      if (n < N) s0 += fabsf(__ldg(w + (int64_t)n * K + k));
    }
    part[warp][lane] = s0 + s1;
    __syncthreads();
    if (warp == 0) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s += part[g][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
      if (lane == 0) atomicMax(out + WSTAT_COL_L1, __float_as_uint(s));
    }
  }
}

// hi / lo [N, ld] and transposed hi_t / lo_t [K, ldt] halves of W [N, K] with the scale of stats[WSTAT_AMAX]; the padding
// columns (K..ld-1, N..ldt-1) are zero-initialised by the caller when it allocates the matrices.
__device__ __forceinline__ void weight_split_block(const float* __restrict__ w, int N, int K, const float* __restrict__ stats,
                                                   __half* __restrict__ hi, __half* __restrict__ lo, int ld,
                                                   __half* __restrict__ hi_t, __half* __restrict__ lo_t, int ldt, int block, int nblocks) {
  const float s = f16x3_scale(__ldg(stats + WSTAT_AMAX));
  const int total = N * K;
  for (int i = block * blockDim.x + threadIdx.x; i < total; i += nblocks * blockDim.x) {
    const int n = i / K, k = i - n * K;
    const float x = w[i] * s;
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi[(int64_t)n * ld + k] = h;
    lo[(int64_t)n * ld + k] = l;
    if (hi_t) {
      hi_t[(int64_t)k * ldt + n] = h;
      lo_t[(int64_t)k * ldt + n] = l;
    }
  }
}

__global__ void __launch_bounds__(256) weight_stats_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ bias,
                                                           float* __restrict__ stats, int row_blocks) {
  weight_stats_block(w, N, K, bias, stats, row_blocks, (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) weight_split_f16_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ stats,
                                                               __half* __restrict__ hi, __half* __restrict__ lo, int ld,
                                                               __half* __restrict__ hi_t, __half* __restrict__ lo_t, int ldt) {
  weight_split_block(w, N, K, stats, hi, lo, ld, hi_t, lo_t, ldt, (int)blockIdx.x, (int)gridDim.x);
}

// The same two passes for up to 8 weight matrices in ONE launch each (blockIdx.y = matrix): the layers of a network are
// re-split after every optimizer step, and three launches (+ a memset) per layer were ~0.3 ms of fixed cost per minibatch
// on the six dense layers of the PPO agent -- a quarter of a rank's step when 65536 environments are split over 8 GPUs.
constexpr int kPrepMax = 8;
struct WeightPrepMulti {
  const float* W[kPrepMax];
  const float* bias[kPrepMax];
  float* stats[kPrepMax];
  __half *hi[kPrepMax], *lo[kPrepMax], *hi_t[kPrepMax], *lo_t[kPrepMax];
  int N[kPrepMax], K[kPrepMax], ld[kPrepMax], ldt[kPrepMax], row_blocks[kPrepMax], stat_blocks[kPrepMax], split_blocks[kPrepMax];
  int count;
};
__global__ void __launch_bounds__(256) weight_stats_multi_kernel(const WeightPrepMulti p) {
  const int m = blockIdx.y;
  if ((int)blockIdx.x >= p.stat_blocks[m]) return;
  weight_stats_block(p.W[m], p.N[m], p.K[m], p.bias[m], p.stats[m], p.row_blocks[m], (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) weight_split_multi_kernel(const WeightPrepMulti p) {
  const int m = blockIdx.y;
  if ((int)blockIdx.x >= p.split_blocks[m]) return;
  weight_split_block(p.W[m], p.N[m], p.K[m], p.stats[m], p.hi[m], p.lo[m], p.ld[m], p.hi_t[m], p.lo_t[m], p.ldt[m], (int)blockIdx.x,
                     p.split_blocks[m]);
}
__global__ void weight_stats_zero_kernel(const WeightPrepMulti p) {
  if ((int)threadIdx.x < 4 * p.count) p.stats[threadIdx.x >> 2][threadIdx.x & 3] = 0.f;
}

namespace tc {

typedef CUresult (*EncodeTiledFn16)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn16 get_encoder16() {
  static EncodeTiledFn16 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn16>(ptr);
  }
  return fn;
}

int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
                       uint32_t box_outer, int swizzle) {
  EncodeTiledFn16 enc = get_encoder16();
  CUSRL_REQUIRE(enc != nullptr, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == TMAP_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUSRL_REQUIRE(r == CUDA_SUCCESS, CUSRL_B200_EDRIVER, "cuTensorMapEncodeTiled (f16) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tc
}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_amax_f32(const float* x, int64_t ld, int64_t rows, int64_t width, float* bound, void* stream) {
  CUSRL_REQUIRE(x && bound, CUSRL_B200_EINVAL, "amax: null pointer");
  CUSRL_REQUIRE(rows > 0 && width > 0 && ld >= width && width < (1ll << 30), CUSRL_B200_EINVAL, "amax: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t me = cudaMemsetAsync(bound, 0, sizeof(float), s);
  CUSRL_REQUIRE(me == cudaSuccess, (int)me, "amax: cudaMemsetAsync: %s", cudaGetErrorString(me));
  int64_t blocks = (rows * width / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  amax_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, ld, rows, (int)width, bound);
  return check_launch("amax_kernel");
}

int cusrl_b200_split_f16(const float* x, int64_t ld, int64_t rows, int64_t width, const float* bound, uint16_t* hi, uint16_t* lo,
                         int64_t ldh, void* stream) {
  CUSRL_REQUIRE(x && bound && hi && lo, CUSRL_B200_EINVAL, "split_f16: null pointer");
  CUSRL_REQUIRE(rows > 0 && width > 0 && ld >= width && ldh >= width && (ldh % 8) == 0 && ldh < (1ll << 30), CUSRL_B200_EINVAL,
                "split_f16: bad sizes (ldh must be a multiple of 8 halves covering the row)");
  CUSRL_REQUIRE(aligned_to(hi, 16) && aligned_to(lo, 16), CUSRL_B200_EALIGN, "split_f16: outputs must be 16-byte aligned");
  int64_t blocks = (rows * (ldh / 8) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, (int)width, bound, (__half*)hi, (__half*)lo, ldh);
  return check_launch("split_f16_kernel");
}

int cusrl_b200_weight_prep_f16(const float* W, int64_t N, int64_t K, const float* bias, uint16_t* hi, uint16_t* lo, int64_t ld,
                               uint16_t* hi_t, uint16_t* lo_t, int64_t ldt, float* stats, void* stream) {
  CUSRL_REQUIRE(W && hi && lo && stats, CUSRL_B200_EINVAL, "weight_prep_f16: null pointer");
  CUSRL_REQUIRE(N > 0 && K > 0 && ld >= K && (ld % 8) == 0 && N * K < (1ll << 31), CUSRL_B200_EINVAL, "weight_prep_f16: bad sizes");
  CUSRL_REQUIRE((hi_t == nullptr) == (lo_t == nullptr) && (!hi_t || (ldt >= N && (ldt % 8) == 0)), CUSRL_B200_EINVAL,
                "weight_prep_f16: transposed outputs must be given together with ldt >= N, a multiple of 8");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t me = cudaMemsetAsync(stats, 0, 4 * sizeof(float), s);
  CUSRL_REQUIRE(me == cudaSuccess, (int)me, "weight_prep_f16: cudaMemsetAsync: %s", cudaGetErrorString(me));
  const int row_blocks = (int)((N + 7) / 8), col_blocks = (int)((K + 31) / 32);
  weight_stats_kernel<<<row_blocks + col_blocks, 256, 0, s>>>(W, (int)N, (int)K, bias, stats, row_blocks);
  if (int e = check_launch("weight_stats_kernel")) return e;
  // the padding columns (K..ld-1 of the pair, N..ldt-1 of the transposed pair) are never written: the CALLER zero-initialises
  // the four matrices once, when it allocates them
  int blocks = (int)((N * K + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  weight_split_f16_kernel<<<blocks, 256, 0, s>>>(W, (int)N, (int)K, stats, (__half*)hi, (__half*)lo, (int)ld, (__half*)hi_t,
                                                 (__half*)lo_t, (int)ldt);
  return check_launch("weight_split_f16_kernel");
}

int cusrl_b200_weight_prep_f16_multi(int64_t count, const float* const* W, const int64_t* N, const int64_t* K, const float* const* bias,
                                     uint16_t* const* hi, uint16_t* const* lo, const int64_t* ld, uint16_t* const* hi_t,
                                     uint16_t* const* lo_t, const int64_t* ldt, float* const* stats, void* stream) {
  CUSRL_REQUIRE(count >= 1 && count <= kPrepMax && W && N && K && hi && lo && ld && stats, CUSRL_B200_EINVAL,
                "weight_prep_f16_multi: 1..8 matrices");
  WeightPrepMulti p{};
  int max_stat = 0, max_split = 0;
  for (int m = 0; m < (int)count; ++m) {
    CUSRL_REQUIRE(W[m] && hi[m] && lo[m] && stats[m], CUSRL_B200_EINVAL, "weight_prep_f16_multi: null pointer");
    CUSRL_REQUIRE(N[m] > 0 && K[m] > 0 && ld[m] >= K[m] && (ld[m] % 8) == 0 && N[m] * K[m] < (1ll << 31), CUSRL_B200_EINVAL,
                  "weight_prep_f16_multi: bad sizes");
    const bool t = hi_t && hi_t[m];
    CUSRL_REQUIRE(!t || (lo_t && lo_t[m] && ldt && ldt[m] >= N[m] && (ldt[m] % 8) == 0), CUSRL_B200_EINVAL,
                  "weight_prep_f16_multi: transposed outputs must be given together with ldt >= N, a multiple of 8");
    p.W[m] = W[m], p.bias[m] = bias ? bias[m] : nullptr, p.stats[m] = stats[m];
    p.hi[m] = (__half*)hi[m], p.lo[m] = (__half*)lo[m], p.hi_t[m] = t ? (__half*)hi_t[m] : nullptr, p.lo_t[m] = t ? (__half*)lo_t[m] : nullptr;
    p.N[m] = (int)N[m], p.K[m] = (int)K[m], p.ld[m] = (int)ld[m], p.ldt[m] = t ? (int)ldt[m] : 0;
    p.row_blocks[m] = (int)((N[m] + 7) / 8);
    p.stat_blocks[m] = p.row_blocks[m] + (int)((K[m] + 31) / 32);
    int sb = (int)((N[m] * K[m] + 255) / 256);
    p.split_blocks[m] = sb > 1184 ? 1184 : sb;
    max_stat = p.stat_blocks[m] > max_stat ? p.stat_blocks[m] : max_stat;
    max_split = p.split_blocks[m] > max_split ? p.split_blocks[m] : max_split;
  }
  p.count = (int)count;
  cudaStream_t s = (cudaStream_t)stream;
  weight_stats_zero_kernel<<<1, 32, 0, s>>>(p);
  if (int e = check_launch("weight_stats_zero_kernel")) return e;
  weight_stats_multi_kernel<<<dim3((unsigned)max_stat, (unsigned)count), 256, 0, s>>>(p);
  if (int e = check_launch("weight_stats_multi_kernel")) return e;
  weight_split_multi_kernel<<<dim3((unsigned)max_split, (unsigned)count), 256, 0, s>>>(p);
  return check_launch("weight_split_multi_kernel");
}

}  // extern "C"
