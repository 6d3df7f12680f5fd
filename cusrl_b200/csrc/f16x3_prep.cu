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
  if (ld == width && (width & 3) == 0 && aligned_to(x, 16)) {
    const int64_t n4 = rows * width / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = ldg_stream4(x + 4 * i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
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
// zero-padded TMA sources.  One warp per row, lanes stride 2-element units.
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int width,
                                                        const float* __restrict__ bound, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t ldh) {
  const float s = f16x3_scale(__ldg(bound));
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int units = (int)(ldh >> 1);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* src = x + r * ld;
    __half2* h2 = reinterpret_cast<__half2*>(hi + r * ldh);
    __half2* l2 = reinterpret_cast<__half2*>(lo + r * ldh);
    for (int u = lane; u < units; u += 32) {
      const int c = 2 * u;
      const float a = c < width ? ldg_stream(src + c) * s : 0.f;
      const float b = c + 1 < width ? ldg_stream(src + c + 1) * s : 0.f;
      const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
      h2[u] = __halves2half2(ha, hb);
      l2[u] = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
    }
  }
}

// ---- weights ---------------------------------------------------------------------------------------------------------
// stats[4] = { max|W|, max_n sum_k |W[n,k]|, max_k sum_n |W[n,k]|, max|bias| }  (one block; the matrices are <= ~130 k entries)
__global__ void __launch_bounds__(1024) weight_stats_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ bias,
                                                            float* __restrict__ stats) {
  __shared__ float red[3][32];
  float amax = 0.f, row_l1 = 0.f, col_l1 = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
      const float a = fabsf(w[(int64_t)n * K + k]);
      s += a, amax = fmaxf(amax, a);
    }
    row_l1 = fmaxf(row_l1, s);
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += fabsf(w[(int64_t)n * K + k]);
    col_l1 = fmaxf(col_l1, s);
  }
  float bmax = 0.f;
  if (bias)
    for (int n = threadIdx.x; n < N; n += blockDim.x) bmax = fmaxf(bmax, fabsf(bias[n]));
  float v[4] = {amax, row_l1, col_l1, bmax};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
  __shared__ float red4[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red4[i][warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = lane < nw ? red4[i][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
      // the sums above are rounded: widen the L1 norms by a few ulps so they stay upper bounds
      if (lane == 0) stats[i] = (i == WSTAT_ROW_L1 || i == WSTAT_COL_L1) ? t * 1.0001f : t;
    }
  }
  (void)red;
}

// hi / lo [N, ld] and transposed hi_t / lo_t [K, ldt] halves of W [N, K] with the scale of stats[WSTAT_AMAX]; the padding
// columns (K..ld-1, N..ldt-1) are zeroed by the launcher's memsets.
__global__ void __launch_bounds__(256) weight_split_f16_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ stats,
                                                               __half* __restrict__ hi, __half* __restrict__ lo, int ld,
                                                               __half* __restrict__ hi_t, __half* __restrict__ lo_t, int ldt) {
  const float s = f16x3_scale(__ldg(stats + WSTAT_AMAX));
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
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
  int64_t blocks = (rows + 7) / 8;
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
  weight_stats_kernel<<<1, 1024, 0, s>>>(W, (int)N, (int)K, bias, stats);
  if (int e = check_launch("weight_stats_kernel")) return e;
  if (ld > K) {  // zero the padding columns once per call (cheap: the matrices are tiny)
    cudaMemsetAsync(hi, 0, (size_t)N * ld * 2, s);
    cudaMemsetAsync(lo, 0, (size_t)N * ld * 2, s);
  }
  if (hi_t && ldt > N) {
    cudaMemsetAsync(hi_t, 0, (size_t)K * ldt * 2, s);
    cudaMemsetAsync(lo_t, 0, (size_t)K * ldt * 2, s);
  }
  int blocks = (int)((N * K + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  weight_split_f16_kernel<<<blocks, 256, 0, s>>>(W, (int)N, (int)K, stats, (__half*)hi, (__half*)lo, (int)ld, (__half*)hi_t,
                                                 (__half*)lo_t, (int)ldt);
  return check_launch("weight_split_f16_kernel");
}

}  // extern "C"
