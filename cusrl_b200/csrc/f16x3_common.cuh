// fp16 hi/lo split ("f16x3") operands for the dense layers: x * s = hi + lo with hi = fp16(x s), lo = fp16(x s - hi), s a
// per-tensor power of two chosen from an UPPER BOUND of max|x| that lives in device memory, so that no value can overflow
// fp16 and everything within 2^-28 of the bound keeps ~22 significant bits (lo is stored unscaled: its gradual underflow
// costs an absolute error of 2^-40 of the bound, far below fp32 rounding of the dot products it enters).
//   A B^T = (Ah Bh^T + Ah Bl^T + Al Bh^T) / (sA sB)          three kind::f16 MMAs, fp32 accumulation in TMEM
// fp16 MMAs run at twice the TF32 rate, so the fp32-equivalent product costs 1.5 TF32 passes instead of 3xTF32's 3
// (measured against fp64 at M = 393 216: profiles/r02_fp16_split_probe.jsonl -- the split is as accurate as cuBLAS SGEMM on
// the forward shapes; the weight gradient's error is set by the accumulation chain, exactly as for 3xTF32).
// Bounds, not exact maxima, for tensors produced by a GEMM epilogue: bound(act(x W^T + b)) <= bound(x) max_n sum_k|W_nk| +
// max|b| is known BEFORE the kernel runs (all factors are device scalars), needs no extra pass, can never be exceeded,
// and its looseness (a few bits per layer) is irrelevant given the 28 bits of range below the bound.
#pragma once
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace cusrl_b200 {

// weight statistics written by weight_prep_f16 (float[4])
enum { WSTAT_AMAX = 0, WSTAT_ROW_L1 = 1, WSTAT_COL_L1 = 2, WSTAT_BIAS_MAX = 3 };

// power-of-two scale s with s * bound in (2^14, 2^15] (fp16 max is 65504): exact in fp32, no rounding anywhere
__host__ __device__ __forceinline__ float f16x3_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;  // all-zero tensor (or inf / nan: propagates regardless)
  int e;
  frexpf(bound, &e);  // bound = m 2^e, m in [0.5, 1)  ->  bound <= 2^e
  int k = 15 - e;
  k = k > 60 ? 60 : (k < -60 ? -60 : k);
  return ldexpf(1.f, k);
}

#ifdef __CUDACC__
namespace tc {

// Instruction descriptor for kind::f16 with F16 inputs and fp32 accumulation (cf. cute::UMMA::InstrDescriptor):
// c_format = F32 (bits 4-5 = 1), a_format / b_format = F16 (0) at bits 7-9 / 10-12, a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem], fp16 inputs (K = 16 per instruction), fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace tc
#endif

namespace tc {
// 2-D tensor map over a row-major [outer, inner] fp16 array (leading dimension ld_elems halves, a multiple of 8);
// swizzle: TMAP_SW128 or TMAP_SW64.
int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
                       uint32_t box_outer, int swizzle);
}  // namespace tc

}  // namespace cusrl_b200
