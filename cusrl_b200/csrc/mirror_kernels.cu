// Symmetry transforms: out[r, v, j] = x[r, dest[v][j]] * mult[v][j].
//
// Reference: MirrorDef.__call__ (cusrl/hook/auxiliary/symmetry.py:58-61: `input[..., destination] * multiplier`) and the
// tensors the symmetry hooks build from it -- `_build_mirrored` (:84-95, [V, N, C]) and `_build_augmented_tensor`
// (:334-339, torch.cat([original.unsqueeze(1), mirrored.movedim(0, 1)], dim=1) = [N, 1 + V, C]).  The reference runs one
// fancy-index gather, one multiply, one movedim copy and one cat per tensor per environment step; here one launch writes the
// final layout (the identity variant is just another table row), straight into 16-byte-padded rows when the destination
// is a rollout-buffer slot.  The multiply is a round-to-nearest fp32 product by +-1, i.e. bit-identical (including -0).
//
// HBM-bound: 4 B read + 4 V B written per element; every source row is read from DRAM once (its V re-reads hit L1).
#include "common.cuh"

namespace cusrl_b200 {

__global__ void __launch_bounds__(256) mirror_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int width,
                                                          const int* __restrict__ dest, const float* __restrict__ mult,
                                                          int variants, float* __restrict__ out, int64_t stride_r,
                                                          int64_t stride_v, int pad_to) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* src = x + r * ldx;
    for (int v = 0; v < variants; ++v) {
      float* dst = out + r * stride_r + v * stride_v;
      const int* d = dest + (int64_t)v * width;
      const float* m = mult + (int64_t)v * width;
      for (int j = lane; j < width; j += 32) dst[j] = __fmul_rn(__ldg(src + __ldg(d + j)), __ldg(m + j));
      for (int j = width + lane; j < pad_to; j += 32) dst[j] = 0.f;  // padding columns of a 16-byte-padded row
    }
  }
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_mirror_rows_f32(const float* x, int64_t ldx, int64_t rows, int64_t width, const int32_t* dest, const float* mult,
                               int64_t variants, float* out, int64_t stride_r, int64_t stride_v, int64_t pad_to, void* stream) {
  CUSRL_REQUIRE(x && dest && mult && out, CUSRL_B200_EINVAL, "mirror_rows: null pointer");
  CUSRL_REQUIRE(rows > 0 && width > 0 && width < (1ll << 24) && ldx >= width && variants > 0 && variants <= 64,
                CUSRL_B200_EINVAL, "mirror_rows: bad sizes (1..64 variants)");
  CUSRL_REQUIRE(pad_to == 0 || pad_to >= width, CUSRL_B200_EINVAL, "mirror_rows: pad_to must be 0 or >= width");
  CUSRL_REQUIRE(stride_v >= (pad_to ? pad_to : width) || variants == 1, CUSRL_B200_EINVAL,
                "mirror_rows: variant stride smaller than a row");
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  mirror_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, (int)width, dest, mult, (int)variants, out,
                                                                         stride_r, stride_v, (int)pad_to);
  return check_launch("mirror_rows_kernel");
}

}  // extern "C"
