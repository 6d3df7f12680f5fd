// K7: LSTM cell arithmetic (reference: nn.LSTM inside cusrl/nn/module/rnn.py:62-97, trained on episode-segmented
// sequences by cusrl/nn/utils/recurrent.py:160-272).  The matrix products run on the K6 tcgen05 kernels (input
// projection of all T steps as one GEMM, one recurrent GEMM per step); these kernels fuse everything per step:
//   forward : gates = xp + hp (biases already added by the GEMM epilogues) -> i,f,g,o -> c_t, h_t, and the RESET of the
//             state that enters step t+1 where done[t] (the in-line form of the reference's split / pad / scatter,
//             pinned by cusrl_test/nn/module/test_rnn.py:145-164)
//   backward: dh, dc -> pre-activation gate gradients and dc of the previous step (BPTT), with the same reset masks.
// Gate order is torch's (i, f, g, o); all tensors fp32; H multiple of 4.
#include "common.cuh"

namespace cusrl_b200 {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmFwdParams {
  const float* xp;  int64_t ldxp;   // [Nb, 4H] slice of the all-steps input projection
  const float* hp;                  // [Nb, 4H] recurrent projection of the (masked) previous hidden state
  const float* c_in;                // [Nb, H]  (masked) previous cell state
  const uint8_t* done;              // [Nb] done flags of THIS step (nullable): mask for the state handed to step t+1
  float* gates;                     // [Nb, 4H] activated gates (saved for backward)
  float* c_out;                     // [Nb, H]  c_t (unmasked)
  float* h_out;                     // [Nb, H]  h_t (unmasked)  -> output sequence
  float* c_next;                    // [Nb, H]  c_t * (1 - done)   (nullable: last step)
  float* h_next;                    // [Nb, H]  h_t * (1 - done)   (nullable)
  int Nb, H;
};

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const LstmFwdParams p) {
  const int H4 = p.H >> 2;
  const int64_t total = (int64_t)p.Nb * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int n = (int)(idx / H4), j = (int)(idx - (int64_t)n * H4) * 4;
    const float* xr = p.xp + (int64_t)n * p.ldxp;
    const float* hr = p.hp + (int64_t)n * 4 * p.H;
    float4 g4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xr + q * p.H + j));
      const float4 b = __ldg(reinterpret_cast<const float4*>(hr + q * p.H + j));
      g4[q] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    const float4 cp = *reinterpret_cast<const float4*>(p.c_in + (int64_t)n * p.H + j);
    float iv[4] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w}, fv[4] = {g4[1].x, g4[1].y, g4[1].z, g4[1].w};
    float gv[4] = {g4[2].x, g4[2].y, g4[2].z, g4[2].w}, ov[4] = {g4[3].x, g4[3].y, g4[3].z, g4[3].w};
    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
    float c[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      iv[k] = sigmoidf_(iv[k]), fv[k] = sigmoidf_(fv[k]), gv[k] = tanhf(gv[k]), ov[k] = sigmoidf_(ov[k]);
      c[k] = fv[k] * cpv[k] + iv[k] * gv[k];
      h[k] = ov[k] * tanhf(c[k]);
    }
    float* gr = p.gates + (int64_t)n * 4 * p.H;
    *reinterpret_cast<float4*>(gr + j) = make_float4(iv[0], iv[1], iv[2], iv[3]);
    *reinterpret_cast<float4*>(gr + p.H + j) = make_float4(fv[0], fv[1], fv[2], fv[3]);
    *reinterpret_cast<float4*>(gr + 2 * p.H + j) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    *reinterpret_cast<float4*>(gr + 3 * p.H + j) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    const int64_t o = (int64_t)n * p.H + j;
    *reinterpret_cast<float4*>(p.c_out + o) = make_float4(c[0], c[1], c[2], c[3]);
    *reinterpret_cast<float4*>(p.h_out + o) = make_float4(h[0], h[1], h[2], h[3]);
    if (p.c_next) {
      const float m = (p.done && p.done[n]) ? 0.f : 1.f;
      *reinterpret_cast<float4*>(p.c_next + o) = make_float4(c[0] * m, c[1] * m, c[2] * m, c[3] * m);
      *reinterpret_cast<float4*>(p.h_next + o) = make_float4(h[0] * m, h[1] * m, h[2] * m, h[3] * m);
    }
  }
}

struct LstmBwdParams {
  const float* dh_above;  int64_t lddh;  // [Nb, H] gradient w.r.t. h_t from the layer above / the head
  const float* dh_rec;                   // [Nb, H] dgates_{t+1} @ W_hh (nullable at the last step)
  const float* dc_rec;                   // [Nb, H] dc_{t+1} * f_{t+1}  (nullable at the last step)
  const uint8_t* done;                   // [Nb] done flags of THIS step: masks dh_rec / dc_rec (nullable)
  const float* gates;                    // [Nb, 4H] activated gates of this step
  const float* c;                        // [Nb, H] c_t
  const float* c_in;                     // [Nb, H] masked c_{t-1}
  float* dgates;                         // [Nb, 4H] pre-activation gate gradients
  float* dc_prev;                        // [Nb, H] dc_t * f_t   (unmasked; masked by the previous step's done there)
  int Nb, H;
};

__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const LstmBwdParams p) {
  const int H4 = p.H >> 2;
  const int64_t total = (int64_t)p.Nb * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int n = (int)(idx / H4), j = (int)(idx - (int64_t)n * H4) * 4;
    const int64_t o = (int64_t)n * p.H + j;
    const float m = (p.done && p.done[n]) ? 0.f : 1.f;
    float4 t = *reinterpret_cast<const float4*>(p.dh_above + (int64_t)n * p.lddh + j);
    float dh[4] = {t.x, t.y, t.z, t.w};
    float dc[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.dh_rec) {
      t = *reinterpret_cast<const float4*>(p.dh_rec + o);
      dh[0] += m * t.x, dh[1] += m * t.y, dh[2] += m * t.z, dh[3] += m * t.w;
      t = *reinterpret_cast<const float4*>(p.dc_rec + o);
      dc[0] = m * t.x, dc[1] = m * t.y, dc[2] = m * t.z, dc[3] = m * t.w;
    }
    const float* gr = p.gates + (int64_t)n * 4 * p.H;
    const float4 i4 = *reinterpret_cast<const float4*>(gr + j), f4 = *reinterpret_cast<const float4*>(gr + p.H + j);
    const float4 g4 = *reinterpret_cast<const float4*>(gr + 2 * p.H + j), o4 = *reinterpret_cast<const float4*>(gr + 3 * p.H + j);
    const float4 c4 = *reinterpret_cast<const float4*>(p.c + o), cp4 = *reinterpret_cast<const float4*>(p.c_in + o);
    const float iv[4] = {i4.x, i4.y, i4.z, i4.w}, fv[4] = {f4.x, f4.y, f4.z, f4.w};
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
    const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
    float di[4], df[4], dg[4], dov[4], dcp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float tc = tanhf(cv[k]);
      const float dct = dc[k] + dh[k] * ov[k] * (1.f - tc * tc);
      dov[k] = dh[k] * tc * ov[k] * (1.f - ov[k]);
      di[k] = dct * gv[k] * iv[k] * (1.f - iv[k]);
      df[k] = dct * cpv[k] * fv[k] * (1.f - fv[k]);
      dg[k] = dct * iv[k] * (1.f - gv[k] * gv[k]);
      dcp[k] = dct * fv[k];
    }
    float* dr = p.dgates + (int64_t)n * 4 * p.H;
    *reinterpret_cast<float4*>(dr + j) = make_float4(di[0], di[1], di[2], di[3]);
    *reinterpret_cast<float4*>(dr + p.H + j) = make_float4(df[0], df[1], df[2], df[3]);
    *reinterpret_cast<float4*>(dr + 2 * p.H + j) = make_float4(dg[0], dg[1], dg[2], dg[3]);
    *reinterpret_cast<float4*>(dr + 3 * p.H + j) = make_float4(dov[0], dov[1], dov[2], dov[3]);
    *reinterpret_cast<float4*>(p.dc_prev + o) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
  }
}

static unsigned lstm_grid(int64_t work) {
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace cusrl_b200

using namespace cusrl_b200;

extern "C" {

int cusrl_b200_lstm_cell_fwd_f32(const float* xp, int64_t ldxp, const float* hp, const float* c_in, const uint8_t* done,
                                 float* gates, float* c_out, float* h_out, float* c_next, float* h_next, int64_t Nb,
                                 int64_t H, void* stream) {
  CUSRL_REQUIRE(xp && hp && c_in && gates && c_out && h_out, CUSRL_B200_EINVAL, "lstm_cell_fwd: null pointer");
  CUSRL_REQUIRE((c_next == nullptr) == (h_next == nullptr), CUSRL_B200_EINVAL, "lstm_cell_fwd: c_next/h_next go together");
  CUSRL_REQUIRE(Nb > 0 && H > 0 && (H % 4) == 0 && (ldxp % 4) == 0 && ldxp >= 4 * H, CUSRL_B200_EINVAL,
                "lstm_cell_fwd: H and ldxp must be positive multiples of 4");
  CUSRL_REQUIRE(aligned_to(xp, 16) && aligned_to(hp, 16) && aligned_to(c_in, 16) && aligned_to(gates, 16) &&
                    aligned_to(c_out, 16) && aligned_to(h_out, 16) && (!c_next || (aligned_to(c_next, 16) && aligned_to(h_next, 16))),
                CUSRL_B200_EALIGN, "lstm_cell_fwd: pointers must be 16-byte aligned");
  LstmFwdParams p{xp, ldxp, hp, c_in, done, gates, c_out, h_out, c_next, h_next, (int)Nb, (int)H};
  lstm_cell_fwd_kernel<<<lstm_grid(Nb * (H / 4)), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("lstm_cell_fwd_kernel");
}

int cusrl_b200_lstm_cell_bwd_f32(const float* dh_above, int64_t lddh, const float* dh_rec, const float* dc_rec,
                                 const uint8_t* done, const float* gates, const float* c, const float* c_in, float* dgates,
                                 float* dc_prev, int64_t Nb, int64_t H, void* stream) {
  CUSRL_REQUIRE(dh_above && gates && c && c_in && dgates && dc_prev, CUSRL_B200_EINVAL, "lstm_cell_bwd: null pointer");
  CUSRL_REQUIRE((dh_rec == nullptr) == (dc_rec == nullptr), CUSRL_B200_EINVAL, "lstm_cell_bwd: dh_rec/dc_rec go together");
  CUSRL_REQUIRE(Nb > 0 && H > 0 && (H % 4) == 0 && (lddh % 4) == 0 && lddh >= H, CUSRL_B200_EINVAL,
                "lstm_cell_bwd: H and lddh must be positive multiples of 4");
  CUSRL_REQUIRE(aligned_to(dh_above, 16) && aligned_to(gates, 16) && aligned_to(c, 16) && aligned_to(c_in, 16) &&
                    aligned_to(dgates, 16) && aligned_to(dc_prev, 16) && (!dh_rec || (aligned_to(dh_rec, 16) && aligned_to(dc_rec, 16))),
                CUSRL_B200_EALIGN, "lstm_cell_bwd: pointers must be 16-byte aligned");
  LstmBwdParams p{dh_above, lddh, dh_rec, dc_rec, done, gates, c, c_in, dgates, dc_prev, (int)Nb, (int)H};
  lstm_cell_bwd_kernel<<<lstm_grid(Nb * (H / 4)), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("lstm_cell_bwd_kernel");
}

}  // extern "C"
