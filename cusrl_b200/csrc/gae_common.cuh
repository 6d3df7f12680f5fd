// Parameter block shared by the GAE scan kernels (rollout_kernels.cu: register-resident LDG variant;
// gae_tma.cu: TMA-staged shared-memory variant).
#pragma once
#include "common.cuh"

namespace cusrl_b200 {

struct GaeParams {
  const float* reward;
  const uint8_t* done;        // !FUSED: done flags; FUSED: terminated
  const uint8_t* truncated;   // FUSED only
  const float* value;
  const float* next_value;    // !FUSED only
  const float* boot;          // FUSED only, [N*Dv]
  float* next_value_out;      // FUSED, optional
  float* advantage;
  float* ret;                 // optional
  int64_t T, N, Dv;
  float gamma, c_adv, c_ret;  // f32(gamma), f32(gamma*lamda), f32(gamma*lamda_value)
  float termination_value;
  int two_lambda;
};

// gae_tma.cu.  `launch_gae_tma` returns CUSRL_B200_EUNSUPPORTED (without touching the last-error string) when the
// problem does not meet the TMA variant's layout requirements; the caller then uses the LDG kernel.
struct GaeTmaConfig {
  int warps;        // tile width = 32 * warps columns; 0 = choose per problem size
  int stages;       // shared-memory stages per CTA (tiles in flight)
  int ctas_per_sm;  // resident CTAs per SM the grid is sized for
};
int launch_gae_tma(const GaeParams& p, const GaeTmaConfig& cfg, cudaStream_t s);

}  // namespace cusrl_b200
