/*
 * cusrl_b200.h -- C ABI of the B200-native (sm_100a) on-policy PPO hot path for CusRL.
 *
 * The reference (chengruiz/cusrl) is pure Python/PyTorch: it has no FFI of its own.  The drop-in
 * boundary is therefore its Python plugin surface (Hook / Sampler / ModuleFactory, see DESIGN.md and
 * INTEGRATION.md); THIS header is the C-ABI layer underneath that surface (SURVEY.md section 8b,
 * "C-ABI layer underneath").  Each entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer; nothing is allocated, freed or retained by the library;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); no call synchronises
 *     the device or the host;
 *   - return value: 0 on success, a negative CUSRL_B200_E* code for argument errors, or a positive
 *     cudaError_t when the launch itself failed.  cusrl_b200_last_error() gives a message for the
 *     calling thread.  Nothing throws, nothing exits;
 *   - tensors are dense row-major with the reference's layouts: rollout leaves are time-major
 *     [T, N, C] (template/buffer.py:144), minibatch leaves are [B, C];
 *   - bool leaves (terminated / truncated / done) are passed as uint8 (torch.bool storage);
 *   - compute entry points are stateless and re-entrant: safe to call from several host threads on
 *     different streams.  The only process-wide state are the *_set_config / *_set_variant /
 *     *_set_schedule tuning knobs (they select between bit-identical kernels and are meant to be set
 *     once at start-up) and one-time per-kernel attribute setup, which assumes the reference's
 *     process model: ONE device per process (utils/config.py:37-38, cuda:{LOCAL_RANK}).
 */
#ifndef CUSRL_B200_H_
#define CUSRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUSRL_B200_ABI_VERSION 1

#define CUSRL_B200_OK 0
#define CUSRL_B200_EINVAL (-1)       /* null pointer, negative/zero size, bad hyper-parameter */
#define CUSRL_B200_EALIGN (-2)       /* pointer or leading dimension not aligned as documented */
#define CUSRL_B200_EUNSUPPORTED (-3) /* shape outside what the kernel supports               */
#define CUSRL_B200_ESCRATCH (-4)     /* scratch buffer too small                               */
#define CUSRL_B200_EDRIVER (-5)      /* CUDA driver entry point (TMA descriptor) unavailable  */

int cusrl_b200_abi_version(void);
const char* cusrl_b200_last_error(void);
/* Number of SMs of the current device (grid sizing); <0 on error. */
int cusrl_b200_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * K3  next_value construction -- replaces ValueComputation.pre_update, hook/on_policy/value.py:68-82
 *   nv[t]   = value[t+1]           (t < T-1)
 *   nv[T-1] = boot_value[n]        (critic(next_state[-1]) computed by the caller)
 *   nv[t,n] = termination_value    where terminated[t,n]
 *   nv[t,n] = trunc_value[t,n]     where truncated[t,n]   (trunc_value==NULL: value[t,n], i.e. the
 *                                   bootstrap_truncated_states=False branch, value.py:81-82)
 *   applied in that order.  value,next_value: [T,N,Dv] f32; flags: [T,N] u8; boot_value: [N,Dv]. */
int cusrl_b200_next_value_f32(const float* value, const uint8_t* terminated, const uint8_t* truncated,
                              const float* boot_value, const float* trunc_value, float* next_value,
                              int64_t T, int64_t N, int64_t Dv, float termination_value, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1  GAE backward-in-time scan + return -- replaces _generalized_advantage_estimation and
 *     GeneralizedAdvantageEstimation._compute_advantage_and_return, hook/on_policy/gae.py:8-20,85-110
 *   adv[t] = (reward[t] + next_value[t]*f32(gamma)) - value[t]
 *   adv[t] = adv[t] + ((done[t]?0:1) * f32(gamma*lamda)) * adv[t+1]      t = T-2..0, no FMA contraction
 *   ret[t] = value[t] + adv[t]            (lamda_value < 0)
 *   ret[t] = value[t] + adv_lv[t]         (lamda_value >= 0: second scan with lamda_value)
 *   Bit-exact with the reference's op order.  reward,value,next_value,advantage,ret: [T,N,Dv] f32;
 *   done: [T,N] u8 (broadcast over Dv).  `ret` may be NULL (advantage only). */
int cusrl_b200_gae_f32(const float* reward, const uint8_t* done, const float* value,
                       const float* next_value, float* advantage, float* ret, int64_t T, int64_t N,
                       int64_t Dv, double gamma, double lamda, double lamda_value, void* stream);

/* K1+K3 fused: next_value is formed on the fly from value/terminated/truncated/boot_value exactly as
 * cusrl_b200_next_value_f32 would (trunc_value==NULL branch only), done = terminated | truncated
 * (template/actor_critic.py:277).  `next_value_out` may be NULL (not materialised). */
int cusrl_b200_gae_fused_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated,
                             const float* value, const float* boot_value, float termination_value,
                             float* next_value_out, float* advantage, float* ret, int64_t T, int64_t N,
                             int64_t Dv, double gamma, double lamda, double lamda_value, void* stream);

/* Tuning knob for K1 (process-wide): columns per thread (1, 2 or 4) and threads per block
 * (multiple of 32, <= 128).  Results are bit-identical for every setting. */
/* K3 + K1 + K2 statistics in ONE launch (+ a one-block finalisation), Dv == 1: next_value formed on the fly and published
 * (next_value_out may be NULL), advantage / return exactly as cusrl_b200_gae_fused_f32 (bit-identical), and
 * mean_var[2] = [mean | unbiased variance] of the advantages (torch.var_mean(correction=1), advantage.py:110-111) from fp64
 * block partials -- the input of cusrl_b200_advantage_normalize_f32.  T must be one of 8, 12, 16, 24, 32
 * (cusrl_b200_gae_chain_supported); scratch: cusrl_b200_gae_chain_scratch_bytes(N) bytes, 8-byte aligned.
 * cusrl_b200_gae_set_chain_threads selects the CTA size (128 / 256 / 448; bit-identical advantages). */
int cusrl_b200_gae_chain_supported(int64_t T, int64_t Dv);
size_t cusrl_b200_gae_chain_scratch_bytes(int64_t N);
int cusrl_b200_gae_set_chain_threads(int threads);
int cusrl_b200_gae_chain_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated, const float* value,
                             const float* boot_value, float termination_value, float* next_value_out, float* advantage,
                             float* ret, int64_t T, int64_t N, double gamma, double lamda, double lamda_value,
                             float* mean_var, void* scratch, size_t scratch_bytes, void* stream);
int cusrl_b200_gae_set_config(int vec, int threads);
/* Instruction schedule of the register-resident K1 kernel (process-wide): 0 = chunked kernel with a run-time T,
 * 1 = exact-length kernel (T in {8, 12, 16, 24, 32}: every load issued before the first dependent instruction), other
 * T fall back to 0.  Bit-identical results. */
int cusrl_b200_gae_set_schedule(int schedule);
/* Kernel variant of cusrl_b200_gae_f32 (process-wide): 0 = register-resident scan fed by vector loads, 1 = TMA-staged
 * scan (column tiles of 32*warps environments x T steps moved by bulk-tensor loads/stores through `stages`
 * shared-memory stages, grid sized for `ctas_per_sm` resident CTAs; warps = 0 picks the tile width per problem).
 * Variant 1 needs Dv == 1, N % 16 == 0 and 16-byte aligned leaves, otherwise variant 0 runs.  Bit-identical results. */
int cusrl_b200_gae_set_variant(int variant, int warps, int stages, int ctas_per_sm);

/* ------------------------------------------------------------------------------------------------
 * K2  advantage statistics + normalisation -- replaces AdvantageNormalization.normalize_,
 *     hook/on_policy/advantage.py:108-115 (torch.var_mean with correction=1 over all but the last dim)
 *   stats:      mean_var[0:Dv] = mean, mean_var[Dv:2Dv] = unbiased variance of adv[E,Dv]
 *   normalize:  adv = (adv - mean) / sqrt(var + eps)      (true division, in place)
 *   Between the two calls the caller may merge mean_var across ranks (utils/distributed.py:175-183).
 *   scratch: cusrl_b200_advantage_stats_scratch_bytes(Dv) bytes, any content, 16-byte aligned. */
size_t cusrl_b200_advantage_stats_scratch_bytes(int64_t Dv);
int cusrl_b200_advantage_stats_f32(const float* advantage, int64_t E, int64_t Dv, float* mean_var,
                                   void* scratch, size_t scratch_bytes, void* stream);
int cusrl_b200_advantage_normalize_f32(float* advantage, int64_t E, int64_t Dv, const float* mean_var,
                                       float eps, void* stream);
/* Equal-weight cross-rank merge of W stacked [mean(Dv) | var(Dv)] rows (utils/distributed.py:175-183):
 *   mean_g = mean_r(mean_r);  var_g = mean_r(var_r + (mean_r - mean_g)^2).  gathered: [W, 2*Dv]. */
int cusrl_b200_merge_mean_var_f32(const float* gathered, int64_t W, int64_t Dv, float* mean_var,
                                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * K8  minibatch gather -- replaces MiniBatchSampler._sample `data.flatten(0,1)[indices]`,
 *     sampler/mini_batch_sampler.py:76,89, for several leaves in ONE launch.
 *   For field f and output row i:  dst_f[i, 0:row_bytes_f] = src_f[index[i], 0:row_bytes_f];
 *   bytes row_bytes_f..dst_stride_f of each destination row are zero-filled (TMA-friendly padding).
 *   Strides are in bytes.  src/dst row starts must be aligned to min(16, largest power of two
 *   dividing both strides and row_bytes).  `fields_host` is a HOST array (copied into the launch). */
typedef struct {
  const void* src;     /* [E rows] device */
  void* dst;           /* [n_index rows] device */
  int64_t row_bytes;   /* payload bytes per row */
  int64_t src_stride;  /* bytes between source rows (>= row_bytes) */
  int64_t dst_stride;  /* bytes between destination rows (>= row_bytes) */
} cusrl_b200_gather_field;
#define CUSRL_B200_MAX_GATHER_FIELDS 24
int cusrl_b200_gather_rows(const cusrl_b200_gather_field* fields_host, int n_fields,
                           const int64_t* index, int64_t n_index, int64_t n_src_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  fused PPO objective, forward + unit gradients, one launch over the minibatch -- replaces
 *     OnPolicyPreparation.objective (hook/on_policy/common.py:29-43) with NormalDist log_prob/entropy
 *     (nn/module/distribution.py:195-213), _ppo_surrogate_loss / PpoSurrogateLoss / EntropyLoss
 *     (hook/on_policy/ppo.py:10-18,50-55,82-84) and ValueLoss / _clipped_value_loss
 *     (hook/on_policy/value.py:85-89,121-137), plus the autograd gradients of their weighted sum.
 *   inputs  mean,action [B,A]; std [A] (state-independent StddevVector, distribution.py:232-245);
 *           logp_old, advantage [B,1]; ret, value_old, curr_value [B,Dv]
 *   per-sample outputs (any may be NULL): logp, entropy, logp_ratio, prob_ratio [B,1]
 *   losses[0..2]  = value_loss*w_v, surrogate_loss*w_s, entropy_loss*w_e   (already weighted)
 *   metrics[0..2] = mean|logp_ratio|, mean entropy, mean over B of sum_Dv curr_value
 *                   (agent.record in common.py:45-49, value.py:139-141)
 *   unit gradients d(loss_k)/d(.):  d_mean [B,A] (surrogate), d_std_surr [A], d_std_ent [A],
 *                   d_value [B,Dv] (value loss).  Any gradient pointer may be NULL.
 *   value_clip <= 0 selects plain MSE (value.py:131-133).  1 <= A <= 32, Dv >= 1.
 *   has_value=0 skips the value loss (curr_value/ret/value_old may be NULL).
 *   scratch: cusrl_b200_ppo_loss_scratch_bytes(A) bytes, 16-byte aligned. */
size_t cusrl_b200_ppo_loss_scratch_bytes(int64_t A);
int cusrl_b200_ppo_loss_f32(const float* mean, const float* std, const float* action,
                            const float* logp_old, const float* advantage, const float* ret,
                            const float* value_old, const float* curr_value, int64_t B, int64_t A,
                            int64_t Dv, int has_value, float clip_ratio, float w_surrogate,
                            float w_entropy, float w_value, float value_clip, float* logp,
                            float* entropy, float* logp_ratio, float* prob_ratio, float* losses,
                            float* metrics, float* d_mean, float* d_std_surr, float* d_std_ent,
                            float* d_value, void* scratch, size_t scratch_bytes, void* stream);
/* dz[i] = dy[i] * act'(z_i) computed from the stored post-activation y (act: 0 identity, 1 ELU, 2 ReLU). */
int cusrl_b200_act_grad_mul_f32(const float* dy, const float* y, float* dz, int64_t n, int act, void* stream);
/* x[i] *= *scale_dev (in place, n floats); used to apply an upstream autograd scalar. */
int cusrl_b200_scale_f32(float* x, int64_t n, const float* scale_dev, void* stream);

/* Diagonal-normal KL(old||new) statistics for OnPolicyStatistics.post_update
 * (hook/on_policy/stats.py:29-40, distribution.py:215-218).  All per-sample inputs are [E,A] or [E,1];
 * std_new [A].  out[0..2] = mean KL, mean(advantage*exp(logp_new-logp_old)), mean(std_new). */
size_t cusrl_b200_policy_stats_scratch_bytes(void);
int cusrl_b200_policy_stats_f32(const float* mean_old, const float* std_old, const float* mean_new,
                                const float* std_new, const float* action, const float* logp_old,
                                const float* advantage, int64_t E, int64_t A, float* out,
                                void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9  gradient-norm clip + Adam on a flat parameter arena -- replaces GradientClipping.pre_optim
 *     (hook/on_policy/gradient_clipping.py:56-75, torch clip_grad_norm_) and torch.optim.Adam.step
 *     (preset/optimizer.py:9-23).
 *   grad_sumsq: accumulates sum(g^2) of g[0:n] into *sumsq_dev (double; caller zeroes it first).
 *   clip_coef:  *norm_dev = sqrt(sumsq); *coef_dev = min(1, max_norm / (norm + 1e-6)).
 *   adam_step:  g' = g * (*coef_dev) (coef_dev==NULL: 1); m,v,p updated as torch.optim.Adam
 *               (amsgrad=False, maximize=False); weight_decay is the L2 (non-decoupled) form. */
int cusrl_b200_grad_sumsq_f32(const float* grad, int64_t n, double* sumsq_dev, void* stream);
int cusrl_b200_clip_coef_f32(const double* sumsq_dev, float max_norm, float* norm_dev, float* coef_dev,
                             void* stream);
int cusrl_b200_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t n, const float* coef_dev, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int64_t step, void* stream);
/* The same step with the learning rate (*lr_dev, f32) and the step count (*step_dev, i64, >= 1) read from device
 * memory: a CUDA graph that contains the optimizer step stays valid while both change between replays. */
int cusrl_b200_adam_step_dev_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                 int64_t n, const float* coef_dev, const float* lr_dev,
                                 const int64_t* step_dev, float beta1, float beta2, float eps,
                                 float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6  dense layers of the MLP actor/critic on tcgen05 tensor cores (TMA-fed, TMEM accumulators) --
 *     replaces nn.Linear + activation in Mlp.forward (nn/module/mlp.py:77-90) and its autograd.
 *   precision: 1 = single-pass TF32, 3 = 3xTF32 error-compensated (fp32-equivalent accuracy; the
 *              reference runs these layers as fp32 SGEMM).
 *   act:       0 = identity, 1 = ELU(alpha=1), 2 = ReLU.
 *   weight_prep:  hi = W & ~0x1fff (TF32-representable), lo = W - hi, written with leading dimension
 *                 ld (>= K); optionally the transposed copies hi_t/lo_t [K, ldt] used by the data
 *                 gradient.  Run once per optimizer step per layer.
 *   linear_fwd:   Y[M,N]  = act(X[M,K] @ W[N,K]^T + bias[N])       W given as (W_hi, W_lo) with ldw
 *   linear_dgrad: dX[M,K] = (dY[M,N] @ W[N,K]) * act'(Xact[M,K])   W given as transposed copies
 *                 (WT_hi, WT_lo) [K, ldwt]; Xact = the layer input (= post-activation output of the layer
 *                 below); Xact==NULL: no activation factor.
 *   Leading dimensions are in ELEMENTS and must be multiples of 4 (16-byte rows for TMA); N must be a
 *   multiple of 4; ragged K / M edges are zero-filled by TMA.  All pointers 16-byte aligned. */
/*                 db_below (optional): the bias gradient of the layer BELOW, i.e. the column sums of dX
 *                 (dX is that layer's dZ), produced by the epilogue while the tile is on chip and reduced
 *                 in a fixed order through `workspace` (cusrl_b200_dgrad_workspace_bytes(K));
 *                 accumulate != 0 adds to the existing db_below.  NULL: not computed, workspace unused.
 */
int cusrl_b200_weight_prep_f32(const float* W, int64_t N, int64_t K, float* hi, float* lo, int64_t ld,
                               float* hi_t, float* lo_t, int64_t ldt, void* stream);
int cusrl_b200_linear_fwd_tf32(const float* X, int64_t ldx, const float* W_hi, const float* W_lo,
                               int64_t ldw, const float* bias, float* Y, int64_t ldy, int64_t M,
                               int64_t N, int64_t K, int act, int precision, void* stream);
int cusrl_b200_linear_dgrad_tf32(const float* dY, int64_t lddy, const float* WT_hi, const float* WT_lo,
                                 int64_t ldwt, const float* Xact, int64_t ldxa, float* dX, int64_t lddx,
                                 int64_t M, int64_t N, int64_t K, int act, int precision, float* db_below,
                                 int accumulate, void* workspace, size_t workspace_bytes, void* stream);
size_t cusrl_b200_dgrad_workspace_bytes(int64_t K);

/*   linear_wgrad: dW[N,K] (+)= dZ[M,N]^T @ X[M,K];  db[N] (+)= column sums of dZ (db may be NULL).
 *                 Split-K over CTAs with a deterministic second-stage reduction through `workspace`
 *                 (cusrl_b200_wgrad_workspace_bytes).  accumulate != 0 adds to the existing dW / db
 *                 (flat gradient arena semantics), 0 overwrites.  lddz, ldx multiples of 4. */
size_t cusrl_b200_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int cusrl_b200_linear_wgrad_tf32(const float* dZ, int64_t lddz, const float* X, int64_t ldx, float* dW,
                                 int64_t lddw, float* db, int64_t M, int64_t N, int64_t K, int precision,
                                 int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Small-N output heads (actor mean_head, critic value_head; reference LinearFp32 / nn.Linear in
 * nn/module/distribution.py:56 and nn/module/critic.py:87-88), exact fp32 SIMT kernels, HBM-bound:
 *   head_fwd: Y[M,No] = H[M,K] @ W[No,K]^T + bias[No]
 *   head_bwd: dH[M,K] = (dY[M,No] @ W) * act'(H)   (dH may be NULL);
 *             dW[No,K] (+)= dY^T @ H;  db[No] (+)= column sums of dY (db may be NULL);
 *             dbH[K] (+)= column sums of dH = bias gradient of the trunk's last layer (optional, needs dH)
 *   K multiple of 128, No <= 16, No*K <= 2048.  W, dY, Y dense; ldh / lddh multiples of 4. */
int cusrl_b200_head_fwd_f32(const float* H, int64_t ldh, const float* W, const float* bias, float* Y,
                            int64_t M, int64_t K, int64_t No, void* stream);
size_t cusrl_b200_head_bwd_scratch_bytes(int64_t K, int64_t No);
int cusrl_b200_head_bwd_f32(const float* dY, const float* H, int64_t ldh, const float* W, int act, float* dH,
                            int64_t lddh, float* dW, float* db, int64_t M, int64_t K, int64_t No,
                            int accumulate, float* dbH, int accumulate_dbh, void* scratch, size_t scratch_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  Random Network Distillation arithmetic -- replaces the tensor code of RandomNetworkDistillation
 *     (hook/auxiliary/rnd.py:68-81); the two small MLPs run on the K6 dense-layer kernels.
 *   rnd_reward: r[m] = reward_scale * mean_d (target[m,d] - pred[m,d])^2;  reward[m, 0..Dr) += r[m]
 *               (rnd.py:72-74); rnd_reward [M] and mean_out (mean of r, the recorded metric) optional.
 *   mse:        loss = mean_{M*D} (pred - target)^2 (nn.MSELoss, rnd.py:80); d_pred = dloss/dpred (optional).
 *   scratch: cusrl_b200_rnd_scratch_bytes() bytes, 8-byte aligned. */
size_t cusrl_b200_rnd_scratch_bytes(void);
int cusrl_b200_rnd_reward_f32(const float* target, const float* pred, int64_t M, int64_t D, float reward_scale,
                              float* reward, int64_t Dr, float* rnd_reward, float* mean_out, void* scratch,
                              size_t scratch_bytes, void* stream);
int cusrl_b200_mse_f32(const float* pred, const float* target, int64_t M, int64_t D, float* loss, float* d_pred,
                       void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  LSTM cell arithmetic -- with the K6 GEMMs replaces nn.LSTM forward/backward as used by Rnn._forward_sequence
 *     (nn/module/rnn.py:62-97,264-299; episode segmentation of nn/utils/recurrent.py:160-272 expressed as an
 *     in-line reset of the state where done, the equivalence pinned by cusrl_test/nn/module/test_rnn.py:145-164).
 *   cell_fwd: gates = xp + hp (biases already included), torch gate order (i,f,g,o);
 *             c_t = f*c_in + i*g; h_t = o*tanh(c_t); gates (activated), c_t, h_t stored;
 *             c_next/h_next = c_t/h_t * (1 - done)  = the state entering step t+1 (NULL at the last step).
 *   cell_bwd: dh = dh_above + (1-done)*dh_rec; dc = (1-done)*dc_rec + dh*o*(1-tanh(c_t)^2);
 *             dgates = pre-activation gate gradients; dc_prev = dc*f (masked by the previous step's done THERE).
 *   All [Nb, H] / [Nb, 4H] dense except xp (row pitch ldxp) and dh_above (row pitch lddh); H % 4 == 0. */
int cusrl_b200_lstm_cell_fwd_f32(const float* xp, int64_t ldxp, const float* hp, const float* c_in, const uint8_t* done,
                                 float* gates, float* c_out, float* h_out, float* c_next, float* h_next, int64_t Nb,
                                 int64_t H, void* stream);
int cusrl_b200_lstm_cell_bwd_f32(const float* dh_above, int64_t lddh, const float* dh_rec, const float* dc_rec,
                                 const uint8_t* done, const float* gates, const float* c, const float* c_in, float* dgates,
                                 float* dc_prev, int64_t Nb, int64_t H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rollout side -- replaces what ActorCritic.act / ActorCritic.step do around the network
 *     forward passes (template/actor_critic.py:227-291) and Buffer.push (template/buffer.py:124-151: one indexed copy
 *     per leaf), writing straight into the time-major buffer slots of the current step.
 *   copy_rows_padded:   dst[r, :width] = src[r, :width], dst[r, width:ldd] = 0 for r < rows (row pitches lds / ldd in
 *                       floats): a dense [N, 235] observation into its 16-byte-padded slot.
 *   rollout_store_step: next_observation (and next_state) rows as above, reward [N, reward_dim], terminated, truncated and
 *                       done = terminated | truncated (actor_critic.py:277) into their slots, ONE launch.  Any group may
 *                       be omitted by passing NULL for its source.
 *   sample_logp:        Normal.rsample + Normal.log_prob summed over the action dim (nn/module/distribution.py:195-213)
 *                       from the mean the head kernel wrote, the state-independent std vector sigma[A] and the
 *                       standard-normal draw eps[N, A]:  std = sigma;  action = mean + eps * sigma (deterministic != 0:
 *                       action = mean);  logp = sum_d( -((a-mu)^2) / (2 sigma^2) - log(sigma) - log(sqrt(2 pi)) ),
 *                       torch's operation order, no FMA contraction.  mean / eps / std / action: dense [N, A], A <= 64. */
int cusrl_b200_copy_rows_padded_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t width,
                                    void* stream);
int cusrl_b200_rollout_store_step_f32(const float* next_obs, int64_t ld_next_obs, float* next_obs_slot, int64_t ld_next_obs_slot,
                                      int64_t obs_dim, const float* next_state, int64_t ld_next_state, float* next_state_slot,
                                      int64_t ld_next_state_slot, int64_t state_dim, const float* reward, float* reward_slot,
                                      int64_t reward_dim, const uint8_t* terminated, const uint8_t* truncated,
                                      uint8_t* terminated_slot, uint8_t* truncated_slot, uint8_t* done_slot, int64_t N,
                                      void* stream);
int cusrl_b200_sample_logp_f32(const float* mean, const float* sigma, const float* eps, int64_t N, int64_t A,
                               int deterministic, float* std_out, float* action_out, float* logp_out, void* stream);
/* step() of a recurrent agent (Module.reset_memory + the rollout buffer's copies of the recurrent memories,
 * nn/module/module.py, hook/on_policy/value.py:42-56, template/actor_critic.py:283-289): for `count` (1..4) dense fp32
 * [N, width] memory tensors, rows where done[n] are zeroed IN PLACE and the resulting rows are also written to dst_a[k] /
 * dst_b[k] (either array or any entry may be null).  Host arrays of device pointers; width a multiple of 4. */
int cusrl_b200_memory_reset_store_f32(float* const* mem, float* const* dst_a, float* const* dst_b, int64_t count, const uint8_t* done,
                                      int64_t N, int64_t width, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6, precision 2 ("f16x3") -- the same dense layers (nn.Linear + activation, nn/module/mlp.py:77-90, and their autograd)
 *     on tcgen05 kind::f16 with fp16 hi / lo SPLIT operands:  x s = hi + lo,  hi = fp16(x s),  lo = fp16(x s - hi),  s a
 *     per-tensor power of two derived from a device-resident upper bound of max|x|;
 *     A B^T = (Ah Bh^T + Ah Bl^T + Al Bh^T) / (sA sB) with fp32 accumulation: fp32-equivalent results at 1.5 TF32 passes
 *     instead of 3xTF32's three (accuracy and rationale: csrc/f16x3_common.cuh).
 *   A "pair" is two dense fp16 matrices [rows, ld] (ld a multiple of 8 halves, columns >= width zero) plus one float bound.
 *   amax:            bound[0] = max|x| of an fp32 [rows, width] array (row pitch ld floats) -- the exact bound of an input.
 *   split_f16:       the pair of x with the scale of *bound.
 *   weight_prep_f16: pair of W[N,K] (+ transposed pair [K, ldt]; the caller zero-initialises the padding columns once) and stats[4] = { max|W|, max_n sum_k|W_nk|,
 *                    max_k sum_n|W_nk|, max|bias| }: the factors of the ANALYTIC output bounds
 *                    bound(act(x W^T + b)) <= bound(x) stats[1] + stats[3]   and   bound((dz W) act') <= bound(dz) stats[2],
 *                    which the GEMM kernels compute and publish (y_bound / dx_bound) without any extra pass.
 *   linear_fwd_f16x3:   Y = act(X W^T + b); output EITHER fp32 (Y, ldy) OR a pair (Yhi, Ylo, ldyh, y_bound).  N % 4 == 0.
 *   linear_dgrad_f16x3: dX = (dY W) * act'(Xact) with W given as the transposed pair; Xact pair optional; output fp32 or
 *                       pair; db_below (+)= column sums of dX (optional; workspace as cusrl_b200_dgrad_workspace_bytes).
 *   linear_wgrad_f16x3: dW (+)= dZ^T X from the two pairs (reduction over the M batch rows). */
/* db[n] (+)= sum_m dZ[m, n]: the bias gradient of a dense layer from its fp32 pre-activation gradient (fixed-order, fp64
 * finalisation); workspace: cusrl_b200_colsum_workspace_bytes(N) bytes. */
size_t cusrl_b200_colsum_workspace_bytes(int64_t N);
int cusrl_b200_colsum_f32(const float* dZ, int64_t lddz, int64_t M, int64_t N, float* db, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream);
int cusrl_b200_amax_f32(const float* x, int64_t ld, int64_t rows, int64_t width, float* bound, void* stream);
int cusrl_b200_split_f16(const float* x, int64_t ld, int64_t rows, int64_t width, const float* bound, uint16_t* hi, uint16_t* lo,
                         int64_t ldh, void* stream);
int cusrl_b200_weight_prep_f16(const float* W, int64_t N, int64_t K, const float* bias, uint16_t* hi, uint16_t* lo, int64_t ld,
                               uint16_t* hi_t, uint16_t* lo_t, int64_t ldt, float* stats, void* stream);
/* cusrl_b200_weight_prep_f16 for `count` (1..8) matrices in three launches instead of 3 x count (host arrays of per-matrix
 * arguments; bias / hi_t / lo_t / ldt may be null or hold null entries): the layers of a network are re-split together
 * after every optimizer step. */
int cusrl_b200_weight_prep_f16_multi(int64_t count, const float* const* W, const int64_t* N, const int64_t* K, const float* const* bias,
                                     uint16_t* const* hi, uint16_t* const* lo, const int64_t* ld, uint16_t* const* hi_t,
                                     uint16_t* const* lo_t, const int64_t* ldt, float* const* stats, void* stream);
/* K8 for precision 2: pair[i] = split(src[index[i]]) for a wide fp32 leaf (row pitch lds floats, `width` valid columns): the
 * gathered minibatch rows emitted directly as the fp16 pair the first trunk layer consumes (scale of *bound, e.g. the amax
 * of the whole leaf); out-of-range indices are clamped like cusrl_b200_gather_rows. */
int cusrl_b200_gather_split_f16(const float* src, int64_t lds, const int64_t* index, int64_t n, int64_t n_src_rows, int64_t width,
                                const float* bound, uint16_t* hi, uint16_t* lo, int64_t ldh, void* stream);
/* cusrl_b200_head_bwd_f32 with dH emitted as the fp16 pair the f16x3 trunk kernels consume (no fp32 dH, no separate split
 * pass): scale from the analytic bound max|dY| * max_k sum_o|W[o,k]| built from the device scalar *dy_amax (cusrl_b200_amax_f32
 * of dY) and published in *dh_bound.  lddh in halves, a multiple of 8. */
int cusrl_b200_head_bwd_f16pair(const float* dY, const float* dy_amax, const float* H, int64_t ldh, const float* W, int act,
                                uint16_t* dH_hi, uint16_t* dH_lo, int64_t lddh, float* dh_bound, float* dW, float* db, int64_t M,
                                int64_t K, int64_t No, int accumulate, float* dbH, int accumulate_dbh, void* scratch,
                                size_t scratch_bytes, void* stream);
/* k-blocks the L2 prefetch of the HBM-streamed operand tiles runs ahead of the TMA loads (0 = off; process-wide knob). */
int cusrl_b200_f16x3_set_prefetch(int k_blocks);
/* output-tile width of the forward / data-gradient kernels: 0 = by problem (256 when N > 128), 128 or 256 = forced. */
int cusrl_b200_f16x3_set_tile(int bn);
int cusrl_b200_linear_fwd_f16x3(const uint16_t* Xhi, const uint16_t* Xlo, int64_t ldx, const float* x_bound, const uint16_t* Whi,
                                const uint16_t* Wlo, int64_t ldw, const float* w_stats, const float* bias, float* Y, int64_t ldy,
                                uint16_t* Yhi, uint16_t* Ylo, int64_t ldyh, float* y_bound, int64_t M, int64_t N, int64_t K,
                                int act, void* stream);
int cusrl_b200_linear_dgrad_f16x3(const uint16_t* dYhi, const uint16_t* dYlo, int64_t lddy, const float* dy_bound,
                                  const uint16_t* WThi, const uint16_t* WTlo, int64_t ldwt, const float* w_stats,
                                  const uint16_t* Xact_hi, const uint16_t* Xact_lo, int64_t ldxa, const float* xact_bound,
                                  float* dX, int64_t lddx32, uint16_t* dXhi, uint16_t* dXlo, int64_t lddx, float* dx_bound,
                                  int64_t M, int64_t N, int64_t K, int act, float* db_below, int accumulate, void* workspace,
                                  size_t workspace_bytes, void* stream);
size_t cusrl_b200_wgrad_f16x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
int cusrl_b200_linear_wgrad_f16x3(const uint16_t* dZhi, const uint16_t* dZlo, int64_t lddz, const float* dz_bound,
                                  const uint16_t* Xhi, const uint16_t* Xlo, int64_t ldx, const float* x_bound, float* dW,
                                  int64_t lddw, int64_t M, int64_t N, int64_t K, int accumulate, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Observation normalisation -- replaces the per-step arithmetic of RunningMeanStd
 *     (nn/layer/rms.py:121-214, nn/utils/normalization.py:15-49,78-93) used by ObservationNormalization
 *     (hook/mdp/observation.py:161-215):
 *   column_stats:  mean_var[0:C] = column means, mean_var[C:2C] = POPULATION variances (torch.var_mean(dim=0, correction=0))
 *                  of an fp32 [rows, C] array (row pitch ld); fp64 block partials, fixed-order finalisation.
 *                  scratch: cusrl_b200_column_stats_scratch_bytes(C) bytes, 8-byte aligned.
 *   rms_merge:     merge_mean_var_(mean, var, w_old, batch_mean, batch_var, w_new) in place (weights are host doubles, like
 *                  the reference's python ints), then std = sqrt(var + eps).
 *   rms_normalize: out = clamp((x - mean) / std, -clamp, clamp) (clamp <= 0: none); x / out may be pitched, zero_padding != 0
 *                  also zeroes out's columns C..ldo-1 (padded rollout-buffer rows). */
size_t cusrl_b200_column_stats_scratch_bytes(int64_t C);
int cusrl_b200_column_stats_f32(const float* x, int64_t ld, int64_t rows, int64_t C, float* mean_var, void* scratch,
                                size_t scratch_bytes, void* stream);
int cusrl_b200_rms_merge_f32(float* mean, float* var, float* std_, const float* batch_mean, const float* batch_var, int64_t C,
                             double w_old, double w_new, float eps, void* stream);
int cusrl_b200_rms_normalize_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int64_t C, const float* mean,
                                 const float* std_, float clamp, int zero_padding, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Symmetry transforms -- replaces MirrorDef.__call__ (hook/auxiliary/symmetry.py:58-61:
 *     input[..., destination] * multiplier) and the per-step gather + multiply + movedim + cat behind `_build_mirrored`
 *     (:84-95) and `_build_augmented_tensor` (:334-339):
 *   out[r * stride_r + v * stride_v + j] = x[r * ldx + dest[v * width + j]] * mult[v * width + j],  0 <= v < variants.
 *   An identity row in the tables makes variant 0 the original (the augmented [N, 1 + V, C] layout: stride_r = (1 + V) * ldo,
 *   stride_v = ldo); stride_v = rows * width, stride_r = width gives the stacked [V, N, C] layout.  pad_to > width also zeroes
 *   columns width..pad_to-1 of every output row (16-byte-padded rollout-buffer rows).  dest entries must lie in [0, width):
 *   the caller validates them (they are a property of the environment spec, checked once).  Bit-identical to the reference. */
int cusrl_b200_mirror_rows_f32(const float* x, int64_t ldx, int64_t rows, int64_t width, const int32_t* dest, const float* mult,
                               int64_t variants, float* out, int64_t stride_r, int64_t stride_v, int64_t pad_to, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7, sequence-resident LSTM layer (csrc/lstm_seq.cu) -- replaces nn.LSTM's per-step recurrence over a whole
 *     [T, Nb, .] sequence (cusrl/nn/module/rnn.py:62-97,264-299 driven by nn/utils/recurrent.py:160-272) in ONE launch:
 *   xp [T*Nb, 4H] (pitch ldxp) = input projection of every step incl. b_ih;  W_hh as the fp16 pair + statistics of
 *   cusrl_b200_weight_prep_f16 ([4H, H], pitch ldw halves);  b_hh [4H] nullable;  h0 / c0 [Nb, H] nullable (zeros);
 *   done [T, Nb] nullable: the state handed to step t+1 is zeroed where done[t] (in-line episode reset).
 *   Writes out = h_t [T,Nb,H] and (nullable) hin [T,Nb,H] = the hidden state that ENTERED each step, both row-major (they
 *   feed dense-layer calls), h_last / c_last = h_{T-1} / c_{T-1} (nullable, row pitch ld_last; h0 / c0 have row pitch ld0: a
 *   layer's slice of the reference's flat [N, layers * H] memory is read and written in place), and -- given together for
 *   training, all null for inference -- the tensors only cusrl_b200_lstm_seq_bwd_f32 reads --
 *   gates (activated i,f,g,o), cseq = c_t, cin = the cell state that entered each step -- in a PRIVATE tiled layout
 *   [T][ceil(Nb/128)][H/4][(4 gates)][128 rows][4 floats]: allocate T * ceil(Nb/128)*128 * 4H (gates) / * H (cseq, cin) floats.
 *   H must be a multiple of 64, at most 256 (cusrl_b200_lstm_seq_supported); workspace: cusrl_b200_lstm_seq_workspace_bytes
 *   bytes, 256-byte aligned, contents irrelevant (cleared by the call).  All CTAs of the launch must be co-resident: the grid
 *   is sized to the SM count; do not run it concurrently with another kernel that spins on it. */
int cusrl_b200_lstm_seq_supported(int64_t H);
/* timing experiments only: non-zero bits drop parts of the kernels' work (results are then wrong); default 0 */
int cusrl_b200_lstm_seq_set_debug(int bits);
size_t cusrl_b200_lstm_seq_workspace_bytes(int64_t T, int64_t Nb, int64_t H);
int cusrl_b200_lstm_seq_fwd_f32(const float* xp, int64_t ldxp, const uint16_t* Whi, const uint16_t* Wlo, int64_t ldw,
                                const float* w_stats, const float* b_hh, const float* h0, const float* c0, const uint8_t* done,
                                int64_t ld0, float* gates, float* cseq, float* out, float* hin, float* cin, float* h_last, float* c_last,
                                int64_t ld_last, int64_t T, int64_t Nb, int64_t H, void* workspace, size_t workspace_bytes,
                                void* stream);

/* Backward through time of the same layer in one launch: dgates [T,Nb,4H] = pre-activation gate gradients of every step,
 * from dout [T*Nb, H] (pitch lddo; gradient w.r.t. h_t from above) and the forward's saved gates / cseq / cin; the recurrent
 * term dgates_{t+1} @ W_hh is formed on the SMs from W_hh^T as the TRANSPOSED fp16 pair of cusrl_b200_weight_prep_f16
 * ([H, 4H], pitch ldwt halves).  done[t] cuts the gradient flowing from step t+1 into step t.  Deterministic.  The weight
 * gradients and the gradient w.r.t. the layer input follow as ordinary dense-layer calls on dgates. */
size_t cusrl_b200_lstm_seq_bwd_workspace_bytes(int64_t T, int64_t Nb, int64_t H);
int cusrl_b200_lstm_seq_bwd_f32(const float* dout, int64_t lddo, const float* gates, const float* cseq, const float* cin,
                                const uint8_t* done, const uint16_t* WThi, const uint16_t* WTlo, int64_t ldwt, const float* w_stats,
                                float* dgates, int64_t T, int64_t Nb, int64_t H, void* workspace, size_t workspace_bytes,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUSRL_B200_H_ */
