"""CPU oracle for the CusRL on-policy PPO hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU fp32 restatement of the reference's arithmetic for the path named in
BASELINE.json (rollout buffer -> next_value -> GAE -> advantage normalisation -> PPO / value /
entropy objective -> RND -> MLP / LSTM actor-critic -> grad clip -> Adam).  Every function cites the
reference file:line it follows (paths relative to the reference checkout, ``cusrl/...``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product package ``cusrl_b200`` never does.

Parity pinning: the oracle is checked against (a) the reference's own known-answer tests
(cusrl_test/hook/on_policy/test_gae.py:8-31, test_ppo.py:8-32, test_advantage.py:37-48,
cusrl_test/sampler/test_mini_batch_sampler.py:8-90) and (b) golden vectors produced by importing the
live reference in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``),
see tests/test_oracle_golden.py; and (c) the reference's own functions and hooks run LIVE next to the
oracle on randomised inputs of many shapes (tests/test_oracle_live_reference.py, bit-identical; the
reference package travels as baseline/_ref).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import Tensor

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# =================================================================================================
# Rollout-side arithmetic
# =================================================================================================
def next_value_ref(
    value: Tensor,
    terminated: Tensor,
    truncated: Tensor,
    boot_value: Tensor,
    termination_value: float = 0.0,
    trunc_value: Tensor | None = None,
) -> Tensor:
    """hook/on_policy/value.py:68-82 (ValueComputation.pre_update).

    value [T,N,Dv]; terminated/truncated [T,N,1] bool; boot_value [N,Dv] = critic(next_state[-1]).
    ``trunc_value`` None selects the bootstrap_truncated_states=False branch (value.py:81-82).
    """
    nv = torch.empty_like(value)
    nv[:-1] = value[1:]                                    # value.py:68
    nv[-1] = boot_value                                    # value.py:69-70
    term = terminated.squeeze(-1)
    trunc = truncated.squeeze(-1)
    nv[term] = value.new_full([value.size(-1)], termination_value)   # value.py:71-72
    if trunc_value is None:
        nv[trunc] = value[trunc]                           # value.py:82
    else:
        nv[trunc] = trunc_value[trunc]                     # value.py:74-80
    return nv


def gae_ref(reward: Tensor, done: Tensor, value: Tensor, next_value: Tensor, gamma: float, lamda: float) -> Tensor:
    """hook/on_policy/gae.py:8-20 (_generalized_advantage_estimation); time-major [T,N,Dv]."""
    keep = done.logical_not()
    adv = reward + next_value * gamma - value              # gae.py:17
    coef = gamma * lamda                                   # python double product, gae.py:19
    for t in reversed(range(adv.size(0) - 1)):
        adv[t] += keep[t] * coef * adv[t + 1]
    return adv


def advantage_and_return_ref(
    reward: Tensor, done: Tensor, value: Tensor, next_value: Tensor, gamma: float, lamda: float,
    lamda_value: float | None = None,
) -> tuple[Tensor, Tensor]:
    """hook/on_policy/gae.py:85-110 (_compute_advantage_and_return)."""
    adv = gae_ref(reward, done, value, next_value, gamma, lamda)
    tail = adv if lamda_value is None else gae_ref(reward, done, value, next_value, gamma, lamda_value)
    return adv, value + tail


def merge_mean_var_ref(means: Tensor, variances: Tensor) -> tuple[Tensor, Tensor]:
    """utils/distributed.py:175-183 (reduce_mean_var_): equal-weight merge of per-rank stats [W,Dv]."""
    mean = means.mean(dim=0)
    var = (variances + (means - mean).square()).mean(dim=0)
    return mean, var


def normalize_advantage_ref(advantage: Tensor, rank_stats: tuple[Tensor, Tensor] | None = None) -> Tensor:
    """hook/on_policy/advantage.py:108-115 (AdvantageNormalization.normalize_), out of place.

    ``rank_stats`` = (means [W,Dv], vars [W,Dv]) of ALL ranks (this rank included) emulates the
    synchronised branch; None is the single-process path.
    """
    dims = tuple(range(advantage.ndim - 1))
    var, mean = torch.var_mean(advantage, dim=dims)        # unbiased, advantage.py:111
    if rank_stats is not None:
        mean, var = merge_mean_var_ref(*rank_stats)
    std = (var + 1e-8).sqrt()
    return (advantage - mean) / std


# =================================================================================================
# Distribution + objective
# =================================================================================================
def normal_log_prob_ref(mean: Tensor, std: Tensor, sample: Tensor) -> Tensor:
    """nn/module/distribution.py:207-209 -> torch Normal.log_prob summed over the action dim."""
    var = std.square()
    lp = -((sample - mean).square()) / (2 * var) - std.log() - LOG_SQRT_2PI
    return lp.sum(dim=-1, keepdim=True)


def normal_entropy_ref(std: Tensor) -> Tensor:
    """nn/module/distribution.py:211-213 -> torch Normal.entropy summed over the action dim."""
    return (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).sum(dim=-1, keepdim=True)


def normal_kl_ref(mean_p: Tensor, std_p: Tensor, mean_q: Tensor, std_q: Tensor) -> Tensor:
    """nn/module/distribution.py:215-218 -> torch kl_divergence(Normal, Normal), summed."""
    var_ratio = (std_p / std_q).square()
    t1 = ((mean_p - mean_q) / std_q).square()
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(dim=-1, keepdim=True)


def surrogate_loss_ref(advantage: Tensor, prob_ratio: Tensor, clip_ratio: float) -> Tensor:
    """hook/on_policy/ppo.py:10-18 (_ppo_surrogate_loss)."""
    unclipped = advantage * prob_ratio
    clipped = advantage * prob_ratio.clamp(1.0 - clip_ratio, 1.0 + clip_ratio)
    return -torch.min(unclipped, clipped).mean()


def value_loss_ref(value_old: Tensor, curr_value: Tensor, ret: Tensor, loss_clip: float | None) -> Tensor:
    """hook/on_policy/value.py:85-89,131-135 (MSE or clipped value loss, un-weighted)."""
    if loss_clip is None:
        return torch.nn.functional.mse_loss(ret, curr_value)
    clipped = value_old + (curr_value - value_old).clamp(-loss_clip, loss_clip)
    return torch.max((curr_value - ret).square(), (clipped - ret).square()).mean()


@dataclass
class ObjectiveOut:
    value_loss: Tensor
    surrogate_loss: Tensor
    entropy_loss: Tensor
    logp: Tensor
    entropy: Tensor
    logp_ratio: Tensor
    prob_ratio: Tensor
    d_mean: Tensor | None = None
    d_std: Tensor | None = None
    d_std_surr: Tensor | None = None
    d_std_ent: Tensor | None = None
    d_value: Tensor | None = None


def ppo_objective_ref(
    mean: Tensor, std_param: Tensor, action: Tensor, logp_old: Tensor, advantage: Tensor, ret: Tensor,
    value_old: Tensor, curr_value: Tensor, clip_ratio: float = 0.2, w_surrogate: float = 1.0,
    w_entropy: float = 0.01, w_value: float = 0.5, value_clip: float | None = None, with_grads: bool = True,
) -> ObjectiveOut:
    """One minibatch objective exactly as the hook chain evaluates it, with autograd gradients.

    hook/on_policy/value.py:121-137 (ValueLoss), common.py:29-43 (OnPolicyPreparation),
    ppo.py:50-55,82-84 (PpoSurrogateLoss, EntropyLoss), actor_critic.py:309 (python sum of the dict).
    ``std_param`` is the state-independent StddevVector parameter [A] (distribution.py:232-245).
    """
    mean = mean.detach().clone().requires_grad_(with_grads)
    std_param = std_param.detach().clone().requires_grad_(with_grads)
    curr_value = curr_value.detach().clone().requires_grad_(with_grads)
    std = std_param.repeat(*mean.shape[:-1], 1)            # distribution.py:245
    logp = normal_log_prob_ref(mean, std, action)
    entropy = normal_entropy_ref(std)
    logp_ratio = logp - logp_old                           # common.py:35
    prob_ratio = logp_ratio.exp()                          # common.py:41
    l_v = value_loss_ref(value_old, curr_value, ret, value_clip) * w_value
    l_s = surrogate_loss_ref(advantage, prob_ratio, clip_ratio) * w_surrogate
    l_e = -entropy.mean() * w_entropy                      # ppo.py:83-84
    out = ObjectiveOut(l_v.detach(), l_s.detach(), l_e.detach(), logp.detach(), entropy.detach(),
                       logp_ratio.detach(), prob_ratio.detach())
    if with_grads:
        g_mean, g_std_s = torch.autograd.grad(l_s, [mean, std_param], retain_graph=True)
        (g_std_e,) = torch.autograd.grad(l_e, [std_param], retain_graph=True)
        (g_val,) = torch.autograd.grad(l_v, [curr_value])
        out.d_mean, out.d_std_surr, out.d_std_ent, out.d_value = g_mean, g_std_s, g_std_e, g_val
        out.d_std = g_std_s + g_std_e
    return out


def policy_stats_ref(mean_old, std_old, mean_new, std_new, action, logp_old, advantage) -> tuple[Tensor, Tensor, Tensor]:
    """hook/on_policy/stats.py:29-40: mean KL(old||new), mean importance-weighted advantage, mean std."""
    kl = normal_kl_ref(mean_old, std_old, mean_new, std_new)
    logp = normal_log_prob_ref(mean_new, std_new, action)
    iwa = advantage * (logp - logp_old).exp()
    return kl.mean(), iwa.mean(), std_new.mean()


# =================================================================================================
# Sampler
# =================================================================================================
def minibatch_slices_ref(perm: Tensor, num_mini_batches: int) -> list[Tensor]:
    """sampler/mini_batch_sampler.py:66,76: B = E // k, contiguous slices, remainder dropped."""
    size = perm.numel() // num_mini_batches
    return [perm[k * size : (k + 1) * size] for k in range(num_mini_batches)]


def gather_ref(leaf: Tensor, indices: Tensor, temporal: bool = False) -> Tensor:
    """mini_batch_sampler.py:89 (`flatten(0,1)[idx]`) and :114 (`[:, idx]`)."""
    return leaf[:, indices] if temporal else leaf.flatten(0, 1)[indices]


# =================================================================================================
# Networks
# =================================================================================================
def activation_ref(x: Tensor, name: str) -> Tensor:
    if name == "ELU":
        return torch.nn.functional.elu(x)
    if name == "ReLU":
        return torch.relu(x)
    if name in ("Identity", "none"):
        return x
    raise ValueError(f"unsupported activation {name!r}")


def mlp_trunk_ref(params: dict[str, Tensor], prefix: str, x: Tensor, n_layers: int, activation: str) -> Tensor:
    """nn/module/mlp.py:77-90 with ends_with_activation=True (preset/ppo.py:137-140): layers.{0,2,4}."""
    h = x
    for i in range(n_layers):
        w, b = params[f"{prefix}backbone.layers.{2 * i}.weight"], params[f"{prefix}backbone.layers.{2 * i}.bias"]
        h = activation_ref(torch.nn.functional.linear(h, w, b), activation)
    return h


def actor_forward_ref(params: dict[str, Tensor], x: Tensor, n_layers: int, activation: str) -> tuple[Tensor, Tensor]:
    """nn/module/actor.py:181-205 + distribution.py:272-273: (mean [B,A], std [B,A])."""
    h = mlp_trunk_ref(params, "actor.", x, n_layers, activation)
    mean = torch.nn.functional.linear(h, params["actor.distribution.mean_head.weight"], params["actor.distribution.mean_head.bias"])
    std = params["actor.distribution.std.param"].repeat(*mean.shape[:-1], 1)
    return mean, std


def critic_forward_ref(params: dict[str, Tensor], x: Tensor, n_layers: int, activation: str) -> Tensor:
    """nn/module/critic.py:70-89: trunk + fp32 value head."""
    h = mlp_trunk_ref(params, "critic.", x, n_layers, activation)
    return torch.nn.functional.linear(h, params["critic.value_head.weight"], params["critic.value_head.bias"])


def init_mlp_params_ref(obs_dim: int, act_dim: int, hidden: tuple[int, ...], value_dim: int = 1,
                        generator: torch.Generator | None = None, init_std: float = 1.0) -> dict[str, Tensor]:
    """Default nn.Linear init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    weight and bias); orthogonal_init=False in the Anymal preset (zoo/isaaclab/locomotion.py:56).
    Parameter names follow the reference modules so reference checkpoints map 1:1."""
    params: dict[str, Tensor] = {}

    def linear(name: str, fan_in: int, fan_out: int):
        bound = 1.0 / math.sqrt(fan_in)
        params[f"{name}.weight"] = (torch.rand(fan_out, fan_in, generator=generator) * 2 - 1) * bound
        params[f"{name}.bias"] = (torch.rand(fan_out, generator=generator) * 2 - 1) * bound

    for net in ("actor", "critic"):
        d = obs_dim
        for i, h in enumerate(hidden):
            linear(f"{net}.backbone.layers.{2 * i}", d, h)
            d = h
        if net == "actor":
            linear("actor.distribution.mean_head", d, act_dim)
            params["actor.distribution.std.param"] = torch.full((act_dim,), float(init_std))
        else:
            linear("critic.value_head", d, value_dim)
    return params


def lstm_sequence_ref(
    x: Tensor, done: Tensor, h0: Tensor, c0: Tensor, weights: list[tuple[Tensor, Tensor, Tensor, Tensor]]
) -> tuple[Tensor, Tensor, Tensor]:
    """nn/module/rnn.py:264-299 + nn/utils/recurrent.py:160-272 restated as the mathematically
    identical in-line reset form (pinned by cusrl_test/nn/module/test_rnn.py:145-164):
    the state entering step t+1 is zeroed where done[t].  x [T,N,I]; done [T,N,1]; h0,c0 [L,N,H];
    weights per layer (w_ih [4H,I], w_hh [4H,H], b_ih, b_hh) in torch gate order i,f,g,o."""
    T = x.size(0)
    h = [h0[l] for l in range(len(weights))]
    c = [c0[l] for l in range(len(weights))]
    outs = []
    for t in range(T):
        inp = x[t]
        for l, (w_ih, w_hh, b_ih, b_hh) in enumerate(weights):
            gates = torch.nn.functional.linear(inp, w_ih, b_ih) + torch.nn.functional.linear(h[l], w_hh, b_hh)
            i, f, g, o = gates.chunk(4, dim=-1)
            c[l] = torch.sigmoid(f) * c[l] + torch.sigmoid(i) * torch.tanh(g)
            h[l] = torch.sigmoid(o) * torch.tanh(c[l])
            inp = h[l]
        outs.append(inp)
        keep = done[t].logical_not().to(x.dtype)
        h = [hl * keep for hl in h]
        c = [cl * keep for cl in c]
    return torch.stack(outs), torch.stack(h), torch.stack(c)


# =================================================================================================
# Optimiser side
# =================================================================================================
def clip_grad_norm_ref(grads: list[Tensor], max_norm: float) -> tuple[Tensor, list[Tensor]]:
    """hook/on_policy/gradient_clipping.py:74 -> torch clip_grad_norm_ (L2, eps 1e-6, coef clamped to 1)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, [g * coef for g in grads]


def adam_step_ref(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
                  betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> tuple[Tensor, Tensor, Tensor]:
    """torch.optim.Adam single-tensor update (preset/optimizer.py:9-23 builds torch.optim.Adam)."""
    b1, b2 = betas
    m = m + (g - m) * (1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


# =================================================================================================
# RND (hook/auxiliary/rnd.py)
# =================================================================================================
def small_mlp_ref(ws: list[tuple[Tensor, Tensor]], x: Tensor, activation: str = "ReLU") -> Tensor:
    """nn/module/mlp.py:77-90 with ends_with_activation=False (plain Mlp.Factory used by RND tests)."""
    h = x
    for i, (w, b) in enumerate(ws):
        h = torch.nn.functional.linear(h, w, b)
        if i != len(ws) - 1:
            h = activation_ref(h, activation)
    return h


def rnd_reward_ref(target_ws, predictor_ws, next_state: Tensor, reward_scale: float, activation: str = "ReLU") -> Tensor:
    """hook/auxiliary/rnd.py:68-75: reward_scale * mean_d (target - prediction)^2, keepdim."""
    t, p = small_mlp_ref(target_ws, next_state, activation), small_mlp_ref(predictor_ws, next_state, activation)
    return reward_scale * (t - p).square().mean(dim=-1, keepdim=True)


def rnd_loss_ref(target_ws, predictor_ws, next_state: Tensor, activation: str = "ReLU") -> Tensor:
    """hook/auxiliary/rnd.py:77-81: MSELoss(predictor(x), target(x))."""
    return torch.nn.functional.mse_loss(small_mlp_ref(predictor_ws, next_state, activation),
                                        small_mlp_ref(target_ws, next_state, activation))


# =================================================================================================
# Whole iteration (CPU baseline "port" and end-to-end parity oracle)
# =================================================================================================
@dataclass
class PpoConfig:
    """Anymal-C-rough preset: zoo/isaaclab/locomotion.py:48-59 over preset/ppo.py:79-130 defaults."""

    obs_dim: int = 235
    act_dim: int = 12
    hidden: tuple[int, ...] = (512, 256, 128)
    activation: str = "ELU"
    num_steps: int = 24
    epochs: int = 5
    mini_batches: int = 4
    gamma: float = 0.99
    lamda: float = 0.95
    lamda_value: float | None = None
    clip_ratio: float = 0.2
    value_loss_weight: float = 0.5
    value_loss_clip: float | None = None
    surrogate_loss_weight: float = 1.0
    entropy_loss_weight: float = 0.005
    max_grad_norm: float | None = 1.0
    lr: float = 1e-3
    desired_kl: float | None = 0.015
    normalize_advantage: bool = True


@dataclass
class OracleRollout:
    """Time-major rollout leaves as template/buffer.py:124-151 stores them."""

    leaves: dict[str, Tensor] = field(default_factory=dict)


class OraclePpo:
    """The reference's ActorCritic PPO iteration restated on CPU (template/actor_critic.py:227-320).

    Networks are evaluated with torch autograd on CPU; all other arithmetic uses the *_ref functions
    above.  Used (1) as the end-to-end parity oracle at small sizes and (2) as the timed CPU baseline.
    """

    def __init__(self, cfg: PpoConfig, params: dict[str, Tensor]):
        self.cfg = cfg
        self.params = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.names = list(self.params)
        self.m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.step_count = 0
        self.lr_scale = 1.0
        self._acc_log_err = 0.0
        self._acc_count = 0
        self.metrics: dict[str, float] = {}

    # -- rollout ---------------------------------------------------------------------------------
    @torch.no_grad()
    def act(self, obs: Tensor, noise: Tensor) -> dict[str, Tensor]:
        """actor_critic.py:227-253 + value.py:42-51: explore (mean + std * noise) and critic value."""
        c = self.cfg
        mean, std = actor_forward_ref(self.params, obs, len(c.hidden), c.activation)
        action = mean + std * noise                         # Normal.rsample
        return {
            "observation": obs, "action": action, "action_logp": normal_log_prob_ref(mean, std, action),
            "action_dist.mean": mean, "action_dist.std": std,
            "value": critic_forward_ref(self.params, obs, len(c.hidden), c.activation),
        }

    @torch.no_grad()
    def boot_value(self, next_obs_last: Tensor) -> Tensor:
        c = self.cfg
        return critic_forward_ref(self.params, next_obs_last, len(c.hidden), c.activation)

    # -- update ----------------------------------------------------------------------------------
    def pre_update(self, buf: dict[str, Tensor]) -> None:
        """preset/ppo.py:37-49 order: ValueComputation -> GAE -> AdvantageNormalization."""
        c = self.cfg
        boot = self.boot_value(buf["next_observation"][-1])
        buf["next_value"] = next_value_ref(buf["value"], buf["terminated"], buf["truncated"], boot)
        buf["advantage"], buf["return"] = advantage_and_return_ref(
            buf["reward"], buf["done"], buf["value"], buf["next_value"], c.gamma, c.lamda, c.lamda_value)
        if c.normalize_advantage:
            buf["advantage"] = normalize_advantage_ref(buf["advantage"])

    def train_step(self, batch: dict[str, Tensor]) -> dict[str, float]:
        """actor_critic.py:302-320 for one minibatch (single process: reduce_gradients is a no-op)."""
        c = self.cfg
        L = len(c.hidden)
        curr_value = critic_forward_ref(self.params, batch["observation"], L, c.activation)
        l_v = value_loss_ref(batch["value"], curr_value, batch["return"], c.value_loss_clip) * c.value_loss_weight
        mean, std = actor_forward_ref(self.params, batch["observation"], L, c.activation)
        logp = normal_log_prob_ref(mean, std, batch["action"])
        entropy = normal_entropy_ref(std)
        ratio = (logp - batch["action_logp"]).exp()
        l_s = surrogate_loss_ref(batch["advantage"], ratio, c.clip_ratio) * c.surrogate_loss_weight
        l_e = -entropy.mean() * c.entropy_loss_weight
        loss = sum([l_v, l_s, l_e])                         # actor_critic.py:309
        grads = torch.autograd.grad(loss, [self.params[k] for k in self.names])
        norm = None
        if c.max_grad_norm is not None:
            norm, grads = clip_grad_norm_ref(list(grads), c.max_grad_norm)
        self.step_count += 1
        lr = c.lr * self.lr_scale
        with torch.no_grad():
            for k, g in zip(self.names, grads):
                p, self.m[k], self.v[k] = adam_step_ref(self.params[k], g, self.m[k], self.v[k], self.step_count, lr)
                self.params[k].copy_(p)
        out = {"value_loss": float(l_v.detach()), "surrogate_loss": float(l_s.detach()), "entropy_loss": float(l_e.detach())}
        if norm is not None:
            out["grad_norm/default"] = float(norm)
        return out

    @torch.no_grad()
    def post_update(self, buf: dict[str, Tensor]) -> dict[str, float]:
        """stats.py:29-40 + lr_schedule.py:60-67,229-239 (AdaptiveLRSchedule, threshold 1, factor 0.2)."""
        c = self.cfg
        obs = buf["observation"].flatten(0, 1)
        mean, std = actor_forward_ref(self.params, obs, len(c.hidden), c.activation)
        kl, iwa, sd = policy_stats_ref(
            buf["action_dist.mean"].flatten(0, 1), buf["action_dist.std"].flatten(0, 1), mean, std,
            buf["action"].flatten(0, 1), buf["action_logp"].flatten(0, 1), buf["advantage"].flatten(0, 1))
        out = {"kl_divergence": float(kl), "importance_weighted_advantage": float(iwa), "action_std": float(sd)}
        if c.desired_kl is not None:
            k = max(float(kl), 1e-5)
            self._acc_log_err += math.log(k / c.desired_kl)
            self._acc_count += 1
            if not (1.0 > self._acc_log_err > -1.0):
                avg = self._acc_log_err / self._acc_count
                self.lr_scale *= math.exp(-min(max(avg, -1.0), 1.0) * 0.2)
                self._acc_log_err, self._acc_count = 0.0, 0
            out["lr_scale"] = self.lr_scale
        return out

    def update(self, buf: dict[str, Tensor], perms: list[Tensor]) -> list[dict[str, float]]:
        """actor_critic.py:293-300 with the sampler loop of mini_batch_sampler.py:52-78.
        ``perms``: one permutation of range(T*N) per epoch (the reference draws torch.randperm)."""
        c = self.cfg
        self.pre_update(buf)
        logs = []
        for epoch in range(c.epochs):
            for idx in minibatch_slices_ref(perms[epoch], c.mini_batches):
                batch = {k: gather_ref(v, idx) for k, v in buf.items()}
                logs.append(self.train_step(batch))
        self.metrics = self.post_update(buf)
        return logs
