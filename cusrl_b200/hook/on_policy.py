"""B200 implementations of the on-policy PPO hooks, under the reference's class names.

Every class keeps the reference hook's name (-> snake_case identity, template/hook.py), constructor
arguments, mutable attributes, metric keys and error behaviour, and reads / writes the same buffer and
batch keys, so by-name addressing (``register_hook(before=...)``), schedules and checkpoints keep working.
Reference files: cusrl/hook/on_policy/{value,gae,advantage,common,ppo,gradient_clipping,stats,lr_schedule}.py
and cusrl/hook/control/initialization.py.

Objective fusion.  The reference evaluates ValueLoss -> OnPolicyPreparation -> PpoSurrogateLoss ->
EntropyLoss as ~30 elementwise kernels plus their autograd twins.  Here ``ValueLoss`` and
``OnPolicyPreparation`` only run the networks and publish ``curr_value`` / ``curr_action_dist``;
``PpoSurrogateLoss.objective`` launches ONE kernel (K4) that produces all three weighted losses, the
per-sample tensors the reference publishes (``curr_action_logp``, ``curr_entropy``, ``action_logp_ratio``,
``action_prob_ratio``), the metric sums and the gradients w.r.t. mean / std / value; ``EntropyLoss`` returns its
slice of that result.  Dict insertion order (value, surrogate, entropy) and therefore the float order of
``sum(objectives.values())`` (actor_critic.py:309) is preserved.  The hooks must appear in the preset order
(preset/ppo.py:50-53); anything else is refused with a clear error instead of silently degrading.
"""

from __future__ import annotations

import copy
import math
from typing import Any

import torch
from torch import Tensor, nn

from .. import distributed, ops
from ..template.buffer import Buffer, Sampler
from ..template.hook import Hook

__all__ = [
    "AdaptiveLRSchedule",
    "AdvantageNormalization",
    "AdvantageReduction",
    "EntropyLoss",
    "GeneralizedAdvantageEstimation",
    "GradientClipping",
    "ModuleInitialization",
    "OnPolicyPreparation",
    "OnPolicyStatistics",
    "PpoSurrogateLoss",
    "ValueComputation",
    "ValueLoss",
]


def _first(mapping, *keys):
    for key in keys:
        if (value := mapping.get(key)) is not None:
            return value
    raise KeyError(f"none of {keys} found")


def _active_neighbor(hook: Hook, offset: int):
    """The active hook `offset` places after (+) / before (-) `hook` in the agent's hook list, or None."""
    hooks = list(hook.agent.hook.active_hooks())
    for i, h in enumerate(hooks):
        if h is hook:
            j = i + offset
            return hooks[j] if 0 <= j < len(hooks) else None
    return None


def _std_vector(std: Tensor, param: Tensor) -> Tensor:
    """The state-independent std of the current action distribution as the ``[A]`` vector the fused kernels take, WITH its
    autograd history.  For the plain ``NormalDist`` the distribution's std is an expanded view of the parameter vector and
    the parameter itself is handed over (no extra autograd nodes on the hot path).  A wrapper that post-processes the std
    (``SymmetricActor``: ``(std + |mirror(std)|) / 2``, symmetry.py:455-458) still yields identical rows, so row 0 carries
    the value, and -- every row being the same function of the parameters -- the gradient of a sum over samples w.r.t. the
    shared vector flows through row 0 exactly."""
    if std.data_ptr() == param.data_ptr() and all(s == 0 for s in std.stride()[:-1]):
        return param
    return std.reshape(-1, std.shape[-1])[0]


def _leaf(buffer: Buffer, name: str, like: Tensor) -> Tensor:
    """Persistent ``[T, N, Dv]`` leaf `name` in the buffer (allocated on first use, then overwritten in place)."""
    if name not in buffer:
        buffer[name] = torch.zeros_like(like)
    return buffer[name]


# =================================================================================================
class ModuleInitialization(Hook):
    """Orthogonal initialisation at ``init`` of the linear AND recurrent layers (and attention / convolution layers of user
    modules) of actor and critic (reference hook/control/initialization.py:11-125: same arguments, same per-module rules,
    same traversal order -> same consumption of the random stream).  The Anymal preset sets ``orthogonal_init=False``
    (zoo/isaaclab/locomotion.py:56) -> no-op there; the recurrent preset keeps the default, so its LSTM weights are
    orthogonal with gain sqrt(2) and its biases zero."""

    def __init__(self, scale: float = math.sqrt(2), scale_dist: float = math.sqrt(2) * 0.1, zero_bias: bool = True,
                 conv_a: float = 0.0, conv_mode: str = "fan_in", conv_nonlinearity: str = "leaky_relu",
                 init_actor: bool = True, init_critic: bool = True):
        super().__init__()
        self.scale, self.scale_dist, self.zero_bias = scale, scale_dist, zero_bias
        self.conv_a, self.conv_mode, self.conv_nonlinearity = conv_a, conv_mode, conv_nonlinearity
        self.init_actor, self.init_critic = init_actor, init_critic

    def init(self) -> None:
        if self.init_actor:
            for module in self.agent.actor.modules():
                self._init_module(module, self.scale, self.zero_bias)
            if self.scale_dist != self.scale:
                self._init_linear(self.agent.actor.distribution.mean_head, self.scale_dist, self.zero_bias)
        if self.init_critic:
            for module in self.agent.critic.modules():
                self._init_module(module, self.scale, self.zero_bias)

    def _init_module(self, module: nn.Module, scale: float, zero_bias: bool) -> None:
        if isinstance(module, nn.Linear):
            self._init_linear(module, scale, zero_bias)
        elif isinstance(module, (nn.RNN, nn.LSTM, nn.GRU)):
            self._init_rnn(module, scale, zero_bias)
        elif isinstance(module, nn.MultiheadAttention):
            self._init_mha(module, scale, zero_bias)
        elif isinstance(module, nn.Conv2d):
            self._init_conv2d(module, zero_bias)

    @staticmethod
    def _zero(*tensors) -> None:
        for tensor in tensors:
            if tensor is not None:
                nn.init.zeros_(tensor)

    def _init_linear(self, module: nn.Linear, scale: float, zero_bias: bool) -> None:
        nn.init.orthogonal_(module.weight, gain=scale)
        if zero_bias:
            self._zero(module.bias)

    def _init_rnn(self, module: nn.RNNBase, scale: float, zero_bias: bool) -> None:
        for layer in range(module.num_layers):   # recurrent matrix first, like the reference (random-stream order)
            for kind in ("hh", "ih"):
                nn.init.orthogonal_(getattr(module, f"weight_{kind}_l{layer}"), gain=scale)
            if zero_bias:
                self._zero(getattr(module, f"bias_hh_l{layer}", None), getattr(module, f"bias_ih_l{layer}", None))

    def _init_mha(self, module: nn.MultiheadAttention, scale: float, zero_bias: bool) -> None:
        packed = module.in_proj_weight is not None
        for weight in ((module.in_proj_weight,) if packed else (module.q_proj_weight, module.k_proj_weight, module.v_proj_weight)):
            nn.init.orthogonal_(weight, gain=scale)
        if zero_bias:
            self._zero(module.in_proj_bias, module.bias_k, module.bias_v)

    def _init_conv2d(self, module: nn.Conv2d, zero_bias: bool) -> None:
        nn.init.kaiming_normal_(module.weight, a=self.conv_a, mode=self.conv_mode, nonlinearity=self.conv_nonlinearity)
        if zero_bias:
            self._zero(module.bias)


# =================================================================================================
class ValueComputation(Hook):
    """Critic values at act time and ``next_value`` construction before the update (K3).

    Reference: hook/on_policy/value.py:14-82.  ``pre_update`` is one critic forward on ``next_state[-1]`` plus
    one kernel; the reference's masked scatters and its ``truncated.any()`` host sync are gone."""

    def __init__(self, *, termination_value: float = 0.0, bootstrap_truncated_states: bool = True):
        super().__init__()
        self.termination_value = termination_value
        self.bootstrap_truncated_states = bootstrap_truncated_states
        self._critic_memory = None

    def init(self) -> None:
        if self.agent.environment_spec.final_state_is_missing:
            self.bootstrap_truncated_states = False

    def post_act(self, transition) -> None:
        state = _first(transition, "state", "observation")
        value, next_memory = self.agent.critic(state, memory=self._critic_memory)
        transition["value"] = value
        transition["critic_memory"] = self._critic_memory
        transition["next_critic_memory"] = next_memory
        self._critic_memory = next_memory

    def post_step(self, transition) -> None:
        self.agent.critic.reset_memory(self._critic_memory, transition["done"])

    @torch.no_grad()
    def pre_update(self, buffer: Buffer) -> None:
        critic = self.agent.critic
        value = buffer["value"]
        next_value = _leaf(buffer, "next_value", value)
        next_state = _first(buffer, "next_state", "next_observation")
        boot = critic.evaluate(next_state[-1], memory=self._critic_memory)
        buffer.private.pop("pending_next_value", None)
        nxt = _active_neighbor(self, +1)
        if (not self.bootstrap_truncated_states and isinstance(nxt, GeneralizedAdvantageEstimation) and not nxt.recompute
                and value.dim() == 3 and ops.gae_chain_supported(value.shape[0], value.shape[2])):
            # The hook that runs NEXT is the B200 GAE hook: it forms next_value on the fly inside its scan and publishes it
            # to buffer["next_value"] from the same launch (K3 + K1 + the K2 statistics fused, ops.gae_chain), so nothing
            # is launched here.  No other hook runs in between, so nobody can observe the leaf before it is written.
            buffer.private["pending_next_value"] = (boot.contiguous(), float(self.termination_value), next_value)
            return
        trunc_value = None
        if self.bootstrap_truncated_states:
            # value.py:74-80 evaluates the critic on next_state[truncated] (data-dependent size -> host sync);
            # evaluating every next state keeps the stream free of syncs and selects the same entries in-kernel
            next_memory = buffer.get("next_critic_memory")
            if next_memory is not None:  # recurrent critic: every (t, n) is an independent single step with its own memory
                flat_mem = {k: v.flatten(0, 1) for k, v in next_memory.items()}
                trunc_value = critic.evaluate(next_state.flatten(0, 1), memory=flat_mem).reshape(*next_state.shape[:2], -1)
            else:
                trunc_value = critic.evaluate(next_state)
            trunc_value = trunc_value.contiguous()
        ops.next_value(value, buffer["terminated"], buffer["truncated"], boot.contiguous(), self.termination_value,
                       trunc_value=trunc_value, out=next_value)


class GeneralizedAdvantageEstimation(Hook):
    """GAE advantages and returns in one launch (K1).  Reference: hook/on_policy/gae.py:23-110."""

    def __init__(self, gamma: float = 0.99, lamda: float = 0.95, lamda_value: float | None = None,
                 recompute: bool = False):
        if gamma < 0 or gamma >= 1:
            raise ValueError(f"'gamma' must be in [0, 1); got {gamma}")
        if lamda < 0 or lamda > 1:
            raise ValueError(f"'lamda' must be in [0, 1]; got {lamda}")
        if lamda_value is not None and (lamda_value < 0 or lamda_value > 1):
            raise ValueError(f"'lamda_value' must be in [0, 1]; got {lamda_value}")
        super().__init__(training_only=True)
        self.recompute = recompute
        self.gamma, self.lamda, self.lamda_value = gamma, lamda, lamda_value
        for name in ("gamma", "lamda", "lamda_value"):
            self.register_mutable(name)

    def pre_update(self, buffer) -> None:
        if not self.recompute:
            self._compute_advantage_and_return(buffer)

    def objective(self, metadata, batch):
        if self.recompute:
            self._compute_advantage_and_return(batch)

    @torch.no_grad()
    def _compute_advantage_and_return(self, data) -> None:
        value = data["value"]
        if isinstance(data, Buffer):
            advantage, ret = _leaf(data, "advantage", value), _leaf(data, "return", value)
            data.private.pop("advantage_stats", None)
            pending = data.private.pop("pending_next_value", None)
            if pending is not None:
                boot, termination_value, next_value = pending
                mean_var = ops.gae_chain(data["reward"], data["terminated"], data["truncated"], value, boot, self.gamma,
                                         self.lamda, self.lamda_value, termination_value, next_value, advantage, ret)
                data.private["advantage_stats"] = (mean_var, advantage.data_ptr())
                return
        else:
            advantage, ret = torch.empty_like(value), torch.empty_like(value)
            data["advantage"], data["return"] = advantage, ret
        ops.gae(data["reward"], data["done"], value, data["next_value"], self.gamma, self.lamda, self.lamda_value,
                advantage=advantage, ret=ret)


class AdvantageReduction(Hook):
    """Reduces a vector advantage (one channel per reward term) to the scalar the surrogate needs: optional per-channel
    weights, then sum or mean over the last dim, on the minibatch.  Reference: hook/on_policy/advantage.py:13-71.  The
    Dv > 1 GAE scan that produces the vector advantage is K1 (bit-exact for every Dv); this reduction is a [B, Dv] -> [B, 1]
    weighted row sum evaluated by the output-head kernel (a 1 x Dv "weight matrix"), so it needs no kernel of its own."""

    def __init__(self, reduction: str = "sum", weight=None):
        if reduction not in ("sum", "mean"):
            raise ValueError(f"Unsupported reduction '{reduction}'")
        super().__init__(training_only=True)
        self.reduction = reduction
        self.weight = None if weight is None else tuple(weight)
        self.register_mutable("weight")
        self._weight_tensor: Tensor | None = None

    def init(self) -> None:
        self._weight_tensor = None if self.weight is None else self.agent.to_tensor(self.weight).float()

    def objective(self, metadata, batch):
        advantage = batch["advantage"]
        if self._weight_tensor is not None:
            advantage = advantage * self._weight_tensor
        if self.reduction == "sum":
            advantage = advantage.sum(-1, keepdim=True)
        elif self.reduction == "mean":
            advantage = advantage.mean(-1, keepdim=True)
        else:
            raise ValueError(f"Unsupported reduction '{self.reduction}'")
        batch["advantage"] = advantage
        return None

    def update_attribute(self, name: str, value) -> None:
        super().update_attribute(name, value)
        if name == "weight":
            if value is None:
                self.weight = self._weight_tensor = None
            else:
                self.weight = tuple(value)
                self._weight_tensor = self.agent.to_tensor(self.weight).float()


class AdvantageNormalization(Hook):
    """Standardise advantages (K2) with the reference's cross-rank statistic merge.
    Reference: hook/on_policy/advantage.py:74-115."""

    def __init__(self, mini_batch_wise: bool = False, synchronize: bool = True):
        super().__init__(training_only=True)
        self.mini_batch_wise, self.synchronize = mini_batch_wise, synchronize

    def pre_update(self, buffer) -> None:
        if not self.mini_batch_wise:
            advantage = buffer["advantage"]
            stats = buffer.private.pop("advantage_stats", None)
            if (stats is not None and stats[1] == advantage.data_ptr()
                    and isinstance(_active_neighbor(self, -1), GeneralizedAdvantageEstimation)):
                # the GAE hook that ran immediately before already reduced the statistics in its own launch
                self.normalize_(advantage, mean_var=stats[0])
            else:
                self.normalize_(advantage)

    def objective(self, metadata, batch):
        if self.mini_batch_wise:
            self.normalize_(batch["advantage"])

    @torch.no_grad()
    def normalize_(self, advantage: Tensor, mean_var: Tensor | None = None) -> None:
        if mean_var is None:
            mean_var = ops.advantage_stats(advantage)
        if self.synchronize:
            distributed.reduce_mean_var_(mean_var)
        ops.advantage_normalize_(advantage, mean_var, 1e-8)


# =================================================================================================
class _FusedObjective(torch.autograd.Function):
    """Autograd node around the K4 result: forward hands out the three weighted losses, backward scales the
    precomputed unit gradients by the upstream scalars (device-side, no host read)."""

    @staticmethod
    def forward(ctx, mean, std_param, curr_value, out):
        ctx.out = out
        ctx.has_value = curr_value is not None
        losses = out["losses"]
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_value, g_surr, g_ent):
        out = ctx.out
        d_mean = d_std = d_value = None
        if g_surr is not None:
            d_mean = ops.scale_(out["d_mean"], g_surr.reshape(1).contiguous())
            d_std = out["d_std_surr"] * g_surr
        if g_ent is not None:
            term = out["d_std_ent"] * g_ent
            d_std = term if d_std is None else d_std + term
        if ctx.has_value and g_value is not None:
            d_value = ops.scale_(out["d_value"], g_value.reshape(1).contiguous())
        return d_mean, d_std, d_value, None


class ValueLoss(Hook):
    """Critic forward on the minibatch; the loss arithmetic itself is part of K4 (see module docstring).
    Reference: hook/on_policy/value.py:92-144."""

    def __init__(self, weight: float = 0.5, loss_clip: float | None = None):
        if weight <= 0:
            raise ValueError("'weight' must be positive")
        if loss_clip is not None and loss_clip <= 0:
            raise ValueError("'loss_clip' must be positive or None")
        super().__init__()
        self.weight, self.loss_clip = weight, loss_clip
        self.register_mutable("weight")
        self.register_mutable("loss_clip")

    def objective(self, metadata, batch):
        state = _first(batch, "state", "observation")
        curr_value = self.agent.critic.evaluate(state, memory=batch.get("critic_memory"), done=batch["done"])
        batch["curr_value"] = curr_value
        batch["_b200_value_loss"] = (self.weight, self.loss_clip)
        return None  # "value_loss" is emitted by PpoSurrogateLoss from the fused kernel, ahead of "surrogate_loss"

    def post_objective(self, metadata, batch) -> None:
        fused = batch.get("_b200_fused")
        if fused is None:
            raise RuntimeError("ValueLoss needs a PpoSurrogateLoss hook after it (fused objective, see cusrl_b200.hook)")
        curr_value = batch["curr_value"]
        self.agent.metrics.record_mean("value", fused["metrics"][2], curr_value.numel() // curr_value.shape[-1])
        if (dv := curr_value.size(-1)) != 1:
            with torch.no_grad():
                self.agent.record(**{f"value.{i}": curr_value[..., i] for i in range(dv)})


class OnPolicyPreparation(Hook):
    """Actor forward on the minibatch; publishes ``curr_action_dist`` (the per-sample log-prob / entropy / ratio
    tensors are published by the fused kernel).  Reference: hook/on_policy/common.py:13-49."""

    def __init__(self, calculate_kl_divergence: bool = False):
        super().__init__(training_only=True)
        self.calculate_kl_divergence = calculate_kl_divergence

    def objective(self, metadata, batch):
        actor = self.agent.actor
        action_dist, _ = actor(batch["observation"], memory=batch.get("actor_memory"), done=batch["done"])
        batch["curr_action_dist"] = action_dist
        if self.calculate_kl_divergence:
            with torch.no_grad():
                batch["kl_divergence"] = actor.compute_kl_div(batch["action_dist"], action_dist)
        return None

    def post_objective(self, metadata, batch) -> None:
        fused = batch.get("_b200_fused")
        if fused is None:
            raise RuntimeError("OnPolicyPreparation needs a PpoSurrogateLoss hook after it (fused objective)")
        n = batch["action_logp_ratio"].numel()
        self.agent.metrics.record_mean("ratio", fused["metrics"][0], n)
        self.agent.metrics.record_mean("entropy", fused["metrics"][1], n)


class PpoSurrogateLoss(Hook):
    """Clipped surrogate; launches the fused objective kernel (K4).  Reference: hook/on_policy/ppo.py:21-55."""

    def __init__(self, clip_ratio: float = 0.2, weight: float = 1.0):
        if clip_ratio <= 0:
            raise ValueError("'clip_ratio' must be positive")
        if weight < 0:
            raise ValueError("'weight' must be non-negative")
        super().__init__(training_only=True)
        self.clip_ratio, self.weight = clip_ratio, weight
        self.register_mutable("clip_ratio")
        self.register_mutable("weight")

    def _entropy_weight(self) -> float:
        for hook in self.agent.hook.active_hooks():
            if isinstance(hook, EntropyLoss):
                return hook.weight
        return 0.0

    def objective(self, metadata, batch):
        advantage = batch["advantage"]
        if advantage.size(-1) != 1:
            raise ValueError(f"Expected advantage to have shape [..., 1], got {advantage.shape}")
        dist = batch.get("curr_action_dist")
        if dist is None:
            raise RuntimeError("PpoSurrogateLoss needs an OnPolicyPreparation hook before it")
        std_param = _std_vector(dist["std"], self.agent.actor.distribution.std.param)
        mean = dist["mean"]
        lead = mean.shape[:-1]
        value_cfg = batch.get("_b200_value_loss")
        curr_value = batch.get("curr_value") if value_cfg is not None else None
        w_value, value_clip = value_cfg if value_cfg is not None else (0.0, None)
        flat = lambda t: None if t is None else t.reshape(-1, t.shape[-1]).contiguous()  # noqa: E731
        out = ops.ppo_loss(
            flat(mean.detach()), std_param.detach(), flat(batch["action"]), flat(batch["action_logp"]), flat(advantage),
            flat(batch["return"]) if curr_value is not None else None,
            flat(batch["value"]) if curr_value is not None else None,
            flat(curr_value.detach()) if curr_value is not None else None,
            self.clip_ratio, self.weight, self._entropy_weight(), w_value, value_clip)
        l_value, l_surr, l_ent = _FusedObjective.apply(mean.reshape(-1, mean.shape[-1]), std_param,
                                                       flat(curr_value) if curr_value is not None else None, out)
        batch["_b200_fused"] = out
        batch["_b200_entropy_loss"] = l_ent
        batch["curr_action_logp"] = out["logp"].reshape(*lead, 1)
        batch["curr_entropy"] = out["entropy"].reshape(*lead, 1)
        batch["action_logp_ratio"] = out["logp_ratio"].reshape(*lead, 1)
        batch["action_prob_ratio"] = out["prob_ratio"].reshape(*lead, 1)
        objectives = {}
        if curr_value is not None:
            objectives["value_loss"] = l_value
        objectives["surrogate_loss"] = l_surr
        return objectives


class EntropyLoss(Hook):
    """Entropy bonus; returns its slice of the fused result.  Reference: hook/on_policy/ppo.py:58-84."""

    def __init__(self, weight: float = 0.01):
        if weight < 0:
            raise ValueError("'weight' must be non-negative")
        super().__init__(training_only=True)
        self.weight = weight
        self.register_mutable("weight")

    def objective(self, metadata, batch):
        loss = batch.get("_b200_entropy_loss")
        if loss is None:
            raise RuntimeError("EntropyLoss needs a PpoSurrogateLoss hook before it (fused objective)")
        return {"entropy_loss": loss}


# =================================================================================================
class GradientClipping(Hook):
    """Per-prefix gradient-norm clipping on the flat arena; for the default single group the scaling is folded
    into the Adam kernel (K9).  Reference: hook/on_policy/gradient_clipping.py:8-83."""

    def __init__(self, max_grad_norm: float | None = 1.0, groups: dict[str, float | None] | None = None, **kwargs):
        super().__init__(training_only=True)
        if max_grad_norm is not None and max_grad_norm < 0:
            raise ValueError("'max_grad_norm' must be non-negative")
        self.max_grad_norm = max_grad_norm
        groups = (groups or {}) | kwargs
        for prefix, limit in groups.items():
            if not prefix:
                raise ValueError("Empty prefixes are not allowed; use 'max_grad_norm' for the default group")
            if limit is not None and limit < 0:
                raise ValueError(f"'max_grad_norm' for prefix '{prefix}' must be non-negative")
        self.groups = dict(sorted(groups.items(), key=lambda x: len(x[0]), reverse=True))
        self._scratch: dict[str, tuple[Tensor, Tensor, Tensor]] = {}

    def _match_prefix(self, name: str) -> str:
        for prefix in self.groups:
            if name == prefix or name.startswith(f"{prefix}."):
                return prefix
        return ""

    def pre_optim(self, optimizer) -> None:
        arena = optimizer.arena
        if not self.groups:
            if self.max_grad_norm is not None:
                norm = optimizer.compute_grad_norm(self.max_grad_norm)
                self.agent.metrics.record_mean("grad_norm/default", norm, 1)
            return
        ranges: dict[str, list[tuple[int, int]]] = {"": [], **{p: [] for p in self.groups}}
        for name, param, off in zip(arena.names, arena.params, arena.offsets):
            ranges[self._match_prefix(name)].append((off, param.numel()))
        for prefix, segs in ranges.items():
            limit = self.groups.get(prefix, self.max_grad_norm)
            if not segs or limit is None:
                continue
            if prefix not in self._scratch:
                dev = arena.flat.device
                self._scratch[prefix] = (torch.zeros(1, dtype=torch.float64, device=dev),
                                         torch.zeros(1, device=dev), torch.ones(1, device=dev))
            sumsq, norm, coef = self._scratch[prefix]
            sumsq.zero_()
            for off, n in segs:
                ops.grad_sumsq_(arena.flat_grad[off : off + n], sumsq)
            ops.clip_coef(sumsq, limit, norm, coef)
            for off, n in segs:
                ops.scale_(arena.flat_grad[off : off + n], coef)
            self.agent.metrics.record_mean(f"grad_norm/{prefix or 'default'}", norm, 1)


class OnPolicyStatistics(Hook):
    """KL(old || new), importance-weighted advantage and action std after the update.

    Reference: hook/on_policy/stats.py:8-40 gathers a shuffled copy of every leaf and re-runs the actor on it.
    The recorded means do not depend on the order, so the gather is skipped: one actor forward over the buffer
    in place plus one statistics kernel.  ``preserve_rng_stream`` still draws the ``randperm`` the reference's
    sampler would have drawn, keeping the generator state (and so next iteration's permutations) identical."""

    def __init__(self, sampler: Sampler | None = None, preserve_rng_stream: bool = True):
        super().__init__(training_only=True)
        self.sampler = sampler if sampler is not None else Sampler()
        self.preserve_rng_stream = preserve_rng_stream

    @torch.no_grad()
    def post_update(self) -> None:
        agent = self.agent
        buffer, actor = agent.buffer, agent.actor
        if self.preserve_rng_stream and hasattr(self.sampler, "indices"):
            for _ in self.sampler.indices(buffer):
                pass
        obs = buffer["observation"]
        action_dist, _ = actor(obs, memory=buffer.get("actor_memory"), done=buffer["done"])
        old = buffer["action_dist"]
        E = obs.shape[0] * obs.shape[1]
        # leaves whose rows the buffer pads to 16-byte multiples (action dims 17-19, 21-23, ...) are narrow views of the
        # padded storage: the statistics kernel takes dense rows, so those (and only those) are compacted first
        out = ops.policy_stats(old["mean"].contiguous(), old["std"].contiguous(), action_dist["mean"].contiguous(),
                               _std_vector(action_dist["std"], actor.distribution.std.param).detach().contiguous(),
                               buffer["action"].contiguous(),
                               buffer["action_logp"], buffer["advantage"])
        agent.metrics.record_mean("kl_divergence", out[0], E)
        agent.metrics.record_mean("importance_weighted_advantage", out[1], E)
        agent.metrics.record_mean("action_std", out[2], E * old["std"].shape[-1])


class AdaptiveLRSchedule(Hook):
    """KL-adaptive learning-rate schedule (host logic on one scalar).

    Reference: hook/on_policy/lr_schedule.py:19-112,173-239 (KLDivergenceBasedLRSchedule + AdaptiveLRSchedule):
    accumulate log(kl / target); when |acc| >= threshold rescale the LR of every param group that contains an
    ``actor.*`` parameter by exp(-clip(mean_err, +-1) * scale_factor)."""

    def __init__(self, desired_kl_divergence: float = 0.01, *, max_kl_divergence: float | None = None,
                 threshold: float = 1.0, scale_factor: float = 0.2, scale_all_params: bool = False,
                 warmup_iterations: int = 0, initial_scale: float = 0.0):
        if desired_kl_divergence <= 0:
            raise ValueError("'desired_kl_divergence' must be positive")
        if warmup_iterations < 0:
            raise ValueError("'warmup_iterations' must be non-negative")
        if not 0 <= initial_scale <= 1:
            raise ValueError("'initial_scale' must be within [0, 1]")
        if max_kl_divergence is not None and max_kl_divergence <= 0:
            raise ValueError("'max_kl_divergence' must be positive")
        if threshold <= 0:
            raise ValueError("'threshold' must be positive")
        if scale_factor <= 0:
            raise ValueError("'scale_factor' must be positive")
        super().__init__(training_only=True)
        self.scale_all_params, self.warmup_iterations, self.initial_scale = scale_all_params, warmup_iterations, initial_scale
        self.threshold, self.scale_factor = threshold, scale_factor
        self.desired_kl_divergence, self.max_kl_divergence = desired_kl_divergence, max_kl_divergence
        self.register_mutable("desired_kl_divergence")
        self.register_mutable("max_kl_divergence")
        self._lr_scale = 1.0
        self._base_lrs: list[float] = []
        self._checkpoint: dict | None = None
        self._accumulated_log_error = 0.0
        self._count = 0

    def post_init(self) -> None:
        self._base_lrs = [g["lr"] for g in self.agent.optimizer.param_groups]

    def pre_update(self, buffer) -> None:
        if self.max_kl_divergence is not None:
            self._checkpoint = copy.deepcopy(self.agent.state_dict())

    def post_update(self) -> None:
        kl = self.agent.metrics["kl_divergence"].mean.clone()
        distributed.reduce_mean_(kl)
        kl_value = kl.item()  # the schedule is a host decision (lr_schedule.py:65)
        if self.agent.iteration >= self.warmup_iterations:
            self._scale_lr(self._compute_scale(kl_value))
            self.agent.record(lr_scale=self._lr_scale)
        if self.max_kl_divergence is not None:
            checkpoint, self._checkpoint = self._checkpoint, None
            if kl_value > self.max_kl_divergence:
                lr_scale = self._lr_scale
                self.agent.load_state_dict(checkpoint)
                self._lr_scale = lr_scale
                self._apply_lr_scale()
                self.agent.record(update_rejected=1.0)
            else:
                self.agent.record(update_rejected=0.0)

    def apply_schedule(self, iteration: int) -> None:
        if self.warmup_iterations <= 0 or iteration > self.warmup_iterations:
            return
        progress = min(iteration, self.warmup_iterations) / self.warmup_iterations
        self._lr_scale = self.initial_scale + (1.0 - self.initial_scale) * progress
        self._apply_lr_scale()
        self.agent.record(lr_scale=self._lr_scale)

    def _compute_scale(self, kl: float) -> float | None:
        kl = max(kl, 1e-5)
        self._accumulated_log_error += math.log(kl / self.desired_kl_divergence)
        self._count += 1
        if self.threshold > self._accumulated_log_error > -self.threshold:
            return None
        average = self._accumulated_log_error / self._count
        self._accumulated_log_error, self._count = 0.0, 0
        return math.exp(-min(max(average, -1.0), 1.0) * self.scale_factor)

    def _scale_lr(self, scale: float | None) -> None:
        if scale is None or scale == 1.0:
            return
        self._lr_scale *= scale
        self._apply_lr_scale()

    def _apply_lr_scale(self) -> None:
        for base_lr, group in zip(self._base_lrs, self.agent.optimizer.param_groups):
            if self.scale_all_params or any(n.startswith("actor.") for n in group["param_names"]):
                group["lr"] = base_lr * self._lr_scale

    def state_dict(self) -> dict[str, Any]:
        return {"lr_scale": self._lr_scale}

    def load_state_dict(self, state_dict) -> None:
        self._lr_scale = state_dict["lr_scale"]
