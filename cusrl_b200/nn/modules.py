"""Actor / critic network modules for the PPO hot path.

Mirrors the reference's ``Mlp`` (cusrl/nn/module/mlp.py:31-93), ``Actor`` (actor.py:24-268), ``Value``
(critic.py:27-102) and ``NormalDist`` (distribution.py:195-273): same factories, same constructor
arguments, same parameter names (``backbone.layers.{0,2,4}.*``, ``distribution.mean_head.*``,
``distribution.std.param``, ``value_head.*``) so reference checkpoints map 1:1.  The arithmetic runs on
the kernels behind :mod:`cusrl_b200.nn.functional`.
"""

from __future__ import annotations

import math
from collections.abc import Iterable, Sequence
from dataclasses import dataclass
from typing import Any

import torch
from torch import Tensor, nn

from . import functional as F

__all__ = ["Module", "Mlp", "MlpFactory", "NormalDist", "NormalDistFactory", "Actor", "ActorFactory", "Value", "ValueFactory"]

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


def standard_normal_like(mean: Tensor) -> Tensor:
    """The exploration noise: one Philox draw from torch's global generator of the mean's device, the same draw
    ``torch.distributions.Normal.rsample`` makes (reference distribution.py:203), so the generator advances identically.
    Looked up through the module at call time: the parity tests substitute a seeded stream here."""
    return torch.empty_like(mean).normal_()


def _activation_name(fn: str | type[nn.Module]) -> str:
    name = fn if isinstance(fn, str) else fn.__name__
    if name not in F.ACTIVATIONS:
        raise ValueError(f"cusrl_b200 supports activations {sorted(F.ACTIVATIONS)}; got '{name}'")
    return name


class Module(nn.Module):
    """Base with the attributes the reference's hooks rely on (nn/module/module.py:35-163)."""

    def __init__(self, input_dim: int, output_dim: int, is_recurrent: bool = False):
        super().__init__()
        if input_dim <= 0 or output_dim <= 0:
            raise ValueError("'input_dim' and 'output_dim' must be positive integers")
        self.input_dim, self.output_dim, self.is_recurrent = input_dim, output_dim, is_recurrent
        self.intermediate_repr: dict[str, Any] = {}

    @property
    def device(self) -> torch.device:
        """Device of the first parameter, else of the first buffer, else the CPU (module.py:88-97)."""
        tensor = next(self.parameters(), None)
        if tensor is None:
            tensor = next(self.buffers(), None)
        return torch.device("cpu") if tensor is None else tensor.device

    def clear_intermediate_repr(self) -> None:
        self.intermediate_repr.clear()

    def reset_memory(self, memory, done=None) -> None:
        if memory is None:
            return
        sel = slice(None) if done is None else done.squeeze(-1)
        for leaf in (memory.values() if isinstance(memory, dict) else (memory if isinstance(memory, (tuple, list)) else [memory])):
            leaf[sel] = 0

    def step_memory(self, input, memory=None, **kwargs):
        return None


@dataclass(slots=True)
class MlpFactory:
    hidden_dims: Sequence[int]
    activation_fn: str | type[nn.Module] = "ReLU"
    ends_with_activation: bool = False
    dropout: float = 0.0

    def __call__(self, input_dim: int | None = None, output_dim: int | None = None) -> "Mlp":
        assert input_dim is not None
        return Mlp(input_dim, self.hidden_dims, output_dim, self.activation_fn, self.ends_with_activation, self.dropout)


class Mlp(Module):
    """Linear + activation stack; ``layers`` keeps the reference's Sequential indexing for state_dict parity."""

    Factory = MlpFactory

    def __init__(self, input_dim: int, hidden_dims: Iterable[int], output_dim: int | None = None,
                 activation_fn: str | type[nn.Module] = "ReLU", ends_with_activation: bool = False, dropout: float = 0.0):
        dims = list(hidden_dims)
        if output_dim is not None:
            dims.append(output_dim)
        if not dims:
            raise ValueError("Mlp needs at least one layer")
        if dropout != 0.0:
            raise ValueError("cusrl_b200.Mlp does not implement dropout (unused by the PPO presets)")
        super().__init__(input_dim, dims[-1])
        self.activation = _activation_name(activation_fn)
        self.ends_with_activation = bool(ends_with_activation)
        act_cls = getattr(nn, self.activation)
        layers = nn.Sequential()
        d = input_dim
        for i, h in enumerate(dims):
            layers.append(nn.Linear(d, h))
            if i != len(dims) - 1 or ends_with_activation:
                layers.append(act_cls())  # parameter-free placeholder: keeps indices 0,2,4,... for the Linear layers
            d = h
        self.layers = layers

    def linears(self) -> list[nn.Linear]:
        return [m for m in self.layers if isinstance(m, nn.Linear)]

    def forward(self, input: Tensor, **kwargs) -> Tensor:
        lins = self.linears()
        return F.mlp_forward(input, [m.weight for m in lins], [m.bias for m in lins], self.activation,
                             self.ends_with_activation)


class StddevVector(nn.Module):
    """State-independent standard deviation, identity bijector (reference distribution.py:232-249)."""

    def __init__(self, output_dim: int, init_std: float | None = None):
        super().__init__()
        if init_std is not None and init_std <= 0:
            raise ValueError("'init_std' must be positive")
        self.param = nn.Parameter(torch.ones(output_dim) * (1.0 if init_std is None else init_std))

    def forward(self, input: Tensor) -> Tensor:
        # the reference materialises param.repeat(B, 1); an expanded view carries the same values at zero cost
        return self.param.expand(*input.shape[:-1], -1)


class _DeterministicWrapper(nn.Module):
    """``distribution.deterministic()``: latent -> most likely action (distribution.py:181-187)."""

    def __init__(self, distribution: nn.Module):
        super().__init__()
        self.dist = distribution

    def forward(self, backbone_feat: Tensor, **kwargs) -> Tensor:
        return self.dist.determine(backbone_feat, **kwargs)


@dataclass(slots=True)
class NormalDistFactory:
    init_std: float | None = None
    bijector: str | None = None

    def __call__(self, input_dim: int | None = None, output_dim: int | None = None) -> "NormalDist":
        assert input_dim is not None and output_dim is not None
        if self.bijector is not None:
            raise ValueError("cusrl_b200.NormalDist supports only the identity bijector (the PPO preset default)")
        return NormalDist(input_dim, output_dim, init_std=self.init_std)


class NormalDist(Module):
    """Diagonal normal with a linear mean head and a parameter-vector std."""

    Factory = NormalDistFactory

    def __init__(self, input_dim: int, output_dim: int, init_std: float | None = None):
        super().__init__(input_dim, output_dim)
        self.mean_head = nn.Linear(input_dim, output_dim)
        self.std = StddevVector(output_dim, init_std)

    def params_from_mean(self, mean: Tensor) -> dict[str, Tensor]:
        return {"mean": mean, "std": self.std(mean)}

    # The reference's Distribution interface (distribution.py:49-178).  The agent does not go through it -- the actor fuses
    # the mean head into the trunk's autograd node -- but hooks and deployment code may call the distribution on a latent.
    def forward(self, backbone_feat: Tensor, **kwargs) -> dict[str, Tensor]:
        return self.params_from_mean(F.linear_head(backbone_feat, self.mean_head.weight, self.mean_head.bias))

    def sample(self, backbone_feat: Tensor, **kwargs) -> tuple[dict[str, Tensor], tuple[Tensor, Tensor]]:
        dist_params = self(backbone_feat, **kwargs)
        return dist_params, self.sample_from_dist(dist_params)

    def determine(self, backbone_feat: Tensor, **kwargs) -> Tensor:
        return self(backbone_feat, **kwargs)["mean"]

    def deterministic(self) -> nn.Module:
        return _DeterministicWrapper(self)

    # torch Normal arithmetic (distribution.py:195-218), kept in torch for the small rollout-time tensors
    @staticmethod
    def compute_logp(dist_params: dict[str, Tensor], sample: Tensor) -> Tensor:
        mean, std = dist_params["mean"], dist_params["std"]
        lp = -((sample - mean) ** 2) / (2 * std**2) - std.log() - LOG_SQRT_2PI
        return lp.sum(dim=-1, keepdim=True)

    @staticmethod
    def compute_entropy(dist_params: dict[str, Tensor]) -> Tensor:
        return (0.5 + 0.5 * math.log(2 * math.pi) + dist_params["std"].log()).sum(dim=-1, keepdim=True)

    @staticmethod
    def compute_kl_div(p: dict[str, Tensor], q: dict[str, Tensor]) -> Tensor:
        var_ratio = (p["std"] / q["std"]).pow(2)
        t1 = ((p["mean"] - q["mean"]) / q["std"]).pow(2)
        return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(dim=-1, keepdim=True)

    def sample_from_dist(self, dist_params: dict[str, Tensor]) -> tuple[Tensor, Tensor]:
        mean, std = dist_params["mean"], dist_params["std"]
        eps = standard_normal_like(mean)
        sample = mean + eps * std
        return sample, self.compute_logp(dist_params, sample)



@dataclass(slots=True)
class ActorFactory:
    backbone_factory: Any
    distribution_factory: Any
    latent_dim: int | None = None

    def __call__(self, input_dim: int | None = None, output_dim: int | None = None) -> "Actor":
        backbone = self.backbone_factory(input_dim, self.latent_dim)
        return Actor(backbone, self.distribution_factory(backbone.output_dim, output_dim))


class Actor(Module):
    """backbone -> distribution (reference actor.py:181-268)."""

    Factory = ActorFactory

    def __init__(self, backbone: Module, distribution: NormalDist):
        super().__init__(backbone.input_dim, distribution.output_dim, backbone.is_recurrent)
        self.backbone, self.distribution = backbone, distribution
        self.latent_dim = backbone.output_dim

    def _mean(self, observation: Tensor, memory=None, done=None):
        """backbone + mean head as ONE autograd node (trunk GEMMs on tcgen05, fp32 SIMT head)."""
        head = self.distribution.mean_head
        if self.backbone.is_recurrent:
            # LSTM backbone (K7) + head; rollout calls are single-step (the reference passes sequential=False there)
            latent, memory = self.backbone(observation, memory, done=done, sequential=observation.dim() >= 3)
            self.intermediate_repr["backbone.output"] = latent
            return F.linear_head(latent, head.weight, head.bias), memory
        lins = self.backbone.linears()
        if not self.backbone.ends_with_activation:
            raise ValueError("Actor expects an activation-terminated Mlp backbone (reference preset/ppo.py:137-140)")
        if not F.simt_head_supported(*head.weight.shape):
            # heads outside the fused SIMT shapes (e.g. 21 actions, or a latent width that is not 128 / 256): trunk node
            # + the general head node
            latent = self.backbone(observation)
            self.intermediate_repr["backbone.output"] = latent
            return F.linear_head(latent, head.weight, head.bias), memory
        mean, latent = F.mlp_head_forward(observation, [m.weight for m in lins], [m.bias for m in lins],
                                          self.backbone.activation, head.weight, head.bias)
        self.intermediate_repr["backbone.output"] = latent
        return mean, memory

    def forward(self, observation: Tensor, memory=None, done: Tensor | None = None, forward_type: str | None = "forward",
                deterministic: bool = False, **kw):
        """Router of the reference's actor (actor.py:70-99): ``forward`` -> (distribution parameters, memory), ``explore`` ->
        (parameters, (action, log-prob), memory), ``act`` / ``act_deterministic`` -> (action, memory)."""
        if forward_type == "forward":
            mean, memory = self._mean(observation, memory, done)
            return self.distribution.params_from_mean(mean), memory
        if forward_type == "explore":
            return self.explore(observation, memory, deterministic)
        if forward_type == "act":
            return self.act(observation, memory, deterministic)
        if forward_type == "act_deterministic":
            return self.act(observation, memory, True)
        raise ValueError(f"Unsupported 'forward_type' value: {forward_type!r}")

    def explore(self, observation: Tensor, memory=None, deterministic: bool = False, **kw):
        mean, memory = self._mean(observation, memory, None)
        dist_params = self.distribution.params_from_mean(mean)
        if deterministic:
            action = mean
            logp = self.distribution.compute_logp(dist_params, action)
        else:
            action, logp = self.distribution.sample_from_dist(dist_params)
        return dist_params, (action, logp), memory

    def act(self, observation: Tensor, memory=None, deterministic: bool = False, **kw):
        _, (action, _), memory = self.explore(observation, memory, deterministic, **kw)
        return action, memory

    def compute_logp(self, dist_params, action):
        return self.distribution.compute_logp(dist_params, action)

    def compute_entropy(self, dist_params):
        return self.distribution.compute_entropy(dist_params)

    def compute_kl_div(self, p, q):
        return self.distribution.compute_kl_div(p, q)

    def reset_memory(self, memory, done=None):
        return self.backbone.reset_memory(memory, done)

    def step_memory(self, observation, memory=None, **kwargs):
        return self.backbone.step_memory(observation, memory, **kwargs)


@dataclass(slots=True)
class ValueFactory:
    backbone_factory: Any
    latent_dim: int | None = None
    action_aware: bool = False

    def __call__(self, input_dim: int | None = None, output_dim: int | None = 1) -> "Value":
        if self.action_aware:
            raise ValueError("cusrl_b200.Value implements the state-value critic only (action_aware=False)")
        backbone = self.backbone_factory(input_dim, self.latent_dim)
        return Value(backbone, nn.Linear(backbone.output_dim, output_dim))


class Value(Module):
    """backbone -> fp32 linear value head (reference critic.py:70-89)."""

    Factory = ValueFactory

    def __init__(self, backbone: Module, value_head: nn.Linear):
        super().__init__(backbone.input_dim, value_head.out_features, backbone.is_recurrent)
        self.backbone, self.value_head = backbone, value_head
        self.action_aware = False

    def forward(self, state: Tensor, *, memory=None, done: Tensor | None = None, **kw):
        if self.backbone.is_recurrent:
            latent, memory = self.backbone(state, memory, done=done, sequential=state.dim() >= 3)
            self.intermediate_repr["backbone.output"] = latent
            return F.linear_head(latent, self.value_head.weight, self.value_head.bias), memory
        lins = self.backbone.linears()
        if not self.backbone.ends_with_activation:
            raise ValueError("Value expects an activation-terminated Mlp backbone (reference preset/ppo.py:143-147)")
        if not F.simt_head_supported(*self.value_head.weight.shape):
            latent = self.backbone(state)
            self.intermediate_repr["backbone.output"] = latent
            return F.linear_head(latent, self.value_head.weight, self.value_head.bias), memory
        value, latent = F.mlp_head_forward(state, [m.weight for m in lins], [m.bias for m in lins],
                                           self.backbone.activation, self.value_head.weight, self.value_head.bias)
        self.intermediate_repr["backbone.output"] = latent
        return value, memory

    def evaluate(self, state: Tensor, *, memory=None, done: Tensor | None = None, **kw) -> Tensor:
        return self(state, memory=memory, done=done, **kw)[0]

    def reset_memory(self, memory, done=None):
        return self.backbone.reset_memory(memory, done)

    def step_memory(self, state, memory=None, **kwargs):
        return self.backbone.step_memory(state, memory, **kwargs)
