"""Auxiliary hooks on the hot path: Random Network Distillation (reference cusrl/hook/auxiliary/rnd.py:15-81)."""

from __future__ import annotations

import itertools

import torch
from torch import nn

from .. import ops
from ..template.buffer import Buffer
from ..template.hook import Hook

__all__ = ["RandomNetworkDistillation"]


class _MseLoss(torch.autograd.Function):
    """MSELoss(prediction, target) with the gradient produced by the forward kernel (K5)."""

    @staticmethod
    def forward(ctx, prediction, target):
        loss, grad = ops.mse_loss(prediction.contiguous(), target.contiguous(), want_grad=True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return ops.scale_(grad, g.reshape(1).contiguous()), None


class RandomNetworkDistillation(Hook):
    """Intrinsic reward from Random Network Distillation (https://arxiv.org/abs/1810.12894).

    Same constructor and behaviour as the reference hook: two ``module_factory(input_dim, output_dim)`` networks with
    Xavier-normal weights and zero biases, the target frozen (rnd.py:54-66); ``pre_update`` adds
    ``reward_scale * mean_d (target - prediction)^2`` of the next state to ``buffer["reward"]`` and records
    ``rnd_reward`` (rnd.py:68-75); ``objective`` returns ``{"rnd_loss": MSE(prediction, target)}`` on the minibatch
    (rnd.py:77-81).  Register it ``before="value_computation"`` so GAE sees the augmented reward."""

    def __init__(self, module_factory, output_dim: int, reward_scale: float, state_indices=None):
        super().__init__()
        self.output_dim = output_dim
        self.module_factory = module_factory
        self.state_indices = slice(None) if state_indices is None else state_indices
        self.reward_scale = reward_scale
        self.register_mutable("reward_scale")

    def init(self) -> None:
        input_dim = torch.ones(1, self.agent.state_dim)[..., self.state_indices].numel()
        target = self.module_factory(input_dim, self.output_dim)
        predictor = self.module_factory(input_dim, self.output_dim)
        for module in itertools.chain(target.modules(), predictor.modules()):
            if isinstance(module, nn.Linear):
                nn.init.xavier_normal_(module.weight)
                nn.init.zeros_(module.bias)
        self.register_module("target", target)
        self.register_module("predictor", predictor)
        self.target.requires_grad_(False)
        # the minibatch sampler only gathers the leaves the objective consumes: this hook consumes the next state
        sampler = getattr(self.agent, "sampler", None)
        if getattr(sampler, "fields", None) is not None:
            sampler.fields = tuple(dict.fromkeys((*sampler.fields, "next_state", "next_observation")))
            if getattr(sampler, "_impl", None) is not None:
                sampler._impl = None

    def _next_state(self, data) -> torch.Tensor:
        x = data.get("next_state")
        if x is None:
            x = data["next_observation"]
        return x if self.state_indices == slice(None) else x[..., self.state_indices]

    @torch.no_grad()
    def pre_update(self, buffer: Buffer) -> None:
        x = self._next_state(buffer)
        target, prediction = self.target(x), self.predictor(x)
        rnd_reward, mean = ops.rnd_reward_(target.contiguous(), prediction.contiguous(), buffer["reward"], self.reward_scale)
        self.agent.metrics.record_mean("rnd_reward", mean, rnd_reward.numel())

    def objective(self, metadata, batch):
        x = self._next_state(batch)
        with torch.no_grad():
            target = self.target(x)
        return {"rnd_loss": _MseLoss.apply(self.predictor(x), target)}
