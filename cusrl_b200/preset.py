"""PPO preset: the hook order and hyper-parameters that define "PPO" (reference cusrl/preset/ppo.py:19-182),
with the Anymal-C-rough values of cusrl/zoo/isaaclab/locomotion.py:48-59 available as ``anymal_c_rough_ppo``."""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass, field

import torch

from . import hook as H
from .nn import Actor, Mlp, NormalDist, Rnn, Value
from .sampler import AutoMiniBatchSampler
from .template import ActorCritic, ActorCriticFactory, AdamFactory

__all__ = ["PpoAgentFactory", "RecurrentPpoAgentFactory", "anymal_c_rough_ppo", "ppo_hook_suite", "PPO_MINIBATCH_FIELDS",
           "RECURRENT_PPO_MINIBATCH_FIELDS"]

# leaves the PPO objective consumes: everything else in the buffer is not gathered
PPO_MINIBATCH_FIELDS = ("observation", "state", "action", "action_logp", "advantage", "return", "value", "done")


RECURRENT_PPO_MINIBATCH_FIELDS = PPO_MINIBATCH_FIELDS + ("actor_memory", "critic_memory")


def ppo_hook_suite(
    orthogonal_init: bool = True, normalize_observation: bool = False, gae_gamma: float = 0.99, gae_lamda: float = 0.95,
    gae_lamda_value: float | None = None, normalize_advantage: bool = True, value_loss_weight: float = 0.5,
    value_loss_clip: float | None = None, surrogate_clip_ratio: float = 0.2, surrogate_loss_weight: float = 1.0,
    entropy_loss_weight: float = 0.01, max_grad_norm: float | None = 1.0, grad_clip_groups: dict[str, float] | None = None,
    desired_kl_divergence: float | None = None, max_kl_divergence: float | None = None, empty_cuda_cache: bool = False,
) -> list:
    """Same order as the reference's ``ppo_hook_suite`` (preset/ppo.py:37-65)."""
    hooks = [
        H.ModuleInitialization(init_actor=orthogonal_init, init_critic=orthogonal_init),
        H.ObservationNormalization() if normalize_observation else None,
        H.ValueComputation(),
        H.GeneralizedAdvantageEstimation(gamma=gae_gamma, lamda=gae_lamda, lamda_value=gae_lamda_value),
        H.AdvantageNormalization() if normalize_advantage else None,
        H.ValueLoss(weight=value_loss_weight, loss_clip=value_loss_clip),
        H.OnPolicyPreparation(),
        H.PpoSurrogateLoss(clip_ratio=surrogate_clip_ratio, weight=surrogate_loss_weight),
        H.EntropyLoss(weight=entropy_loss_weight),
        H.GradientClipping(max_grad_norm, grad_clip_groups),
        H.OnPolicyStatistics(sampler=AutoMiniBatchSampler()),
        (H.AdaptiveLRSchedule(desired_kl_divergence, max_kl_divergence=max_kl_divergence)
         if desired_kl_divergence is not None else None),
        H.EmptyCudaCache() if empty_cuda_cache else None,
    ]
    return [h for h in hooks if h is not None]


@dataclass(kw_only=True)
class PpoAgentFactory:
    """Field-for-field the reference's ``PpoAgentFactory`` (preset/ppo.py:77-130) for continuous actions."""

    num_steps_per_update: int = 24
    actor_hidden_dims: Sequence[int] = (256, 128)
    critic_hidden_dims: Sequence[int] = (256, 128)
    activation_fn: str = "ReLU"
    action_space_type: str = "continuous"
    lr: float = 2e-4
    sampler_epochs: int = 5
    sampler_mini_batches: int = 4
    orthogonal_init: bool = True
    init_distribution_std: float | None = None
    normalize_observation: bool = False
    gae_gamma: float = 0.99
    gae_lamda: float = 0.95
    gae_lamda_value: float | None = None
    normalize_advantage: bool = True
    value_loss_weight: float = 0.5
    value_loss_clip: float | None = None
    surrogate_clip_ratio: float = 0.2
    surrogate_loss_weight: float = 1.0
    entropy_loss_weight: float = 0.01
    max_grad_norm: float | None = 1.0
    grad_clip_groups: dict[str, float] = field(default_factory=dict)
    desired_kl_divergence: float | None = None
    max_kl_divergence: float | None = None
    name: str = "Agent"
    device: torch.device | str | None = None
    compile: bool | str = False
    autocast: bool | None | torch.dtype | str = False

    def to_underlying(self) -> ActorCriticFactory:
        if self.action_space_type != "continuous":
            raise ValueError("cusrl_b200 implements the continuous (NormalDist) policy head of the PPO preset")
        return ActorCriticFactory(
            num_steps_per_update=self.num_steps_per_update,
            actor_factory=Actor.Factory(
                backbone_factory=Mlp.Factory(hidden_dims=self.actor_hidden_dims, activation_fn=self.activation_fn,
                                             ends_with_activation=True),
                distribution_factory=NormalDist.Factory(init_std=self.init_distribution_std)),
            critic_factory=Value.Factory(
                backbone_factory=Mlp.Factory(hidden_dims=self.critic_hidden_dims, activation_fn=self.activation_fn,
                                             ends_with_activation=True)),
            optimizer_factory=AdamFactory(defaults={"lr": self.lr}),
            sampler=AutoMiniBatchSampler(num_epochs=self.sampler_epochs, num_mini_batches=self.sampler_mini_batches,
                                         fields=PPO_MINIBATCH_FIELDS),
            hooks=ppo_hook_suite(
                orthogonal_init=self.orthogonal_init, normalize_observation=self.normalize_observation, gae_gamma=self.gae_gamma, gae_lamda=self.gae_lamda,
                gae_lamda_value=self.gae_lamda_value, normalize_advantage=self.normalize_advantage,
                value_loss_weight=self.value_loss_weight, value_loss_clip=self.value_loss_clip,
                surrogate_clip_ratio=self.surrogate_clip_ratio, surrogate_loss_weight=self.surrogate_loss_weight,
                entropy_loss_weight=self.entropy_loss_weight, max_grad_norm=self.max_grad_norm,
                grad_clip_groups=self.grad_clip_groups, desired_kl_divergence=self.desired_kl_divergence,
                max_kl_divergence=self.max_kl_divergence),
            name=self.name, device=self.device, compile=self.compile, autocast=self.autocast)

    def __call__(self, environment_spec) -> ActorCritic:
        return self.to_underlying()(environment_spec)

    def from_environment(self, environment) -> ActorCritic:
        return self(environment.spec)


def anymal_c_rough_ppo(**overrides) -> PpoAgentFactory:
    """Isaac-Velocity-Rough-Anymal-C-v0 PPO preset (reference zoo/isaaclab/locomotion.py:48-59)."""
    kwargs = dict(num_steps_per_update=24, actor_hidden_dims=(512, 256, 128), critic_hidden_dims=(512, 256, 128),
                  activation_fn="ELU", lr=1e-3, sampler_epochs=5, sampler_mini_batches=4, orthogonal_init=False,
                  entropy_loss_weight=0.005, desired_kl_divergence=0.015)
    kwargs.update(overrides)
    return PpoAgentFactory(**kwargs)


@dataclass(kw_only=True)
class RecurrentPpoAgentFactory(PpoAgentFactory):
    """The reference's ``RecurrentPpoAgentFactory`` (preset/ppo.py:185-298): LSTM 2 x 256 actor and critic."""

    rnn_type: str = "LSTM"
    actor_num_layers: int = 2
    actor_hidden_size: int = 256
    critic_num_layers: int = 2
    critic_hidden_size: int = 256
    # The reference defaults this to True (preset/ppo.py:241-243): its recurrent update re-packs variable-length sequences
    # every minibatch and fragments the caching allocator.  Every tensor of this implementation's update has a static shape
    # (episode boundaries are in-line resets), so the default is off; True appends the same `EmptyCudaCache` hook.
    empty_cuda_cache: bool = False

    def to_underlying(self) -> ActorCriticFactory:
        base = super().to_underlying()
        if self.empty_cuda_cache:
            base.register_hook(H.EmptyCudaCache())
        base.actor_factory = Actor.Factory(
            backbone_factory=Rnn.Factory(self.rnn_type, num_layers=self.actor_num_layers, hidden_size=self.actor_hidden_size),
            distribution_factory=NormalDist.Factory(init_std=self.init_distribution_std))
        base.critic_factory = Value.Factory(
            backbone_factory=Rnn.Factory(self.rnn_type, num_layers=self.critic_num_layers, hidden_size=self.critic_hidden_size))
        base.sampler = AutoMiniBatchSampler(num_epochs=self.sampler_epochs, num_mini_batches=self.sampler_mini_batches,
                                            fields=RECURRENT_PPO_MINIBATCH_FIELDS, memory_first_step_only=True)
        return base
