"""The on-policy actor-critic agent: rollout (`act` / `step`) and the hook-driven update loop.

Interface-compatible with the reference's ``ActorCritic`` / ``ActorCriticFactory`` / ``HookList``
(cusrl/template/actor_critic.py:23-320, cusrl/template/agent.py:25-391): same constructor arguments, same
hook call order in ``update`` / ``_train_step``, same transition keys, numpy/torch I/O preservation in
``act``.  Differences are confined to the implementation: parameters live in a flat arena
(template/optimizer.py), the gradient allreduce is one in-place collective on that arena, metrics are
flushed with one device-to-host copy, and ``autocast`` / ``compile`` are refused.
"""

from __future__ import annotations

import os
from collections.abc import Iterable, Mapping
from contextlib import contextmanager, nullcontext
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .. import distributed
from ..environment import EnvironmentSpec
from ..metrics import Metrics
from ..runtime import device as resolve_device
from .buffer import Buffer, Sampler
from .hook import Hook, HookComposite

__all__ = ["ActorCritic", "ActorCriticFactory", "HookList"]


class HookList(list):
    """List of hooks addressable by name (reference actor_critic.py:23-62)."""

    def to_dict(self) -> dict[str, Hook]:
        return {hook.name: hook for hook in self}

    @classmethod
    def from_dict(cls, data: dict[str, Hook]) -> "HookList":
        return cls(hook.name_(name) for name, hook in data.items())

    def __getattr__(self, name: str) -> Any:
        for hook in self:
            if hook.name == name:
                return hook
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    @classmethod
    def coerce(cls, data: Any) -> "HookList":
        if isinstance(data, (cls, list, tuple)):
            return cls(data)
        if isinstance(data, dict):
            return cls.from_dict(data)
        raise TypeError(f"Unsupported hooks payload: {type(data)!r}")


@dataclass(kw_only=True)
class ActorCriticFactory:
    """Configuration that builds an :class:`ActorCritic` for an environment spec."""

    num_steps_per_update: int
    actor_factory: Any
    critic_factory: Any
    optimizer_factory: Any
    sampler: Sampler
    hooks: list
    name: str = "Agent"
    device: torch.device | str | None = None
    compile: bool | str = False
    autocast: bool | None | torch.dtype | str = False

    def __post_init__(self):
        self.hooks = HookList.coerce(self.hooks)

    def __call__(self, environment_spec: EnvironmentSpec) -> "ActorCritic":
        return ActorCritic(
            environment_spec=environment_spec, actor_factory=self.actor_factory, critic_factory=self.critic_factory,
            optimizer_factory=self.optimizer_factory, sampler=self.sampler, hooks=self.hooks,
            num_steps_per_update=self.num_steps_per_update, name=self.name, device=self.device,
            compile=self.compile, autocast=self.autocast)

    def from_environment(self, environment) -> "ActorCritic":
        return self(environment.spec)

    def register_hook(self, hook: Hook, index: int | None = None, before: str | None = None, after: str | None = None):
        """Insert a hook by index or relative to a named hook (reference actor_critic.py:97-137)."""
        if (index is not None) + (before is not None) + (after is not None) > 1:
            raise ValueError("Only one of index, before, or after can be specified")
        if before is not None:
            index = self.get_hook_index(before)
        elif after is not None:
            index = self.get_hook_index(after) + 1
        elif index is None:
            index = len(self.hooks)
        self.hooks.insert(index, hook)
        return self

    def get_hook(self, hook_name: str) -> Hook:
        return self.hooks[self.get_hook_index(hook_name)]

    def get_hook_index(self, hook_name: str) -> int:
        for i, hook in enumerate(self.hooks):
            if hook.name == hook_name:
                return i
        raise ValueError(f"No hook named '{hook_name}' is registered")


class ActorCritic:
    """PPO-family agent whose algorithm is an ordered list of hooks."""

    Factory = ActorCriticFactory
    MODULES = ["actor", "critic", "hook"]
    STATEFULS = ["optimizer", "grad_scaler"]

    def __init__(self, environment_spec: EnvironmentSpec, actor_factory, critic_factory, optimizer_factory,
                 sampler: Sampler, hooks: Iterable[Hook], num_steps_per_update: int, name: str = "Agent",
                 device: torch.device | str | None = None, compile: bool | str = False,
                 autocast: bool | None | torch.dtype | str = False):
        if compile not in (False, None):
            raise ValueError("cusrl_b200 runs hand-written sm_100a kernels: 'compile' must be False")
        if autocast not in (False, None):
            raise ValueError("cusrl_b200 computes in fp32 like the reference presets: 'autocast' must be False")
        self.environment_spec = environment_spec
        self.observation_dim = environment_spec.observation_dim
        self.action_dim = environment_spec.action_dim
        self.has_state = environment_spec.state_dim is not None
        self.state_dim = environment_spec.state_dim or self.observation_dim
        self.parallelism = environment_spec.num_instances
        self.value_dim = environment_spec.reward_dim
        self.num_steps_per_update = num_steps_per_update
        self.buffer_capacity = num_steps_per_update
        self.name = name
        self.device = resolve_device(device)
        self.compile, self.autocast_enabled, self.dtype = False, False, torch.float32
        self.inference_mode = False
        self.deterministic = False
        self.transition: dict[str, Any] = {}
        self.metrics = Metrics()
        self.iteration = 0
        self.step_index = 0
        # opt-in CUDA-graph replay of the train step (template/graphs.py); not part of the reference's surface
        self.cuda_graphs = os.environ.get("CUSRL_B200_CUDA_GRAPHS", "1") not in ("", "0")
        self._train_step_graphs = None
        # fused rollout step (template/rollout.py); CUSRL_B200_FUSED_ROLLOUT=0 forces the generic act / step flow
        self.fused_rollout = os.environ.get("CUSRL_B200_FUSED_ROLLOUT", "1") not in ("", "0")
        self._fused_rollout = None
        self.last_objectives: dict[str, torch.Tensor] | None = None

        self.actor_factory, self.critic_factory, self.optimizer_factory = actor_factory, critic_factory, optimizer_factory
        self.hook = HookComposite(hooks)
        self.hook.pre_init(self)
        # through the attributes: a hook's pre_init may have replaced a factory (SymmetricArchitecture, actor_critic.py:203-206)
        self.actor = self.actor_factory(self.observation_dim, self.action_dim)
        self.critic = self.critic_factory(self.state_dim, self.value_dim)
        self.buffer = Buffer(self.buffer_capacity, self.parallelism, device=self.device)
        self.sampler = sampler
        self.actor_memory = None
        self.hook.init()
        self.actor = self.setup_module(self.actor)
        self.critic = self.setup_module(self.critic)
        self.optimizer = self.optimizer_factory(self.named_parameters())
        # fp32 only: a DISABLED scaler, kept so that checkpoints carry the reference's "grad_scaler" entry (actor_critic.py:171,210)
        self.grad_scaler = torch.GradScaler(device=str(self.device), enabled=False)
        self._set_training_mode(False)
        self.hook.post_init()
        distributed.broadcast_parameters([self.optimizer.flat_param] if hasattr(self.optimizer, "flat_param")
                                         else self.parameters())
        self.hook.apply_schedule(0)
        self._configure_sampler()

    def _configure_sampler(self) -> None:
        """Minibatch gather of the network inputs.  With the f16x3 dense layers the sampler emits ``observation`` / ``state``
        minibatches directly as the fp16 pairs the first layers consume; the fp32 copy of those leaves (940 of the 1 004
        bytes a sample's gather moves) is then read by nobody -- PROVIDED every hook is one of this package's PPO hooks
        (which hand the batch leaf to the networks and nothing else) and both networks are MLPs the pair path covers.  Only
        then is the sampler told that it may skip that copy; any other hook (a user's, the reference's, observation
        normalisation, symmetry, ...) keeps the full gather."""
        from ..hook import on_policy as P
        from ..hook.auxiliary import RandomNetworkDistillation
        from ..nn import functional as F
        from ..nn import modules as M

        sampler = self.sampler
        if not hasattr(sampler, "pair_only"):
            return
        known = (P.ModuleInitialization, P.ValueComputation, P.GeneralizedAdvantageEstimation, P.AdvantageNormalization,
                 P.AdvantageReduction, P.ValueLoss, P.OnPolicyPreparation, P.PpoSurrogateLoss, P.EntropyLoss, P.GradientClipping,
                 P.OnPolicyStatistics, P.AdaptiveLRSchedule, RandomNetworkDistillation)
        ok = all(type(hook) in known for hook in self.hook)
        for net in (self.actor, self.critic):
            backbone = getattr(net, "backbone", None)
            ok = ok and type(net) in (M.Actor, M.Value) and type(backbone) is M.Mlp
            if ok:
                lins = backbone.linears()
                ok = F.f16x3_supported([m.weight for m in lins], [m.bias for m in lins])
        sampler.pair_only = frozenset(("observation", "state")) if ok else frozenset()

    # ---- parameters / modules --------------------------------------------------------------------
    def named_parameters(self):
        for name in self.MODULES:
            module = getattr(self, name, None)
            if module is not None:
                yield from module.named_parameters(prefix=name)

    def parameters(self):
        for _name, param in self.named_parameters():
            yield param

    def setup_module(self, module):
        return module.to(device=self.device)

    def autocast(self):
        return nullcontext()

    @property
    def grad_scaler_enabled(self) -> bool:
        return False

    def record(self, metrics: Mapping[str, Any] | None = None, /, **kwargs) -> None:
        self.metrics.record(metrics, **kwargs)

    def to_tensor(self, value) -> torch.Tensor:
        tensor = torch.as_tensor(value, device=self.device)
        return tensor.clone() if tensor is value else tensor

    def to_nested_tensor(self, value):
        if value is None:
            return None
        if isinstance(value, (tuple, list)):
            return tuple(self.to_nested_tensor(v) for v in value)
        if isinstance(value, Mapping):
            return {k: self.to_nested_tensor(v) for k, v in value.items()}
        return self.to_tensor(value)

    # ---- rollout ---------------------------------------------------------------------------------
    @torch.no_grad()
    def act(self, observation, state=None):
        """actor_critic.py:227-253; output array type follows the input (agent.py:376-391)."""
        if self.fused_rollout:
            if self._fused_rollout is None:
                from .rollout import make_fused_rollout

                self._fused_rollout = make_fused_rollout(self)
            action = self._fused_rollout.act(observation, state)
            if action is not None:
                return action
        self.transition.clear()
        self._save_transition(observation=observation, state=state)
        self.hook.pre_act(self.transition)
        action_dist, (action, action_logp), next_memory = self.actor.explore(
            self.transition["observation"], memory=self.actor_memory, deterministic=self.deterministic)
        self._save_transition(actor_memory=self.actor_memory, action_dist=action_dist, action=action,
                              action_logp=action_logp)
        self.actor_memory = next_memory
        self.hook.post_act(self.transition)
        action = self.transition["action"]
        if isinstance(observation, np.ndarray):
            out = action.cpu().numpy()
            return out.astype(observation.dtype) if np.issubdtype(out.dtype, np.floating) else out
        return action.to(device=observation.device, dtype=observation.dtype if torch.is_floating_point(action) else None)

    @torch.no_grad()
    def step(self, next_observation, reward, terminated, truncated, next_state=None, **kwargs) -> bool:
        """actor_critic.py:255-291."""
        if self._fused_rollout is not None and self._fused_rollout.step(next_observation, reward, terminated, truncated,
                                                                        next_state, kwargs):
            self.step_index += 1
            return self.step_index >= self.num_steps_per_update and self.hook.should_update(self.transition)
        self._save_transition(next_observation=next_observation, next_state=next_state, reward=reward,
                              terminated=terminated, truncated=truncated, **kwargs)
        if self.transition["terminated"].dtype != torch.bool:
            raise TypeError("'terminated' must have dtype bool")
        if self.transition["truncated"].dtype != torch.bool:
            raise TypeError("'truncated' must have dtype bool")
        self.transition["done"] = self.transition["terminated"] | self.transition["truncated"]
        self.hook.post_step(self.transition)
        if not self.inference_mode:
            self.buffer.push(self.transition)
        self.actor.reset_memory(self.actor_memory, self.transition["done"])
        if self.inference_mode:
            return False
        self.step_index += 1
        return self.step_index >= self.num_steps_per_update and self.hook.should_update(self.transition)

    # ---- update ----------------------------------------------------------------------------------
    def update(self) -> dict[str, float]:
        """actor_critic.py:293-300."""
        self.hook.pre_update(self.buffer)
        with self._training_mode():
            for metadata, batch in self.sampler(self.buffer):
                self._train_step(metadata, batch)
        self.hook.post_update()
        self.hook.apply_schedule(self.iteration + 1)
        self.step_index = 0
        self.iteration += 1
        summary = self.metrics.summary(self.name)
        self.metrics.clear()
        return summary

    def _train_step(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None:
        """actor_critic.py:302-320 (no GradScaler: fp32 only).  With ``cuda_graphs`` the two halves of the step are
        captured once per batch layout and replayed (template/graphs.py); the gradient allreduce between them always
        runs eagerly, so the multi-GPU data path is the same in both modes."""
        if self.cuda_graphs and self.device.type == "cuda":
            if self._train_step_graphs is None:
                from .graphs import TrainStepGraphs

                self._train_step_graphs = TrainStepGraphs(self)
            self._train_step_graphs(metadata, batch)
        else:
            self._train_step_eager(metadata, batch)

    def _train_step_eager(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None:
        objectives = self._train_step_forward_backward(metadata, batch)
        if objectives is not None:
            distributed.reduce_gradients(self.optimizer)
        self._train_step_optimize(metadata, batch, objectives)
        # detached: a live reference to the losses would keep the step's autograd graph (and the parameters' AccumulateGrad
        # nodes, with the stream they were created on) alive into the next step
        self.last_objectives = None if objectives is None else {k: v.detach() for k, v in objectives.items()}

    def _train_step_forward_backward(self, metadata: dict[str, Any], batch: dict[str, Any]):
        """First half of the step: objectives of every hook, their sum in dict order, backward into the flat arena."""
        self.actor.clear_intermediate_repr()
        self.critic.clear_intermediate_repr()
        self.hook.pre_objective(metadata, batch)
        objectives = self.hook.objective(metadata, batch)
        if objectives is not None:
            loss = sum(objectives.values())
            self.optimizer.zero_grad()
            loss.backward()
            arena = getattr(self.optimizer, "arena", None)
            if arena is not None:
                # a hook that detached a p.grad from the flat arena must not make the allreduce / clip / Adam kernels
                # work on stale memory: such gradients are folded back into the arena here (pointer compares only)
                arena.rebind_gradients()
        return objectives

    def _train_step_optimize(self, metadata: dict[str, Any], batch: dict[str, Any], objectives) -> None:
        """Second half, after the cross-rank gradient average: clip, Adam, records, post-objective hooks."""
        if objectives is not None:
            self.hook.pre_optim(self.optimizer)
            self.optimizer.step()
            self.hook.post_optim()
            self.record(**objectives)
        self.hook.post_objective(metadata, batch)

    def resize_buffer(self, capacity: int) -> None:
        """actor_critic.py:327-330: a different rollout length; the leaves are re-allocated by the next pushes."""
        if self.buffer_capacity != capacity:
            self.buffer_capacity = capacity
            self.buffer.resize(capacity)

    def set_inference_mode(self, mode: bool = True, deterministic: bool | None = True) -> None:
        self.inference_mode = mode
        if deterministic is not None:
            self.deterministic = mode and deterministic

    def set_iteration(self, iteration: int) -> None:
        if iteration < 0:
            raise ValueError("Iteration must be non-negative")
        if iteration != self.iteration:
            self.iteration = iteration
            self.hook.apply_schedule(iteration)

    # ---- checkpointing (agent.py:283-330) ---------------------------------------------------------
    def state_dict(self) -> dict[str, Any]:
        out = {}
        for name in self.MODULES + self.STATEFULS:
            obj = getattr(self, name, None)
            if obj is not None:
                out[name] = obj.state_dict()
        return out

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        keys = set(state_dict)
        for name in self.MODULES + self.STATEFULS:
            obj = getattr(self, name, None)
            if obj is None:
                continue
            if (state := state_dict.get(name)) is None:
                self.warn(f"No state_dict entry was found for '{name}'")
                continue
            keys.discard(name)
            try:
                obj.load_state_dict(state)
            except (RuntimeError, ValueError) as error:
                self.warn(f"Mismatched state_dict for '{name}': {error}")
        if keys:
            self.warn(f"Unused state_dict keys: {keys}.")

    def export(self, output_dir: str, *, target_format: str = "onnx", with_environment_normalization: bool = True,
               optimize: bool = True, sequence_len: int = 1, batch_size: int = 1, opset_version: int | None = None,
               dynamo: bool = False, verbose: bool = True, **kwargs) -> None:
        """actor_critic.py:332-418: the deployed (deterministic) policy as ONNX / TorchScript (template/export.py)."""
        from .export import export_agent

        export_agent(self, output_dir, target_format=target_format, with_environment_normalization=with_environment_normalization,
                     optimize=optimize, sequence_len=sequence_len, batch_size=batch_size, opset_version=opset_version,
                     dynamo=dynamo, verbose=verbose, **kwargs)

    @classmethod
    def warn(cls, message: str) -> None:
        if distributed.is_main_process():
            print(f"\033[1;33mAgent: {message}\033[0m")

    # ---- internals -------------------------------------------------------------------------------
    def _save_transition(self, **kwargs) -> None:
        for key, value in kwargs.items():
            if value is None:
                continue
            try:
                self.transition[key] = self.to_nested_tensor(value)
            except Exception as error:
                raise ValueError(f"Failed to convert transition field '{key}' to a tensor") from error

    def _set_training_mode(self, mode: bool = True) -> None:
        for name in self.MODULES:
            module = getattr(self, name, None)
            if module is not None:
                module.train(mode)

    @contextmanager
    def _training_mode(self):
        self._set_training_mode(True)
        try:
            yield
        finally:
            self._set_training_mode(False)
