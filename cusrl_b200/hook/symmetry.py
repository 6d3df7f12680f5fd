"""Symmetry hooks under the reference's names: ``MirrorDef``, ``TransitionMirroring``,
``MirrorSymmetryLoss``, ``SymmetricDataAugmentation``, ``SymmetricArchitecture`` (+ ``SymmetricActor``).

Reference: cusrl/hook/auxiliary/symmetry.py:30-508.  Same constructor arguments, hook identities, mutable attributes,
transition / batch keys, loss names, errors.  What changes is where the arithmetic runs:

* a ``MirrorDef`` (index-permute + sign-flip of the last dim) applied to a CUDA tensor is ONE launch of
  ``mirror_rows_kernel`` instead of a fancy-index gather plus a multiply, bit-identical;
* the ``[N, 1 + V, C]`` tensors ``SymmetricDataAugmentation`` stores every environment step (the original next to its
  mirrored variants) are written by one launch per tensor in their final layout, instead of gather + multiply + movedim +
  unsqueeze + cat (5 passes over the observation, twice per step);
* the doubled minibatch then flows through the same gather (K8), dense-layer (K6) and fused objective (K4) kernels as any
  other batch -- leading dims are flattened into GEMM rows; the per-sample leaves (``advantage``, ``action_logp``,
  ``value``, ``return``) are repeated along the new axis as the reference does.

Mirror functions that are arbitrary callables (anything that is not a ``MirrorDef``, here or the reference's own) are
simply called, as in the reference.  Gradients through a ``MirrorDef`` (``MirrorSymmetryLoss`` mirrors the action mean of
the mirrored observation) run on the same kernel: the adjoint of an index permutation with sign flips is another one.
CPU tensors (the spec's mirror definitions are built and self-checked on the CPU, cusrl_test/_helpers.py:17-35) and index
maps that are not permutations take torch indexing.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from dataclasses import dataclass
from typing import Any

import torch
from torch import Tensor

from .. import ops
from ..nn import modules as M
from ..template.hook import Hook

__all__ = ["MirrorDef", "MirrorSymmetryLoss", "SymmetricActor", "SymmetricActorFactory", "SymmetricArchitecture",
           "SymmetricDataAugmentation", "TransitionMirroring"]

MirrorFn = Callable[[Tensor], Tensor]


class MirrorDef:
    """``mirror(x) = x[..., destination_indices] * multiplier`` with ``multiplier = -1`` at ``flipped_indices`` (an index
    list or a boolean mask, as torch indexing takes them), else ``+1``.  Reference symmetry.py:30-64."""

    def __init__(self, destination_indices: Sequence[int], flipped_indices: Sequence[int]):
        self.destination_indices = destination_indices
        self.flipped_indices = flipped_indices
        self.destination = torch.tensor(destination_indices, dtype=torch.long)
        self.multiplier = torch.ones(len(destination_indices))
        self.multiplier[flipped_indices] = -1.0
        n = self.destination.numel()
        if n and (int(self.destination.min()) < -n or int(self.destination.max()) >= n):
            raise IndexError(f"'destination_indices' must index a tensor of width {n}")
        self._tables: dict[torch.device, tuple[Tensor, Tensor]] = {}

    def tables(self, device: torch.device) -> tuple[Tensor, Tensor]:
        """(int32 [1, C] source index, fp32 [1, C] multiplier) on `device` for the kernel."""
        cached = self._tables.get(device)
        if cached is None:
            n = self.destination.numel()
            dest = (self.destination % n).to(device=device, dtype=torch.int32).reshape(1, n).contiguous()
            mult = self.multiplier.to(device=device, dtype=torch.float32).reshape(1, n).contiguous()
            cached = self._tables[device] = (dest, mult)
        return cached

    def inverse_tables(self, device: torch.device) -> tuple[Tensor, Tensor] | None:
        """Tables of the adjoint transform (gradient of the output -> gradient of the input) when the index map is a
        permutation: ``g_in[dest[j]] = g_out[j] * mult[j]``, i.e. another index-permute + sign-flip.  None otherwise."""
        cached = self._tables.get(("inv", device))  # type: ignore[arg-type]
        if cached is None:
            n = self.destination.numel()
            dest = (self.destination % n).tolist()
            if sorted(dest) != list(range(n)):
                cached = (None, None)
            else:
                inv = [0] * n
                for j, d in enumerate(dest):
                    inv[d] = j
                inv_t = torch.tensor(inv, dtype=torch.int32, device=device).reshape(1, n)
                mult = self.multiplier.to(device=device, dtype=torch.float32)[torch.tensor(inv, device=device)].reshape(1, n).contiguous()
                cached = (inv_t.contiguous(), mult)
            self._tables[("inv", device)] = cached  # type: ignore[index]
        return None if cached[0] is None else cached

    def __call__(self, input: Tensor) -> Tensor:
        if _kernel_ok(input, self):
            dest, mult = self.tables(input.device)
            return ops.mirror_rows(input, dest, mult, layout="same")
        if _kernel_ok(input, self, allow_grad=True) and self.inverse_tables(input.device) is not None:
            return _MirrorFunction.apply(input, self)   # differentiable, still the kernel (forward and adjoint)
        self.destination = self.destination.to(input.device)
        self.multiplier = self.multiplier.to(dtype=input.dtype, device=input.device)
        return input[..., self.destination] * self.multiplier

    def __repr__(self) -> str:
        return f"MirrorDef(destination_indices={self.destination_indices}, flipped_indices={self.flipped_indices})"


def _as_mirror_def(mirror: Any) -> MirrorDef | None:
    """`mirror` as a kernel-backed MirrorDef: ours, or the reference's own class (same two tensors, duck-typed) that an
    environment adapter of the reference put on the spec."""
    if isinstance(mirror, MirrorDef):
        return mirror
    dest, mult = getattr(mirror, "destination", None), getattr(mirror, "multiplier", None)
    if isinstance(dest, Tensor) and isinstance(mult, Tensor) and dest.dim() == 1 and dest.shape == mult.shape:
        twin = getattr(mirror, "_b200_twin", None)
        if twin is None:
            twin = MirrorDef(dest.tolist(), (mult < 0).tolist())
            try:
                mirror._b200_twin = twin
            except AttributeError:
                pass
        return twin
    return None


def _kernel_ok(x: Tensor, mirror: MirrorDef, allow_grad: bool = False) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 1 and x.shape[-1] == mirror.destination.numel()
            and x.numel() > 0 and (allow_grad or not (torch.is_grad_enabled() and x.requires_grad)))


class _MirrorFunction(torch.autograd.Function):
    """``mirror(x)`` with autograd: forward and adjoint are both index-permute + sign-flip transforms, both on the kernel."""

    @staticmethod
    def forward(ctx, x, mirror):
        ctx.mirror = mirror
        return ops.mirror_rows(x.detach(), *mirror.tables(x.device), layout="same")

    @staticmethod
    def backward(ctx, grad):
        return ops.mirror_rows(grad.contiguous(), *ctx.mirror.inverse_tables(grad.device), layout="same"), None


def _identity_plus(mirror: MirrorDef, device: torch.device) -> tuple[Tensor, Tensor]:
    """Tables of [identity, mirror]: the augmented layout's variants."""
    cached = mirror._tables.get(("aug", device))  # type: ignore[arg-type]
    if cached is None:
        dest, mult = mirror.tables(device)
        n = dest.shape[1]
        ident = torch.arange(n, dtype=torch.int32, device=device).reshape(1, n)
        cached = (torch.cat([ident, dest]).contiguous(), torch.cat([torch.ones_like(mult), mult]).contiguous())
        mirror._tables[("aug", device)] = cached  # type: ignore[index]
    return cached


class _SymmetryHook(Hook):
    """Reads the mirror functions off the environment spec (symmetry.py:67-95)."""

    mirror_observation: MirrorFn
    mirror_state: MirrorFn | None
    mirror_action: MirrorFn

    def init(self) -> None:
        spec = self.agent.environment_spec
        if getattr(spec, "mirror_observation", None) is None:
            raise ValueError("'mirror_observation' must be defined for symmetry hooks")
        self.mirror_observation = spec.mirror_observation
        if self.agent.has_state and getattr(spec, "mirror_state", None) is None:
            raise ValueError("'mirror_state' must be defined for symmetry hooks")
        self.mirror_state = getattr(spec, "mirror_state", None)
        if getattr(spec, "mirror_action", None) is None:
            raise ValueError("'mirror_action' must be defined for symmetry hooks")
        self.mirror_action = spec.mirror_action

    @staticmethod
    def _build_mirrored(original: Tensor, mirror: MirrorFn) -> Tensor:
        """``[V, *original.shape]``: every mirrored variant of `original` (symmetry.py:84-95)."""
        definition = _as_mirror_def(mirror)
        if definition is not None and _kernel_ok(original, definition):
            return ops.mirror_rows(original, *definition.tables(original.device), layout="stacked")
        mirrored = mirror(original)
        if mirrored.shape[1:] == original.shape:
            return mirrored
        if mirrored.shape[1:] == original.shape[1:]:
            return mirrored.reshape(-1, *original.shape)
        shape = ", ".join(str(s) for s in original.shape)
        raise ValueError(f"Mirrored tensor has incompatible shape: expected (N * {shape}) or (N, {shape}), "
                         f"got {mirrored.shape}")

    def _extend_sampler_fields(self, *names: str) -> None:
        """The presets gather only the leaves the PPO objective consumes (K8); a hook that adds transition fields the
        objective reads must add them to that selection."""
        seen: set[int] = set()
        stack = [getattr(self.agent, "sampler", None)]
        while stack:
            sampler = stack.pop()
            if sampler is None or id(sampler) in seen:
                continue
            seen.add(id(sampler))
            fields = getattr(sampler, "fields", None)
            if fields is not None:
                sampler.fields = tuple(fields) + tuple(n for n in names if n not in fields)
            stack.append(getattr(sampler, "_impl", None))


class TransitionMirroring(_SymmetryHook):
    """Collects the rollout in ONE mirrored frame: the actor sees mirrored observations, the sampled action is mirrored
    back for the environment, and the stored transition is rewritten consistently (symmetry.py:98-159)."""

    def __init__(self, index: int = 0):
        if not isinstance(index, int):
            raise TypeError("'index' must be an int")
        super().__init__()
        self.index = index

    def pre_act(self, transition) -> None:
        transition["observation"] = self._select_mirrored_tensor(transition["observation"], self.mirror_observation, self.index)
        if (state := transition.get("state")) is not None:
            assert self.mirror_state is not None
            transition["state"] = self._select_mirrored_tensor(state, self.mirror_state, self.index)

    def post_act(self, transition) -> None:
        transition["action"] = self._select_mirrored_tensor(transition["action"], self.mirror_action, self.index)

    def post_step(self, transition) -> None:
        transition["next_observation"] = self._select_mirrored_tensor(
            transition["next_observation"], self.mirror_observation, self.index)
        if (next_state := transition.get("next_state")) is not None:
            assert self.mirror_state is not None
            transition["next_state"] = self._select_mirrored_tensor(next_state, self.mirror_state, self.index)

    @classmethod
    def _select_mirrored_tensor(cls, original: Tensor, mirror: MirrorFn, index: int) -> Tensor:
        definition = _as_mirror_def(mirror)
        if definition is not None:  # exactly one variant: skip the stacked intermediate
            if not -1 <= index < 1:
                raise IndexError(f"Mirror index {index} is out of range for 1 symmetry transforms")
            return definition(original)
        mirrored = cls._build_mirrored(original, mirror)
        num_symmetries = mirrored.shape[0]
        if not -num_symmetries <= index < num_symmetries:
            raise IndexError(f"Mirror index {index} is out of range for {num_symmetries} symmetry transforms")
        return mirrored[index]


class MirrorSymmetryLoss(_SymmetryHook):
    """MSE between the action distribution at an observation and the mirrored distribution at the mirrored observation
    ("Learning Symmetric and Low-Energy Locomotion"); symmetry.py:162-232.  Losses ``action_mean_symmetry_loss`` (and
    ``action_std_symmetry_loss``), each scaled by ``weight``."""

    def __init__(self, weight: float | None, symmetrize_action_std: bool = False):
        if weight is not None and weight < 0:
            raise ValueError("'weight' must be None or non-negative")
        super().__init__()
        self.symmetrize_action_std = symmetrize_action_std
        self.weight: float | None = weight
        self.register_mutable("weight")
        self.mirrored_actor_memory = None

    def init(self) -> None:
        super().init()
        self.mirrored_actor_memory = None
        self._extend_sampler_fields("mirrored_actor_memory")

    @torch.no_grad()
    def post_step(self, transition) -> None:
        actor = self.agent.actor
        mirrored_observation = self.mirror_observation(transition["observation"])
        transition["mirrored_actor_memory"] = self.mirrored_actor_memory
        self.mirrored_actor_memory = actor.step_memory(mirrored_observation, memory=self.mirrored_actor_memory)
        actor.reset_memory(self.mirrored_actor_memory, transition["done"])

    def objective(self, metadata, batch):
        if self.weight is None:
            return None
        mirrored_dist, _ = self.agent.actor(self.mirror_observation(batch["observation"]),
                                            memory=batch.get("mirrored_actor_memory"), done=batch["done"])
        current = batch["curr_action_dist"]
        losses = {"action_mean_symmetry_loss":
                  torch.nn.functional.mse_loss(current["mean"], self.mirror_action(mirrored_dist["mean"])) * self.weight}
        if self.symmetrize_action_std:
            losses["action_std_symmetry_loss"] = torch.nn.functional.mse_loss(
                current["std"], self.mirror_action(mirrored_dist["std"]).abs()) * self.weight
        return losses


class SymmetricDataAugmentation(_SymmetryHook):
    """Appends the mirrored twin(s) of every transition to the training batch ("Symmetry Considerations for Learning Task
    Symmetric Robot Policies"); symmetry.py:235-339.  Stored per step: ``augmented_observation`` /
    ``augmented_next_observation`` (/ ``_state``) / ``augmented_action`` of shape ``[N, 1 + V, C]`` and, for recurrent
    networks, the memories of the mirrored streams."""

    def __init__(self, augments_value: bool = True):
        self.augments_value = augments_value
        super().__init__(training_only=True)
        self.mirrored_actor_memory = None
        self.mirrored_critic_memory = None

    def init(self) -> None:
        super().init()
        self.mirrored_actor_memory = None
        self.mirrored_critic_memory = None
        self._extend_sampler_fields("augmented_observation", "augmented_next_observation", "augmented_state",
                                    "augmented_next_state", "augmented_action", "augmented_actor_memory",
                                    "augmented_critic_memory")

    @torch.no_grad()
    def post_step(self, transition) -> None:
        mirrored_observation, transition["augmented_observation"] = self._build_augmented_tensor(
            transition["observation"], self.mirror_observation)
        _, transition["augmented_next_observation"] = self._build_augmented_tensor(
            transition["next_observation"], self.mirror_observation, need_mirrored=False)
        if (state := transition.get("state")) is not None:
            assert self.mirror_state is not None
            mirrored_state, transition["augmented_state"] = self._build_augmented_tensor(state, self.mirror_state)
            _, transition["augmented_next_state"] = self._build_augmented_tensor(
                transition["next_state"], self.mirror_state, need_mirrored=False)
        else:
            mirrored_state = mirrored_observation
        _, transition["augmented_action"] = self._build_augmented_tensor(transition["action"], self.mirror_action,
                                                                         need_mirrored=False)

        actor, critic = self.agent.actor, self.agent.critic
        done = transition["done"]
        if self.mirrored_actor_memory is not None:
            transition["augmented_actor_memory"] = _concat_memory(transition["actor_memory"], self.mirrored_actor_memory)
        self.mirrored_actor_memory = actor.step_memory(mirrored_observation, self.mirrored_actor_memory, sequential=False)
        actor.reset_memory(self.mirrored_actor_memory, done)
        if self.augments_value:
            if self.mirrored_critic_memory is not None:
                transition["augmented_critic_memory"] = _concat_memory(transition["critic_memory"],
                                                                       self.mirrored_critic_memory)
            self.mirrored_critic_memory = critic.step_memory(mirrored_state, self.mirrored_critic_memory, sequential=False)
            critic.reset_memory(self.mirrored_critic_memory, done)

    def objective(self, metadata, batch):
        augmented_observation = batch["augmented_observation"]
        batch["observation"] = augmented_observation
        batch["next_observation"] = batch.get("augmented_next_observation")
        batch["action"] = batch["augmented_action"]
        if self.agent.has_state:
            batch["state"] = batch["augmented_state"]
            batch["next_state"] = batch.get("augmented_next_state")
        dim = 2 if metadata["temporal"] else 1
        factor = augmented_observation.size(dim)
        repeated = ("action_logp", "advantage") + (("value", "return") if self.augments_value else ())
        for key in repeated:
            original = batch.get(key)
            if original is None:
                if key in ("value", "return"):
                    raise KeyError(key)
                continue
            batch[key] = original.unsqueeze(dim).repeat_interleave(factor, dim=dim)
        if (memory := batch.get("augmented_actor_memory")) is not None:
            batch["actor_memory"] = memory
        if self.augments_value and (memory := batch.get("augmented_critic_memory")) is not None:
            batch["critic_memory"] = memory
        return None

    @classmethod
    def _build_augmented_tensor(cls, original: Tensor, mirror: MirrorFn, augmentation_dim: int = 1,
                                need_mirrored: bool = True) -> tuple[Tensor | None, Tensor]:
        """(mirrored ``[N, V, C]``, augmented ``[N, 1 + V, C]``).  With a MirrorDef on CUDA the augmented tensor is one
        launch and `mirrored` is its ``[:, 1:]`` view."""
        definition = _as_mirror_def(mirror)
        if definition is not None and augmentation_dim == 1 and original.dim() == 2 and _kernel_ok(original, definition):
            augmented = ops.mirror_rows(original, *_identity_plus(definition, original.device), layout="augmented")
            return (augmented[:, 1:] if need_mirrored else None), augmented
        mirrored = cls._build_mirrored(original, mirror).movedim(0, augmentation_dim)
        return mirrored, torch.cat([original.unsqueeze(augmentation_dim), mirrored], dim=augmentation_dim)


def _concat_memory(original, mirrored):
    """concat_memory(unsqueeze(original, 1), mirrored, dim=-2) of nn/utils/recurrent.py on {"hidden", "cell"} dicts, tuples
    or plain tensors."""
    if original is None:
        return None
    if isinstance(original, dict):
        return {k: _concat_memory(original[k], mirrored[k]) for k in original}
    if isinstance(original, (tuple, list)):
        return tuple(_concat_memory(o, m) for o, m in zip(original, mirrored))
    return torch.cat([original.unsqueeze(1), mirrored], dim=-2)


# =================================================================================================
class SymmetricArchitecture(_SymmetryHook):
    """Wraps the actor in a :class:`SymmetricActor` at construction ("On Learning Symmetric Locomotion"); symmetry.py:342-359."""

    def pre_init(self, agent) -> None:
        super().pre_init(agent)
        base = agent.actor_factory
        agent.actor_factory = SymmetricActorFactory(
            base.backbone_factory, base.distribution_factory, base.latent_dim,
            mirror_observation=getattr(agent.environment_spec, "mirror_observation", None),
            mirror_action=getattr(agent.environment_spec, "mirror_action", None))


@dataclass(slots=True)
class SymmetricActorFactory(M.ActorFactory):
    mirror_observation: MirrorFn | None = None
    mirror_action: MirrorFn | None = None

    def __call__(self, input_dim: int | None = None, output_dim: int | None = None) -> "SymmetricActor":
        actor = M.ActorFactory.__call__(self, input_dim, output_dim)
        assert self.mirror_observation is not None, "'mirror_observation' must be defined"
        assert self.mirror_action is not None, "'mirror_action' must be defined"
        return SymmetricActor(actor, self.mirror_observation, self.mirror_action)


class SymmetricActor(M.Actor):
    """``dist(o) = (wrapped(o) + mirror_action(wrapped(mirror_observation(o)))) / 2`` (std: with ``abs`` after mirroring);
    symmetry.py:381-508.  Parameters are the wrapped actor's (registered once, under ``backbone`` / ``distribution``, and
    again under ``wrapped.*`` exactly like the reference's module tree)."""

    def __init__(self, wrapped: M.Actor, mirror_observation: MirrorFn, mirror_action: MirrorFn):
        super().__init__(wrapped.backbone, wrapped.distribution)
        if not isinstance(self.distribution, M.NormalDist):
            raise ValueError("SymmetricActor can only be used with Normal distributions")
        self.wrapped = wrapped
        self.mirror_observation = mirror_observation
        self.mirror_action = mirror_action

    @staticmethod
    def _split(memory):
        return (None, None) if memory is None else (memory["original"], memory["mirrored"])

    def forward(self, observation: Tensor, memory=None, done: Tensor | None = None, **kw):
        original_memory, mirrored_memory = self._split(memory)
        mirrored_observation = self.mirror_observation(observation)
        self.wrapped.intermediate_repr.clear()
        mirrored_dist, mirrored_memory = self.wrapped(mirrored_observation, memory=mirrored_memory, done=done)
        mirrored_repr = dict(self.wrapped.intermediate_repr)
        self.wrapped.intermediate_repr.clear()
        original_dist, original_memory = self.wrapped(observation, memory=original_memory, done=done)
        rep = self.intermediate_repr
        rep["original.action_dist"] = original_dist
        rep.update({f"original.{k}": v for k, v in self.wrapped.intermediate_repr.items()})
        rep["mirrored.observation"] = mirrored_observation
        rep["mirrored.action_dist"] = mirrored_dist
        rep.update({f"mirrored.{k}": v for k, v in mirrored_repr.items()})
        action_dist = {
            "mean": (original_dist["mean"] + self.mirror_action(mirrored_dist["mean"])) / 2,
            "std": (original_dist["std"] + self.mirror_action(mirrored_dist["std"]).abs()) / 2,
        }
        if original_memory is None:
            return action_dist, None
        return action_dist, {"original": original_memory, "mirrored": mirrored_memory}

    def explore(self, observation: Tensor, memory=None, deterministic: bool = False, **kw):
        action_dist, memory = self(observation, memory=memory)
        if deterministic:
            # the deterministic action of a Normal is its mean: the average of the two branches' means IS action_dist.mean
            action = action_dist["mean"]
            logp = self.distribution.compute_logp(action_dist, action)
        else:
            action, logp = self.distribution.sample_from_dist(action_dist)
        return action_dist, (action, logp), memory

    def step_memory(self, observation, memory=None, **kwargs):
        original_memory, mirrored_memory = self._split(memory)
        original_memory = self.wrapped.step_memory(observation, memory=original_memory, **kwargs)
        mirrored_memory = self.wrapped.step_memory(self.mirror_observation(observation), memory=mirrored_memory, **kwargs)
        return None if original_memory is None else {"original": original_memory, "mirrored": mirrored_memory}

    def reset_memory(self, memory, done=None):
        if memory is None:
            return
        self.wrapped.reset_memory(memory["original"], done=done)
        self.wrapped.reset_memory(memory["mirrored"], done=done)
