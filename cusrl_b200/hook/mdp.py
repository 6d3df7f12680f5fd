"""Observation-side hooks: running mean / std normalisation of observations and states.

Mirrors the reference's ``ObservationNormalization`` / ``ObservationNanToNum`` (cusrl/hook/mdp/observation.py:16-255): same
constructor arguments, mutable ``frozen`` flag, transition keys (``observation`` / ``state`` / ``next_observation`` /
``next_state`` replaced by their normalised values, the raw ones kept under ``original_*``), update rule (statistics of
every step's next observations, of the first observations of a run, and -- for environments that do deliver the final
state -- of the freshly reset rows), ``defer_synchronization`` (the cross-rank merge runs once per update instead of once
per step) and ``renormalize``.  The arithmetic runs on the kernels behind :class:`cusrl_b200.nn.rms.RunningMeanStd`.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np
import torch
from torch import Tensor

from ..nn.rms import RunningMeanStd, mean_var_count
from ..template.hook import Hook

__all__ = ["ObservationNanToNum", "ObservationNormalization"]


class ObservationNanToNum(Hook):
    """Replaces NaN / inf in observations and states in place (observation.py:16-56)."""

    def __init__(self, nan: float = 0.0, posinf: float = 0.0, neginf: float = 0.0):
        super().__init__()
        self.nan, self.posinf, self.neginf = nan, posinf, neginf

    def nan_to_num_(self, tensor: Tensor | None) -> None:
        if tensor is not None:
            tensor.nan_to_num_(nan=self.nan, posinf=self.posinf, neginf=self.neginf)

    def pre_act(self, transition) -> None:
        self.nan_to_num_(transition.get("observation"))
        self.nan_to_num_(transition.get("state"))

    def post_step(self, transition) -> None:
        self.nan_to_num_(transition.get("next_observation"))
        self.nan_to_num_(transition.get("next_state"))


class ObservationNormalization(Hook):
    def __init__(self, max_count: int | None = None, defer_synchronization: bool = False, renormalize: bool = False):
        if max_count is not None and max_count <= 0:
            raise ValueError("'max_count' must be positive or None")
        super().__init__()
        self.max_count = max_count
        self.defer_synchronization = defer_synchronization
        self.renormalize = renormalize
        self.frozen: bool = False
        self.register_mutable("frozen")
        self.observation_rms: RunningMeanStd
        self.state_rms: RunningMeanStd | None
        self._mirror_observation = None
        self._mirror_state = None
        self._observation_is_subset_of_state = None
        self._last_done: Tensor | None = None

    def freeze(self):
        self.frozen = True
        return self

    def init(self) -> None:
        spec = self.agent.environment_spec
        observation_dim = self.agent.observation_dim
        self._mirror_observation = getattr(spec, "mirror_observation", None)
        self._mirror_state = getattr(spec, "mirror_state", None)
        subset = getattr(spec, "observation_is_subset_of_state", None)
        if subset is not None:
            if not self.agent.has_state:
                raise ValueError("'observation_is_subset_of_state' is set without defining the state")
            if isinstance(subset, (np.ndarray, Sequence)):
                subset = self.agent.to_tensor(np.asarray(subset))
            self._observation_is_subset_of_state = subset
            self.register_module("observation_rms", RunningMeanStd(observation_dim))
        else:
            self.register_module("observation_rms", RunningMeanStd(
                observation_dim, max_count=self.max_count, groups=getattr(spec, "observation_stat_groups", ()),
                excluded_indices=getattr(spec, "observation_normalization_excluded_indices", None)))
        if self.agent.has_state:
            self.register_module("state_rms", RunningMeanStd(
                self.agent.state_dim, max_count=self.max_count, groups=getattr(spec, "state_stat_groups", ()),
                excluded_indices=getattr(spec, "state_normalization_excluded_indices", None)))
        else:
            self.state_rms = None

    def pre_export(self, graph) -> None:
        """The deployed policy normalises its observation with the frozen running statistics (observation.py:248-255).  The
        node is a plain-torch twin of ``observation_rms`` (same mean / std / clamp; the B200 module runs kernels and cannot
        be traced), also when the hook is inactive -- like the reference, export visits every hook."""
        from ..template.export import _Affine

        rms = self.observation_rms
        graph.add_node(_Affine(rms.mean, rms.std, getattr(rms, "clamp", None)), module_name="observation_rms",
                       input_names={"input": "observation"}, output_names="observation", expose_outputs=False)

    def pre_act(self, transition) -> None:
        observation, state = transition["observation"], transition.get("state")
        if self._last_done is None or not self.agent.environment_spec.final_state_is_missing:
            self._update_rms(observation, state, self._last_done)
        transition["original_observation"] = observation
        transition["observation"] = self.observation_rms.normalize(observation)
        if self.state_rms is not None:
            transition["original_state"] = state
            transition["state"] = self.state_rms.normalize(state)

    def post_step(self, transition) -> None:
        next_observation, next_state = transition["next_observation"], transition.get("next_state")
        self._update_rms(next_observation, next_state)
        self._last_done = transition["done"].squeeze(-1)
        transition["original_next_observation"] = next_observation
        transition["next_observation"] = self.observation_rms.normalize(next_observation)
        if self.state_rms is not None:
            transition["original_next_state"] = next_state
            transition["next_state"] = self.state_rms.normalize(next_state)

    def _update_rms(self, observation: Tensor, state: Tensor | None, indices: Tensor | None = None) -> None:
        if self.agent.inference_mode or self.frozen:
            return
        if state is not None:
            self._update_rms_impl(state, self.state_rms, self._mirror_state, indices)
        if self._observation_is_subset_of_state is not None:
            self._copy_observation_stats_from_state()
        else:
            self._update_rms_impl(observation, self.observation_rms, self._mirror_observation, indices)

    def _update_rms_impl(self, observation: Tensor, rms: RunningMeanStd, mirror=None, indices: Tensor | None = None) -> None:
        if indices is not None:
            observation = observation[indices]   # data-dependent row count: one host sync, as in the reference
        mean, var, count = mean_var_count(observation)
        if mirror is not None:
            mirrored_mean = mirror(mean)
            mirrored_var = abs(mirror(var))
            var = (var + mirrored_var) / 2 + (mean - mirrored_mean) ** 2 / 4
            mean = (mean + mirrored_mean) / 2
        rms.update_from_stats(mean, var, count, synchronize=not self.defer_synchronization)

    def _copy_observation_stats_from_state(self) -> None:
        idx = self._observation_is_subset_of_state
        self.observation_rms.mean.copy_(self.state_rms.mean[idx])
        self.observation_rms.var.copy_(self.state_rms.var[idx])
        self.observation_rms.std.copy_(self.state_rms.std[idx])
        self.observation_rms.count = self.state_rms.count

    def pre_update(self, buffer) -> None:
        if self.defer_synchronization:
            if self.state_rms is not None:
                self.state_rms.synchronize()
            if self._observation_is_subset_of_state is not None:
                self._copy_observation_stats_from_state()
            else:
                self.observation_rms.synchronize()

    def objective(self, metadata, batch):
        if self.renormalize:
            batch["observation"] = self.observation_rms.normalize(batch["original_observation"])
            batch["next_observation"] = self.observation_rms.normalize(batch["original_next_observation"])
            if self.state_rms is not None:
                batch["state"] = self.state_rms.normalize(batch["original_state"])
                batch["next_state"] = self.state_rms.normalize(batch["original_next_state"])
        return None
