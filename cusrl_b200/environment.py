"""Environment specification and the synthetic Anymal-C-rough-shaped environment.

``EnvironmentSpec`` carries the fields of the reference's spec that the hot path reads
(cusrl/template/environment.py:24-260: num_instances, observation/state/action/reward dims, autoreset,
final_state_is_missing).  ``SyntheticEnvironment`` follows the reference's dummy environment
(cusrl/testing/environment.py:39-63: random observations / rewards / done flags) with the flags of the
IsaacLab adapter (cusrl/environment/isaaclab.py:42-45: autoreset=True, final_state_is_missing=True) and the
Isaac-Velocity-Rough-Anymal-C-v0 shapes (obs 235, act 12, reward 1) named in BASELINE.json.
"""

from __future__ import annotations

import torch

from .runtime import device as resolve_device

__all__ = ["EnvironmentSpec", "SyntheticEnvironment"]


class EnvironmentSpec:
    """The reference's specification object (cusrl/template/environment.py:24-176): same attribute names and defaults, extra
    keyword arguments become attributes, ``get(key, default)``.  The hot path reads ``num_instances``, the four dimensions,
    ``autoreset`` and ``final_state_is_missing``; the symmetry hooks read ``mirror_*``; ``ObservationNormalization`` reads the
    ``*_stat_groups`` / ``*_excluded_indices`` / ``observation_is_subset_of_state`` fields; ``agent.export`` reads
    ``observation_normalization`` / ``action_denormalization``.

    Positional arguments: the reference's order ``EnvironmentSpec(observation_dim, action_dim, *, num_instances=1, ...)`` when
    exactly two are given, and this package's original order ``EnvironmentSpec(num_instances, observation_dim, action_dim,
    state_dim=None, reward_dim=1, ...)`` otherwise (every call site of the tests and tools)."""

    _POSITIONAL = ("num_instances", "observation_dim", "action_dim", "state_dim", "reward_dim", "autoreset",
                   "final_state_is_missing", "mirror_action", "mirror_observation", "mirror_state")
    _DEFAULTS: dict = dict(
        num_instances=1, state_dim=None, reward_dim=1, action_denormalization=None, action_space=None, autoreset=False,
        demonstration_sampler=None, device="cpu", environment_instance=None, final_state_is_missing=False, mirror_action=None,
        mirror_observation=None, mirror_state=None, observation_is_subset_of_state=None, observation_stat_groups=(),
        observation_normalization=None, observation_normalization_excluded_indices=None, observation_space=None,
        state_stat_groups=(), state_normalization=None, state_normalization_excluded_indices=None, timestep=None)

    def __init__(self, *args, **kwargs):
        if len(args) == 2 and "action_dim" not in kwargs:
            names = ("observation_dim", "action_dim")          # the reference's positional order
        else:
            names = self._POSITIONAL
        if len(args) > len(names):
            raise TypeError(f"EnvironmentSpec takes at most {len(names)} positional arguments ({len(args)} given)")
        for name, value in zip(names, args):
            if name in kwargs:
                raise TypeError(f"EnvironmentSpec got multiple values for argument '{name}'")
            kwargs[name] = value
        for required in ("observation_dim", "action_dim"):
            if required not in kwargs:
                raise TypeError(f"EnvironmentSpec missing required argument '{required}'")
        values = {**self._DEFAULTS, **kwargs}
        values["device"] = torch.device(values["device"])
        values["observation_stat_groups"] = tuple(values["observation_stat_groups"])
        values["state_stat_groups"] = tuple(values["state_stat_groups"])
        for key, value in values.items():     # unknown keywords become attributes, like the reference's **kwargs
            setattr(self, key, value)

    def get(self, key: str, default=None):
        return self.__dict__.get(key, default)

    def __repr__(self) -> str:
        shown = {k: v for k, v in self.__dict__.items() if k in ("num_instances", "observation_dim", "action_dim", "state_dim",
                                                                  "reward_dim", "autoreset", "final_state_is_missing")}
        return "EnvironmentSpec(" + ", ".join(f"{k}={v!r}" for k, v in shown.items()) + ")"


class SyntheticEnvironment:
    """Device-resident random rollouts: obs ~ N(0,1), reward ~ N(0,1), terminated ~ B(p_term), truncated ~ B(p_trunc)."""

    def __init__(self, num_instances: int, observation_dim: int = 235, action_dim: int = 12, reward_dim: int = 1,
                 p_term: float = 0.01, p_trunc: float = 0.001, device=None, seed: int | None = None):
        self.device = resolve_device(device)
        self.num_instances = num_instances
        self.spec = EnvironmentSpec(num_instances, observation_dim, action_dim, None, reward_dim,
                                    autoreset=True, final_state_is_missing=True)
        self.p_term, self.p_trunc = p_term, p_trunc
        self.generator = torch.Generator(device=self.device)
        if seed is not None:
            self.generator.manual_seed(seed)

    def reset(self):
        s = self.spec
        return torch.randn(s.num_instances, s.observation_dim, device=self.device, generator=self.generator), None, {}

    def step(self, action: torch.Tensor):
        s, g, dev = self.spec, self.generator, self.device
        obs = torch.randn(s.num_instances, s.observation_dim, device=dev, generator=g)
        reward = torch.randn(s.num_instances, s.reward_dim, device=dev, generator=g)
        terminated = torch.rand(s.num_instances, 1, device=dev, generator=g) < self.p_term
        truncated = torch.rand(s.num_instances, 1, device=dev, generator=g) < self.p_trunc
        return obs, None, reward, terminated, truncated, {}
