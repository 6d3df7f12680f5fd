"""Environment specification and the synthetic Anymal-C-rough-shaped environment.

``EnvironmentSpec`` carries the fields of the reference's spec that the hot path reads
(cusrl/template/environment.py:24-260: num_instances, observation/state/action/reward dims, autoreset,
final_state_is_missing).  ``SyntheticEnvironment`` follows the reference's dummy environment
(cusrl/testing/environment.py:39-63: random observations / rewards / done flags) with the flags of the
IsaacLab adapter (cusrl/environment/isaaclab.py:42-45: autoreset=True, final_state_is_missing=True) and the
Isaac-Velocity-Rough-Anymal-C-v0 shapes (obs 235, act 12, reward 1) named in BASELINE.json.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass

import torch

from .runtime import device as resolve_device

__all__ = ["EnvironmentSpec", "SyntheticEnvironment"]


@dataclass
class EnvironmentSpec:
    num_instances: int
    observation_dim: int
    action_dim: int
    state_dim: int | None = None
    reward_dim: int = 1
    autoreset: bool = False
    final_state_is_missing: bool = False
    # symmetry transforms read by the symmetry hooks and by ObservationNormalization (template/environment.py:130-132,156-158)
    mirror_action: Callable[[torch.Tensor], torch.Tensor] | None = None
    mirror_observation: Callable[[torch.Tensor], torch.Tensor] | None = None
    mirror_state: Callable[[torch.Tensor], torch.Tensor] | None = None


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
