"""Drop-in registration with the UNMODIFIED reference CLI:

    python -m cusrl train -env Isaac-Velocity-Rough-Anymal-C-v0 -alg ppo-b200 -m cusrl_b200.plugin
    python -m cusrl train -env Synthetic-AnymalC-Rough-v0 -alg ppo-b200 -m cusrl_b200.plugin -- --num-iterations 10
    torchrun --nproc-per-node 8 -m cusrl train -env ... -alg ppo-b200 -m cusrl_b200.plugin

(``-m`` takes the REST of the command line up to ``--`` as the module's own arguments, cusrl/cli/train.py:29-30, so it
goes last.)  Exercised end to end on a B200 by tests/test_cli_gpu.py against the unmodified reference in baseline/_ref.

``-m <module>`` makes the reference import this module before the experiment lookup (cusrl/cli/train.py:29-32,45);
importing it calls the reference's own ``cusrl.zoo.register_experiment`` (cusrl/zoo/registry.py:27-82) with
:class:`cusrl_b200.PpoAgentFactory` as the agent factory.  The reference ``Trainer`` only needs
``agent_factory.from_environment(env)`` and an agent with ``act / step / update / set_iteration / device /
state_dict / load_state_dict`` (cusrl/template/trainer.py:258,270,276,299-318,330,346-348), which
:class:`cusrl_b200.ActorCritic` provides, so the whole hot path runs on the B200 kernels while CLI, zoo, loggers,
checkpoints and environments stay the reference's.

This module needs the reference package (``import cusrl``) and is not imported by ``cusrl_b200`` itself.
"""

from __future__ import annotations

import json
import os

import torch

import cusrl  # the reference package
from cusrl.zoo import register_experiment

import cusrl_b200
from cusrl_b200.preset import anymal_c_rough_ppo

__all__ = ["ALGORITHM_NAME", "RECURRENT_ALGORITHM_NAME", "SyntheticAnymalEnvironment", "make_synthetic_env"]

ALGORITHM_NAME = "ppo-b200"  # experiment names may not contain ':', '_', '/' or '\\' (cusrl/zoo/experiment.py:218-223)
RECURRENT_ALGORITHM_NAME = "ppo-b200-lstm"  # the recurrent preset (LSTM 2 x 256 actor and critic, preset/ppo.py:185-298)


class SyntheticAnymalEnvironment(cusrl.template.Environment):
    """Anymal-C-rough-shaped random environment on the agent's device, with the IsaacLab adapter's flags
    (cusrl/environment/isaaclab.py:42-45); modelled on cusrl/testing/environment.py:39-63."""

    def __init__(self, num_envs: int = 4096, observation_dim: int = 235, action_dim: int = 12, p_term: float = 0.01,
                 p_trunc: float = 0.001, **kwargs):
        device = cusrl_b200.device()
        super().__init__(observation_dim, action_dim, num_instances=num_envs, device=device, autoreset=True,
                         final_state_is_missing=True, **kwargs)
        self._device, self._p_term, self._p_trunc = device, p_term, p_trunc

    def reset(self, *, indices=None, randomize_episode_progress=False):
        n = self.num_instances if indices is None else torch.zeros(self.num_instances)[indices].numel()
        return torch.randn(n, self.observation_dim, device=self._device), None, {}

    def step(self, action):
        n, dev = self.num_instances, self._device
        return (torch.randn(n, self.observation_dim, device=dev), None, torch.randn(n, self.spec.reward_dim, device=dev),
                torch.rand(n, 1, device=dev) < self._p_term, torch.rand(n, 1, device=dev) < self._p_trunc, {})


def make_synthetic_env(id: str = "Synthetic-AnymalC-Rough-v0", argv=None, **kwargs):
    return SyntheticAnymalEnvironment(**kwargs)


class MetricsJsonl(cusrl.template.trainer.TrainerHook):
    """Appends the trainer's per-iteration log dict (``Agent/*``, ``Perf/agent_fps``, ...; cusrl/template/trainer.py:374-397)
    to the file named by ``CUSRL_B200_METRICS_JSONL``, one JSON object per line, on the main process."""

    def __init__(self, path: str):
        self.path = path

    def pre_log_info(self, info):
        if cusrl.utils.distributed.is_main_process():
            with open(self.path, "a") as f:
                f.write(json.dumps({k: float(v) for k, v in info.items()}) + "\n")


def _register() -> None:
    hooks = [MetricsJsonl(path)] if (path := os.environ.get("CUSRL_B200_METRICS_JSONL")) else []
    common = dict(algorithm_name=ALGORITHM_NAME, agent_meta_factory=anymal_c_rough_ppo, num_iterations=1500,
                  checkpoint_interval=100, trainer_hooks=hooks)
    register_experiment(environment_name="Synthetic-AnymalC-Rough-v0", training_env_factory=make_synthetic_env, **common)
    recurrent = dict(common, algorithm_name=RECURRENT_ALGORITHM_NAME, agent_meta_factory=cusrl_b200.RecurrentPpoAgentFactory)
    register_experiment(environment_name="Synthetic-AnymalC-Rough-v0", training_env_factory=make_synthetic_env, **recurrent)
    try:  # the real simulator adapter, when IsaacLab is installed (same env list as cusrl/zoo/isaaclab/locomotion.py:40-47)
        from cusrl.environment import make_isaaclab_env
    except Exception:  # pragma: no cover - IsaacLab absent
        return
    register_experiment(
        environment_name=[f"Isaac-Velocity-Rough-{robot}-v0" for robot in
                          ("Anymal-B", "Anymal-C", "Anymal-D", "Unitree-A1", "Unitree-Go1", "Unitree-Go2")],
        training_env_factory=make_isaaclab_env, playing_env_factory_kwargs={"play": True}, **common)
    register_experiment(
        environment_name=[f"Isaac-Velocity-Rough-{robot}-v0" for robot in
                          ("Anymal-B", "Anymal-C", "Anymal-D", "Unitree-A1", "Unitree-Go1", "Unitree-Go2")],
        training_env_factory=make_isaaclab_env, playing_env_factory_kwargs={"play": True}, **recurrent)


_register()
