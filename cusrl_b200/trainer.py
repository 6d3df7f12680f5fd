"""Rollout/update loop with the reference's timing sections.

Mirrors ``Trainer._rollout_and_update`` (cusrl/template/trainer.py:296-321) and its definition of
``Perf/agent_fps`` = steps * envs * world_size / time inside the "agent" timer (trainer.py:385-393), with
CUDA-event timing like the reference's ``Timer`` (cusrl/utils/timing.py:49-94).  The per-step
``get_done_indices(...).tolist()`` host sync of the reference (trainer.py:306) is not needed for
``autoreset=True`` environments and is omitted."""

from __future__ import annotations

import time

import torch

from . import distributed

__all__ = ["Trainer", "SectionTimer"]


class SectionTimer:
    """Accumulates time per named section; CUDA events on GPU, perf_counter on CPU."""

    def __init__(self, device: torch.device):
        self.cuda = device.type == "cuda"
        self._pending: dict[str, list] = {}
        self._cpu: dict[str, float] = {}

    class _Section:
        def __init__(self, timer, name):
            self.timer, self.name = timer, name

        def __enter__(self):
            if self.timer.cuda:
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record()
            else:
                self.start = time.perf_counter()

        def __exit__(self, *exc):
            if self.timer.cuda:
                end = torch.cuda.Event(enable_timing=True)
                end.record()
                self.timer._pending.setdefault(self.name, []).append((self.start, end))
            else:
                self.timer._cpu[self.name] = self.timer._cpu.get(self.name, 0.0) + time.perf_counter() - self.start

    def record(self, name: str):
        return SectionTimer._Section(self, name)

    def __getitem__(self, name: str) -> float:
        if self.cuda:
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in self._pending.get(name, [])) * 1e-3
        return self._cpu.get(name, 0.0)

    def clear(self) -> None:
        self._pending.clear()
        self._cpu.clear()


class Trainer:
    def __init__(self, environment, agent, num_iterations: int = 1, verbose: bool = False):
        self.environment, self.agent, self.num_iterations, self.verbose = environment, agent, num_iterations, verbose
        if not getattr(getattr(environment, "spec", None), "autoreset", True):
            # the reference resets finished instances itself when the environment does not (trainer.py:306-312, one host
            # sync per step); this loop omits that branch, so it must not silently train on dead episodes
            raise ValueError("cusrl_b200.Trainer drives autoreset environments only (spec.autoreset must be True); use the "
                             "reference Trainer (python -m cusrl train -m cusrl_b200.plugin) for the others")
        self.timer = SectionTimer(agent.device)
        self.iteration = 0
        self.history: list[dict[str, float]] = []

    def run_training_loop(self):
        with self.timer.record("environment"):
            observation, state, _ = self.environment.reset()
        while self.iteration < self.num_iterations:
            observation, state = self._rollout_and_update(observation, state)
            self.iteration += 1
        return self.history

    def _rollout_and_update(self, observation, state):
        steps = 0
        while True:
            with self.timer.record("agent"):
                action = self.agent.act(observation, state)
            with self.timer.record("environment"):
                next_observation, next_state, reward, terminated, truncated, info = self.environment.step(action)
            with self.timer.record("agent"):
                ready = self.agent.step(next_observation, reward, terminated, truncated, next_state, **info)
            observation, state = next_observation, next_state
            steps += 1
            if ready:
                break
        with self.timer.record("agent"):
            info = self.agent.update()
        info["Perf/agent_time"] = self.timer["agent"]
        info["Perf/environment_time"] = self.timer["environment"]
        info = distributed.average_dict(info)
        num_steps = steps * self.environment.num_instances * distributed.world_size()
        info["Perf/agent_fps"] = num_steps / info["Perf/agent_time"]
        info["Perf/environment_fps"] = num_steps / max(info["Perf/environment_time"], 1e-12)
        self.history.append(info)
        if self.verbose and distributed.is_main_process():
            print(f"iteration {self.iteration + 1}: agent_fps {info['Perf/agent_fps']:.0f}")
        self.timer.clear()
        return observation, state
