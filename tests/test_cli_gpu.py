"""GPU: the drop-in boundary end to end -- the UNMODIFIED reference CLI (baseline/_ref, see tools/install_reference.py)
trains with the B200 agent through ``python -m cusrl train ... -m cusrl_b200.plugin``: the reference's own argument
parsing, experiment registry, tyro overrides, Trainer loop, timers and logger around this repository's hot path."""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


def _reference_pythonpath() -> str:
    from install_reference import reference_path

    return os.pathsep.join(reference_path() + [str(ROOT)])


def _run_cli(tmp_path, extra, nproc: int = 1, algorithm: str = "ppo-b200"):
    out = tmp_path / "metrics.jsonl"
    env = dict(os.environ, PYTHONPATH=_reference_pythonpath(), CUSRL_B200_METRICS_JSONL=str(out))
    for key in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE"):
        env.pop(key, None)
    cmd = [sys.executable, "-m", "cusrl", "train", "-env", "Synthetic-AnymalC-Rough-v0", "-alg", algorithm, "--seed", "1",
           "-m", "cusrl_b200.plugin", "--", "--num-iterations", "3", "--log-dir", str(tmp_path / "logs"), *extra]
    res = subprocess.run(cmd, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return [json.loads(line) for line in out.read_text().splitlines()], res


def test_reference_cli_trains_with_the_b200_plugin(tmp_path):
    try:
        _reference_pythonpath()
    except RuntimeError as error:
        pytest.skip(str(error))
    rows, res = _run_cli(tmp_path, ["--env-kwargs", '{"num_envs": 2048}'])
    assert len(rows) == 3
    last = rows[-1]
    for key in ("Agent/value_loss", "Agent/surrogate_loss", "Agent/entropy_loss", "Agent/kl_divergence", "Agent/ratio",
                "Agent/entropy", "Agent/value", "Agent/grad_norm/default", "Agent/action_std", "Agent/lr_scale",
                "Agent/importance_weighted_advantage", "Perf/agent_fps", "Perf/agent_time", "Perf/environment_fps",
                "Metric/episode_length"):
        assert key in last and last[key] == last[key], key
    assert last["Perf/agent_fps"] > 0
    # the reference's loggers / checkpoint layout were used untouched: logs/<experiment>/<timestamp>/ckpt + info
    runs = list((tmp_path / "logs").glob("*/*"))
    assert runs, res.stdout[-2000:]
    assert any((run / "ckpt").is_dir() for run in runs if run.is_dir())


def test_reference_cli_trains_the_recurrent_preset(tmp_path):
    """`-alg ppo-b200-lstm`: the recurrent preset (LSTM 2 x 256) under the reference's Trainer -- sequence-resident LSTM
    kernels, fused recurrent rollout step and temporal minibatches behind the unmodified CLI."""
    try:
        _reference_pythonpath()
    except RuntimeError as error:
        pytest.skip(str(error))
    rows, res = _run_cli(tmp_path, ["--env-kwargs", '{"num_envs": 512}'], algorithm="ppo-b200-lstm")
    assert len(rows) == 3
    last = rows[-1]
    for key in ("Agent/value_loss", "Agent/surrogate_loss", "Agent/entropy_loss", "Agent/kl_divergence", "Agent/entropy",
                "Agent/grad_norm/default", "Perf/agent_fps"):
        assert key in last and last[key] == last[key], key
    assert last["Perf/agent_fps"] > 0
