"""CPU, build container only: the plugin registers with the UNMODIFIED reference (mounted at /root/reference) and the
reference's own Trainer accepts the B200 agent factory.  Skipped wherever the reference checkout is absent (GPU box)."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

REFERENCE = Path(os.environ.get("CUSRL_REFERENCE", "/root/reference"))
pytestmark = pytest.mark.skipif(not (REFERENCE / "cusrl").is_dir(), reason="reference checkout not available")


@pytest.fixture(scope="module")
def reference():
    shims = Path(__file__).resolve().parent / "golden" / "_shims"
    added = [str(shims), str(REFERENCE)]
    sys.path[:0] = added
    try:
        import cusrl

        yield cusrl
    finally:
        for p in added:
            sys.path.remove(p)


def test_plugin_registers_and_reference_trainer_accepts_the_agent(reference):
    import cusrl_b200.plugin as plugin
    from cusrl.zoo import get_experiment

    spec = get_experiment("Synthetic-AnymalC-Rough-v0", plugin.ALGORITHM_NAME)
    factory = spec.to_training_factory()
    assert type(factory.agent_factory).__name__ == "PpoAgentFactory"
    assert factory.agent_factory.actor_hidden_dims == (512, 256, 128) and factory.agent_factory.lr == 1e-3
    factory.agent_factory.device = "cpu"   # construction only: the kernels themselves need a GPU
    env = plugin.SyntheticAnymalEnvironment(num_envs=8)
    trainer = reference.Trainer(env, factory.agent_factory, logger_factory=None, num_iterations=1, verbose=False)
    import cusrl_b200

    assert isinstance(trainer.agent, cusrl_b200.ActorCritic)
    assert trainer.agent.parallelism == 8 and trainer.agent.observation_dim == 235
    assert [h.name for h in trainer.agent.hook][:3] == ["module_initialization", "value_computation",
                                                        "generalized_advantage_estimation"]
