"""ObservationNormalization / RunningMeanStd against the live reference's golden `obsnorm.npz`
(tests/golden/make_golden.py::make_obsnorm): the normalised observations handed to the agent and the running statistics after
every environment step, for an IsaacLab-style environment (final state missing, no state) and for one that delivers final
states and has a critic state (masked update on the freshly reset rows).  Runs on the CPU (host logic, torch arithmetic)
and, marked gpu, through the kernels of csrc/rms_kernels.cu."""

from __future__ import annotations

from types import SimpleNamespace

import pytest
import torch


def _run(golden, device, tag):
    import cusrl_b200 as C

    g = golden("obsnorm")
    final_missing, state_dim = (True, None) if tag == "a" else (False, 7)
    hook = C.ObservationNormalization()
    spec = SimpleNamespace(final_state_is_missing=final_missing)
    agent = SimpleNamespace(environment_spec=spec, observation_dim=19, state_dim=state_dim, has_state=state_dim is not None,
                            inference_mode=False, setup_module=lambda m: m.to(device), to_tensor=torch.as_tensor)
    hook.pre_init(agent)
    hook.init()
    for t in range(5):
        tr = {"observation": g.t(f"{tag}_obs_in_{t}", device)}
        if state_dim:
            tr["state"] = g.t(f"{tag}_state_in_{t}", device)
        hook.pre_act(tr)
        assert torch.equal(tr["original_observation"].cpu(), g.t(f"{tag}_obs_in_{t}"))
        assert torch.allclose(tr["observation"].cpu(), g.t(f"{tag}_obs_norm_{t}"), rtol=1e-5, atol=1e-5), t
        if state_dim:
            assert torch.allclose(tr["state"].cpu(), g.t(f"{tag}_state_norm_{t}"), rtol=1e-5, atol=1e-5), t
        tr2 = {"next_observation": g.t(f"{tag}_next_obs_in_{t}", device), "done": g.t(f"{tag}_done_{t}", device)}
        if state_dim:
            tr2["next_state"] = g.t(f"{tag}_next_state_in_{t}", device)
        hook.post_step(tr2)
        assert torch.allclose(tr2["next_observation"].cpu(), g.t(f"{tag}_next_obs_norm_{t}"), rtol=1e-5, atol=1e-5), t
        assert torch.allclose(hook.observation_rms.mean.cpu(), g.t(f"{tag}_mean_{t}"), rtol=1e-5, atol=1e-6), t
        assert torch.allclose(hook.observation_rms.var.cpu(), g.t(f"{tag}_var_{t}"), rtol=1e-5, atol=1e-6), t
        assert hook.observation_rms.count == int(g.np(f"{tag}_count_{t}"))
    if state_dim:
        assert torch.allclose(hook.state_rms.mean.cpu(), g.t(f"{tag}_state_mean"), rtol=1e-5, atol=1e-6)
        assert torch.allclose(hook.state_rms.var.cpu(), g.t(f"{tag}_state_var"), rtol=1e-5, atol=1e-6)
        assert hook.state_rms.count == int(g.np(f"{tag}_state_count"))
    # state dict layout of the reference (mean / var / std buffers + count as extra state) round-trips
    sd = hook.state_dict()
    assert set(sd["observation_rms"]) == {"mean", "var", "std", "_extra_state"}
    fresh = C.ObservationNormalization()
    fresh.pre_init(agent)
    fresh.init()
    fresh.load_state_dict(sd)
    assert fresh.observation_rms.count == hook.observation_rms.count
    assert torch.equal(fresh.observation_rms.mean, hook.observation_rms.mean)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_observation_normalization_matches_reference_cpu(golden, tag):
    _run(golden, "cpu", tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_observation_normalization_matches_reference_gpu(golden, tag):
    from cusrl_b200 import build

    build.build()
    _run(golden, "cuda", tag)


@pytest.mark.gpu
def test_rms_kernels_at_rollout_size():
    """Column statistics / merge / normalise at 65536 x 235 against fp64 torch (the north-star 1e-5 bound), pitched rows."""
    from cusrl_b200 import build

    build.build()
    from cusrl_b200 import ops
    from cusrl_b200.nn import RunningMeanStd

    g = torch.Generator(device="cuda").manual_seed(0)
    back = torch.randn(65536, 236, device="cuda", generator=g) * 2.5 + 0.7
    x = back[:, :235]
    mv = ops.column_stats(x)
    var, mean = torch.var_mean(x.double(), dim=0, correction=0)
    assert torch.allclose(mv[:235].double(), mean, rtol=1e-5, atol=1e-6) and torch.allclose(mv[235:].double(), var, rtol=1e-5, atol=1e-6)
    rms = RunningMeanStd(235).cuda()
    rms.update(x)
    rms.update(x * 0.5 - 1.0)
    both = torch.cat([x.double(), (x * 0.5 - 1.0).double()])
    var2, mean2 = torch.var_mean(both, dim=0, correction=0)
    assert rms.count == 131072
    assert torch.allclose(rms.mean.double(), mean2, rtol=1e-5, atol=1e-5) and torch.allclose(rms.var.double(), var2, rtol=1e-5, atol=1e-5)
    out = torch.full((65536, 236), float("nan"), device="cuda")
    rms.normalize(x, out=out[:, :235], zero_padding=True)
    ref = ((x.double() - rms.mean.double()) / rms.std.double()).clamp(-10, 10)
    assert torch.allclose(out[:, :235].double(), ref, rtol=1e-5, atol=1e-5) and float(out[:, 235].abs().max()) == 0.0


@pytest.mark.gpu
def test_ppo_with_observation_normalization_trains():
    from cusrl_b200 import build

    build.build()
    import numpy as np

    import cusrl_b200 as C

    env = C.SyntheticEnvironment(512, device="cuda", seed=3)
    agent = C.anymal_c_rough_ppo(device="cuda", normalize_observation=True).from_environment(env)
    assert [h.name for h in agent.hook][:3] == ["module_initialization", "observation_normalization", "value_computation"]
    history = C.Trainer(env, agent, num_iterations=2).run_training_loop()
    for key in ("Agent/value_loss", "Agent/surrogate_loss", "Agent/kl_divergence"):
        assert np.isfinite(history[-1][key]), key
    rms = agent.hook["observation_normalization"].observation_rms
    assert rms.count == 512 * (1 + 2 * 24) and float(rms.std.min()) > 0.5
    assert "original_observation" in agent.buffer and "original_next_observation" in agent.buffer
