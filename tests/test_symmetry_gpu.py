"""GPU: the symmetry transform kernel (bit-exact against the reference formula in every layout) and one whole PPO
iteration per symmetry hook against the live-reference golden ``symmetry.npz`` (tests/golden/make_golden.py::make_symmetry:
the reference's own SymmetricDataAugmentation / MirrorSymmetryLoss / TransitionMirroring / SymmetricArchitecture hooks in
its PPO preset on CPU; inputs regenerated from seeds on both sides)."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import recipes as R  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def C():
    from cusrl_b200 import build

    build.build()
    import cusrl_b200

    return cusrl_b200


@pytest.mark.parametrize("width,rows", [(235, 4096), (12, 1000), (19, 33), (1, 5), (48, 65536)])
def test_mirror_kernel_bit_exact_in_every_layout(C, width, rows):
    from cusrl_b200 import ops
    from cusrl_b200.hook.symmetry import MirrorDef, SymmetricDataAugmentation, _identity_plus

    mirror = MirrorDef(*R.mirror_tables(width, seed=width))
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, width, generator=g)
    x[0, 0] = 0.0   # a flipped zero must come out as -0.0, like torch's multiply
    ref = x[..., mirror.destination] * mirror.multiplier
    xd = x.to(DEV)
    out = mirror(xd)
    assert torch.equal(out.cpu(), ref) and torch.equal(torch.signbit(out.cpu()), torch.signbit(ref))
    # leading dims and pitched (padded-row) inputs
    x3 = xd.reshape(1, rows, width)
    assert torch.equal(mirror(x3).cpu(), ref.reshape(1, rows, width))
    padded = torch.zeros(rows, width + 3, device=DEV)
    padded[:, :width] = xd
    assert torch.equal(mirror(padded[:, :width]).cpu(), ref)
    # stacked [V, N, C] with several transforms
    other = MirrorDef(*R.mirror_tables(width, seed=width + 100))
    dest = torch.cat([mirror.tables(xd.device)[0], other.tables(xd.device)[0]])
    mult = torch.cat([mirror.tables(xd.device)[1], other.tables(xd.device)[1]])
    stacked = ops.mirror_rows(xd, dest, mult, layout="stacked")
    assert torch.equal(stacked[0].cpu(), ref)
    assert torch.equal(stacked[1].cpu(), x[..., other.destination] * other.multiplier)
    # augmented [N, 1 + V, C] = cat([x.unsqueeze(1), mirrored.movedim(0, 1)], 1), also straight into padded rows
    mirrored, augmented = SymmetricDataAugmentation._build_augmented_tensor(xd, mirror)
    expect = torch.cat([x.unsqueeze(1), ref.unsqueeze(1)], dim=1)
    assert torch.equal(augmented.cpu(), expect) and torch.equal(mirrored.cpu(), ref.unsqueeze(1))
    pad_w = (width + 3) // 4 * 4 + 4
    backing = torch.full((rows, 2, pad_w), 7.0, device=DEV)
    ops.mirror_rows(xd, *_identity_plus(mirror, xd.device), layout="augmented", out=backing[..., :width])
    assert torch.equal(backing[..., :width].cpu(), expect)
    assert torch.count_nonzero(backing[..., width:]).item() == 0
    # with autograd the transform stays differentiable (torch indexing) and agrees
    xg = xd.clone().requires_grad_(True)
    yg = mirror(xg)
    assert yg.requires_grad and torch.equal(yg.detach().cpu(), ref)
    gout = torch.randn(rows, width, generator=g).to(DEV)
    (yg * gout).sum().backward()
    xr = x.clone().requires_grad_(True)
    ((xr[..., mirror.destination] * mirror.multiplier) * gout.cpu()).sum().backward()
    assert torch.equal(xg.grad.cpu(), xr.grad)   # the adjoint is an index permutation with sign flips too: exact


def _make_agent(C, variant: str):
    sh = R.SYMMETRY_SHAPE
    factory = C.anymal_c_rough_ppo(num_steps_per_update=sh["T"], actor_hidden_dims=sh["hidden"],
                                   critic_hidden_dims=sh["hidden"], sampler_epochs=3, sampler_mini_batches=2,
                                   device=DEV).to_underlying()
    if variant == "augmentation":
        factory.register_hook(C.SymmetricDataAugmentation(), before="value_loss")
    elif variant == "mirror_loss":
        factory.register_hook(C.MirrorSymmetryLoss(0.5, symmetrize_action_std=True), after="ppo_surrogate_loss")
    elif variant == "transition_mirroring":
        factory.register_hook(C.TransitionMirroring(), index=0)
    else:
        factory.register_hook(C.SymmetricArchitecture(), after="module_initialization")
    spec = C.EnvironmentSpec(sh["N"], sh["obs"], sh["act"], autoreset=True, final_state_is_missing=True,
                             mirror_observation=C.MirrorDef(*R.mirror_tables(sh["obs"], seed=21)),
                             mirror_action=C.MirrorDef(*R.mirror_tables(sh["act"], seed=22)))
    agent = factory(spec)
    R.set_seeded_parameters(agent.named_parameters(), seed=23)
    from cusrl_b200 import ops

    ops.invalidate_weight_cache()
    return agent


@pytest.mark.parametrize("cuda_graphs", [False, True])
@pytest.mark.parametrize("variant", R.SYMMETRY_VARIANTS)
def test_symmetry_hook_iteration_matches_reference(C, golden, monkeypatch, variant, cuda_graphs):
    g = golden("symmetry")
    sh = R.SYMMETRY_SHAPE
    N, T = sh["N"], sh["T"]
    agent = _make_agent(C, variant)
    agent.cuda_graphs = cuda_graphs
    assert [n for n, _ in agent.named_parameters()] == g.np(f"{variant}/param_names").tolist()
    stream = {k: v.to(DEV) for k, v in R.anymal_stream(T, N, seed=24, obs_dim=sh["obs"], p_term=0.1, p_trunc=0.05).items()}
    noise = R.noise_stream(T, N, sh["act"], seed=25).to(DEV)
    step = {"t": 0}
    import cusrl_b200.nn.modules as M

    monkeypatch.setattr(M, "standard_normal_like", lambda mean: noise[step["t"]].reshape(mean.shape).clone())
    returned = []
    for t in range(T):
        step["t"] = t
        returned.append(agent.act(stream["obs"][t]).clone())
        ready = agent.step(stream["obs"][t + 1], stream["reward"][t], stream["terminated"][t], stream["truncated"][t])
    assert ready
    pre = f"{variant}/"
    assert torch.allclose(torch.stack(returned).cpu(), g.t(pre + "returned_action"), rtol=1e-5, atol=2e-5)
    for key in ("observation", "augmented_observation"):
        if pre + f"buffer/{key}" in g.keys():   # pure index / sign work: bit-exact
            assert torch.equal(agent.buffer.storage[key].cpu(), g.t(pre + f"buffer/{key}")), key
    for key in ("action", "action_logp", "action_dist.mean", "value", "augmented_action"):
        if pre + f"buffer/{key}" in g.keys():
            assert torch.allclose(agent.buffer.storage[key].cpu(), g.t(pre + f"buffer/{key}"), rtol=1e-5, atol=2e-5), key
    if variant == "augmentation":   # the mirrored action is an index / sign image of the stored action: exact
        aug = agent.buffer.storage["augmented_action"]
        assert torch.equal(aug[:, :, 0], agent.buffer.storage["action"])
        assert torch.equal(aug[:, :, 1], agent.environment_spec.mirror_action(agent.buffer.storage["action"]))

    import cusrl_b200.sampler as S

    seeded = R.SeededRandperm(seed=26)
    monkeypatch.setattr(S.torch, "randperm", seeded)
    names = g.np(pre + "objective_names").tolist()
    logs = []
    train_step = agent._train_step

    def spy(metadata, batch):
        train_step(metadata, batch)
        o = agent.last_objectives
        assert list(o) == names   # same objectives in the same order (the float order of the loss sum)
        logs.append(torch.stack([o[k].detach().reshape(()) for k in names]))

    agent._train_step = spy
    metrics = agent.update()
    got = torch.stack(logs).double().cpu().numpy()
    ref = g.np(pre + "minibatch_losses")
    assert got.shape == ref.shape
    np.testing.assert_allclose(got[0], ref[0], rtol=2e-5, atol=1e-6)   # before any parameter update
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=5e-6)
    params = dict(agent.named_parameters())
    for name in g.np(pre + "param_names").tolist():
        ours, want = params[name].detach().cpu(), g.t(pre + f"param1/{name}")
        ok = (ours - want).abs() <= 5e-6 + 2e-4 * want.abs()
        assert ok.float().mean().item() >= 0.99 and (ours - want).abs().max().item() <= 4e-3, \
            (name, ok.float().mean().item(), (ours - want).abs().max().item())
    ref_metrics = dict(zip(g.np(pre + "metric_names").tolist(), g.np(pre + "metric_values").tolist()))
    assert set(metrics) == set(ref_metrics), set(metrics) ^ set(ref_metrics)
    for k, v in ref_metrics.items():
        assert metrics[k] == pytest.approx(v, rel=5e-3, abs=5e-6), k
    assert agent.optimizer.param_groups[0]["lr"] == pytest.approx(float(g.np(pre + "lr_after")), rel=1e-9)
    if cuda_graphs:
        assert agent._train_step_graphs.captures >= 1


def test_symmetry_hooks_with_recurrent_agent_and_custom_mirrors(C, monkeypatch):
    """The reference's own smoke tests (cusrl_test/hook/auxiliary/test_symmetry.py:48-64): recurrent agent, custom stacked
    mirror callables (two variants), training runs and stays finite."""
    # same LSTM weights as in the hardware-validated runs of this test (see tests/test_rollout_gpu.py for the reasoning)
    monkeypatch.setattr(C.ModuleInitialization, "_init_rnn", lambda *args, **kwargs: None)
    N, obs_dim, act_dim = 64, 16, 8
    for recurrent in (False, True):
        factory = (C.RecurrentPpoAgentFactory(num_steps_per_update=8, actor_hidden_size=32, critic_hidden_size=32,
                                              actor_num_layers=1, critic_num_layers=1, device=DEV) if recurrent else
                   C.PpoAgentFactory(num_steps_per_update=8, actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128),
                                     device=DEV)).to_underlying()
        factory.register_hook(C.SymmetricDataAugmentation(), before="value_loss")
        env = C.SyntheticEnvironment(N, obs_dim, act_dim, device=DEV, seed=5, p_term=0.05)
        env.spec.mirror_observation = lambda o: torch.stack([o, o.flip(-1)])
        env.spec.mirror_action = lambda a: torch.stack([a, a.flip(-1)])
        agent = factory.from_environment(env)
        history = C.Trainer(env, agent, num_iterations=3).run_training_loop()
        for key in ("Agent/value_loss", "Agent/surrogate_loss", "Agent/kl_divergence"):
            assert np.isfinite(history[-1][key]), (recurrent, key)
        assert agent.buffer.storage["augmented_observation"].shape == (8, N, 3, obs_dim)
