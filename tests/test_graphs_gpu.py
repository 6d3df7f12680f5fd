"""GPU: the CUDA-graph replay of the train step (template/graphs.py, opt-in) against the eager step -- same seeds, same
rollouts, several iterations so that the learning rate (KL-adaptive schedule) and Adam's step count change between
replays -- and the device-scalar Adam kernel against the host-scalar one."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def C():
    from cusrl_b200 import build

    build.build()
    import cusrl_b200

    return cusrl_b200


def test_adam_device_scalars_match_host_scalars(C):
    from cusrl_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(0)
    n = 10007
    p0 = torch.randn(n, device=DEV, generator=g)
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb, vb = (torch.zeros(n, device=DEV) for _ in range(4))
    step_dev = torch.zeros(1, dtype=torch.int64, device=DEV)
    lr_dev = torch.zeros(1, device=DEV)
    coef = torch.full((1,), 0.5, device=DEV)
    for step, lr in enumerate([1e-3, 1e-3, 3e-4, 2e-3, 1e-5], start=1):
        grad = torch.randn(n, device=DEV, generator=g)
        use_coef = coef if step % 2 else None
        ops.adam_step_(pa, grad, ma, va, step, lr, (0.9, 0.999), 1e-8, 0.01, coef=use_coef)
        step_dev.add_(1)
        lr_dev.fill_(lr)
        ops.adam_step_dev_(pb, grad, mb, vb, step_dev, lr_dev, (0.9, 0.999), 1e-8, 0.01, coef=use_coef)
        # the bias corrections are the same doubles rounded to float on the host / on the device
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-8), step
        assert torch.equal(ma, mb) and torch.equal(va, vb)


def _run(C, make_factory, envs, iters, cuda_graphs):
    from bench import RolloutData, run_iteration

    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    env = C.SyntheticEnvironment(envs, device=dev, seed=7)
    agent = make_factory(C, dev).from_environment(env)
    agent.cuda_graphs = cuda_graphs
    data = RolloutData(24, envs, dev, seed=11, pinned_host=False)
    torch.manual_seed(123)  # exploration noise and minibatch permutations
    history = [run_iteration(agent, data) for _ in range(iters)]
    return agent, history


def _mlp(C, dev):
    return C.anymal_c_rough_ppo(device=dev)


def _rnd(C, dev):
    factory = C.anymal_c_rough_ppo(device=dev).to_underlying()
    factory.register_hook(C.RandomNetworkDistillation(C.Mlp.Factory([64, 64]), output_dim=16, reward_scale=0.1),
                          before="value_computation")
    return factory


def _lstm(C, dev):
    return C.RecurrentPpoAgentFactory(device=dev, actor_hidden_size=128, critic_hidden_size=128)


@pytest.mark.parametrize("name,make,envs,iters", [("mlp", _mlp, 512, 3), ("rnd", _rnd, 256, 2), ("lstm", _lstm, 64, 2)])
def test_graph_replay_matches_eager_training(C, monkeypatch, name, make, envs, iters):
    # the kernels' numerics in this comparison were validated on a B200 with the LSTM at torch's default initialisation:
    # keep those weights (ModuleInitialization's orthogonal initialisation of recurrent layers is host-side torch code,
    # covered on the CPU by tests/test_plugin_reference.py against the reference's hook, seed for seed)
    monkeypatch.setattr(C.ModuleInitialization, "_init_rnn", lambda *args, **kwargs: None)
    eager, hist_e = _run(C, make, envs, iters, cuda_graphs=False)
    graphed, hist_g = _run(C, make, envs, iters, cuda_graphs=True)
    runner = graphed._train_step_graphs
    steps = 20 * iters
    assert runner is not None and runner.captures == 1 and runner.replays == steps - runner.WARMUP
    assert graphed.optimizer.step_count == eager.optimizer.step_count == steps
    assert int(graphed.optimizer.step_dev.item()) == steps
    assert graphed.optimizer.param_groups[0]["lr"] == pytest.approx(eager.optimizer.param_groups[0]["lr"], rel=1e-6)
    for it, (me, mg) in enumerate(zip(hist_e, hist_g)):
        assert set(me) == set(mg)
        for key in me:
            assert mg[key] == pytest.approx(me[key], rel=2e-4, abs=2e-6), (it, key)
    for (n1, p1), (n2, p2) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert n1 == n2
        close = (p1 - p2).abs() <= 1e-5 + 1e-4 * p1.abs()
        assert close.float().mean().item() >= 0.999, n1
    sd = graphed.optimizer.state_dict()
    assert float(sd["state"][0]["step"]) == steps


def test_schedule_change_recaptures(C):
    """A hook hyper-parameter is a captured constant: changing it must discard the graph, not replay the old value."""
    agent, _ = _run(C, _mlp, 256, 1, cuda_graphs=True)
    runner = agent._train_step_graphs
    assert runner.captures == 1
    agent.hook["ppo_surrogate_loss"].update_attribute("clip_ratio", 0.1)
    from bench import RolloutData, run_iteration

    data = RolloutData(24, 256, torch.device("cuda", 0), seed=12, pinned_host=False)
    metrics = run_iteration(agent, data)
    assert runner.captures == 2
    assert all(v == v for v in metrics.values())
