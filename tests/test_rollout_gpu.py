"""GPU: the fused rollout step (template/rollout.py) against the generic act / step flow (the
reference's control flow, actor_critic.py:227-291): identical buffer contents from identical inputs and noise, for device
inputs, for pinned-host inputs (including the "next_observation handed back as observation" shortcut that skips the second
H2D copy), with a separate critic state, and with the fall-backs that must take the generic path."""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
T, N, OBS, ACT = 6, 512, 235, 12


@pytest.fixture(scope="module")
def C():
    from cusrl_b200 import build

    build.build()
    import cusrl_b200

    return cusrl_b200


def _agent(C, fused: bool, state_dim=None, seed=0):
    torch.manual_seed(seed)
    spec = C.EnvironmentSpec(N, OBS, ACT, state_dim=state_dim, autoreset=True, final_state_is_missing=True)
    agent = C.anymal_c_rough_ppo(num_steps_per_update=T, device=DEV)(spec)
    agent.fused_rollout = fused
    return agent


def _stream(seed=1, state_dim=None):
    g = torch.Generator().manual_seed(seed)
    d = {"obs": torch.randn(T + 1, N, OBS, generator=g), "reward": torch.randn(T, N, 1, generator=g),
         "terminated": torch.rand(T, N, 1, generator=g) < 0.05, "truncated": torch.rand(T, N, 1, generator=g) < 0.02,
         "noise": torch.randn(T, N, ACT, generator=g)}
    if state_dim:
        d["state"] = torch.randn(T + 1, N, state_dim, generator=g)
    return d


def _rollout(agent, data, monkeypatch, where: str, hand_back: bool = True):
    import cusrl_b200.nn.modules as M

    step = {"t": 0}
    monkeypatch.setattr(M, "standard_normal_like", lambda mean: data["noise"][step["t"]].to(mean.device).reshape(mean.shape).clone())
    put = (lambda x: x.to(DEV)) if where == "cuda" else (lambda x: x.pin_memory())
    obs = [put(data["obs"][t]) for t in range(T + 1)]
    state = [put(data["state"][t]) for t in range(T + 1)] if "state" in data else [None] * (T + 1)
    actions = []
    for t in range(T):
        step["t"] = t
        # hand_back: act() receives the very tensor step() got as next_observation (the Trainer's hand-over)
        o = obs[t] if hand_back else put(data["obs"][t])
        s = state[t] if (hand_back or state[t] is None) else put(data["state"][t])
        a = agent.act(o, s)
        assert a.device == o.device and a.shape == (N, ACT)
        actions.append(a.cpu().clone())
        ready = agent.step(obs[t + 1], put(data["reward"][t]), put(data["terminated"][t]), put(data["truncated"][t]), state[t + 1])
    assert ready and agent.buffer.full
    return actions


def _compare(a, b, exact=("observation", "next_observation", "state", "next_state", "reward", "terminated", "truncated",
                         "done", "action_dist.mean", "action_dist.std", "action", "value")):
    assert set(a.buffer.storage) == set(b.buffer.storage)
    for key in a.buffer.storage:
        x, y = a.buffer.storage[key], b.buffer.storage[key]
        if key in exact:
            assert torch.equal(x, y), key
        else:
            assert torch.allclose(x, y, rtol=1e-6, atol=1e-5), (key, (x - y).abs().max().item())


@pytest.mark.parametrize("where", ["cuda", "host"])
def test_fused_rollout_matches_generic(C, monkeypatch, where):
    data = _stream()
    generic, fused = _agent(C, False), _agent(C, True)
    act_g = _rollout(generic, data, monkeypatch, "cuda")
    act_f = _rollout(fused, data, monkeypatch, where)
    assert generic._fused_rollout is None
    assert fused._fused_rollout.fast_steps == T - 1   # the first step allocates the leaves through the generic push
    _compare(generic, fused)
    for x, y in zip(act_g, act_f):
        assert torch.equal(x, y)
    # padding columns of the observation slots are zero (legal zero-padded TMA rows)
    back = fused.buffer.backing("observation")
    assert back.shape[-1] == 236 and float(back[..., 235:].abs().max()) == 0.0
    # a second rollout (slots overwritten in place, cursor wrapped) and then an update
    _rollout(generic, _stream(seed=2), monkeypatch, "cuda")
    _rollout(fused, _stream(seed=2), monkeypatch, where, hand_back=False)
    _compare(generic, fused)
    torch.manual_seed(5)
    mg = generic.update()
    torch.manual_seed(5)
    mf = fused.update()
    for k in mg:
        assert mf[k] == pytest.approx(mg[k], rel=1e-4, abs=1e-6), k


def test_fused_rollout_with_critic_state(C, monkeypatch):
    data = _stream(state_dim=48)
    generic, fused = _agent(C, False, state_dim=48), _agent(C, True, state_dim=48)
    _rollout(generic, data, monkeypatch, "cuda")
    _rollout(fused, data, monkeypatch, "host")
    assert fused._fused_rollout.fast_steps == T - 1
    _compare(generic, fused)


def test_deterministic_and_fallbacks(C, monkeypatch):
    data = _stream()
    agent = _agent(C, True)
    _rollout(agent, data, monkeypatch, "cuda")
    fast = agent._fused_rollout.fast_steps
    # numpy observations take the generic path and come back as numpy (agent.py:376-391)
    a = agent.act(data["obs"][0].numpy())
    assert isinstance(a, np.ndarray) and a.dtype == np.float32
    agent.step(data["obs"][1].numpy(), data["reward"][0].numpy(), data["terminated"][0].numpy(), data["truncated"][0].numpy())
    assert agent._fused_rollout.fast_steps == fast
    # non-bool flags are refused exactly like the reference does (actor_critic.py:273-276)
    agent.act(data["obs"][1].to(DEV))
    with pytest.raises(TypeError, match="terminated"):
        agent.step(data["obs"][2].to(DEV), data["reward"][1].to(DEV), data["terminated"][1].float().to(DEV), data["truncated"][1].to(DEV))
    # a user hook that post-processes actions disables the fused path for the whole agent
    class Clip(C.Hook):
        def post_act(self, transition):
            transition["action"] = transition["action"].clamp(-1, 1)

    factory = C.anymal_c_rough_ppo(num_steps_per_update=T, device=DEV).to_underlying()
    factory.register_hook(Clip())
    other = factory(C.EnvironmentSpec(N, OBS, ACT, autoreset=True, final_state_is_missing=True))
    acts = _rollout(other, data, monkeypatch, "cuda")
    assert other._fused_rollout.fast_steps == 0 and float(torch.stack(acts).abs().max()) <= 1.0


def test_fused_recurrent_rollout_matches_generic(C, monkeypatch):
    """FusedRecurrentRollout (LSTM actor and critic): same buffer contents -- including the six recurrent-memory leaves with
    their in-place episode resets -- same returned actions and the same update as the generic act / step flow."""
    # the kernels' numerics in this comparison were validated on a B200 with the LSTM at torch's default initialisation:
    # keep those weights (ModuleInitialization's orthogonal initialisation of recurrent layers is host-side torch code,
    # covered on the CPU by tests/test_plugin_reference.py against the reference's hook, seed for seed)
    monkeypatch.setattr(C.ModuleInitialization, "_init_rnn", lambda *args, **kwargs: None)
    Tn, Nn = 8, 384

    def agent_of(fused: bool):
        torch.manual_seed(3)
        spec = C.EnvironmentSpec(Nn, OBS, ACT, autoreset=True, final_state_is_missing=True)
        agent = C.RecurrentPpoAgentFactory(num_steps_per_update=Tn, actor_hidden_size=256, critic_hidden_size=128, actor_num_layers=2,
                                           critic_num_layers=1, sampler_mini_batches=2, sampler_epochs=2, device=DEV)(spec)
        agent.fused_rollout = fused
        return agent

    def stream(seed):
        g = torch.Generator().manual_seed(seed)
        return {"obs": torch.randn(Tn + 1, Nn, OBS, generator=g).to(DEV), "reward": torch.randn(Tn, Nn, 1, generator=g).to(DEV),
                "terminated": (torch.rand(Tn, Nn, 1, generator=g) < 0.1).to(DEV),
                "truncated": (torch.rand(Tn, Nn, 1, generator=g) < 0.05).to(DEV), "noise": torch.randn(Tn, Nn, ACT, generator=g).to(DEV)}

    import cusrl_b200.nn.modules as M

    def rollout(agent, data):
        step = {"t": 0}
        monkeypatch.setattr(M, "standard_normal_like", lambda mean: data["noise"][step["t"]].reshape(mean.shape).clone())
        actions = []
        for t in range(Tn):
            step["t"] = t
            actions.append(agent.act(data["obs"][t]).clone())
            ready = agent.step(data["obs"][t + 1], data["reward"][t], data["terminated"][t], data["truncated"][t])
        assert ready
        return torch.stack(actions)

    generic, fused = agent_of(False), agent_of(True)
    for r, seed in enumerate((11, 12, 13)):   # three rollouts: memories carry over, slots are overwritten in place
        data = stream(seed)
        a_g, a_f = rollout(generic, data), rollout(fused, data)
        # bit-identical until the first update; afterwards to fp32 rounding (the update itself is not bit-reproducible from
        # run to run: the gradient norm is an atomic fp64 sum, so the clip coefficient can differ in its last bit)
        same = torch.equal if r < 2 else (lambda x, y: torch.allclose(x.float(), y.float(), rtol=1e-4, atol=1e-5))
        assert same(a_g, a_f), r
        assert set(generic.buffer.storage) == set(fused.buffer.storage)
        for key in generic.buffer.storage:
            x, y = generic.buffer.storage[key], fused.buffer.storage[key]
            if key == "action_logp":
                assert torch.allclose(x, y, rtol=1e-5, atol=1e-5), (r, key)
            else:
                assert same(x, y), (r, key, (x.float() - y.float()).abs().max().item())
        for mem_g, mem_f in ((generic.actor_memory, fused.actor_memory),
                             (generic.hook["value_computation"]._critic_memory, fused.hook["value_computation"]._critic_memory)):
            assert same(mem_g["hidden"], mem_f["hidden"]) and same(mem_g["cell"], mem_f["cell"])
        if r == 0:
            continue   # the second rollout overwrites the buffer in place before any update
        torch.manual_seed(20 + r)
        mg = generic.update()
        torch.manual_seed(20 + r)
        mf = fused.update()
        for k in mg:
            assert mf[k] == pytest.approx(mg[k], rel=1e-3, abs=1e-5), (r, k)
    from cusrl_b200.template.rollout import FusedRecurrentRollout

    assert isinstance(fused._fused_rollout, FusedRecurrentRollout)
    assert fused._fused_rollout.fast_steps == 3 * Tn - 2   # all but the two allocating steps of the first rollout
    assert generic._fused_rollout is None
