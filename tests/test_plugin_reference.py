"""CPU: the plugin registers with the UNMODIFIED reference (baseline/_ref, or /root/reference in the build container), the
reference's own Trainer accepts the B200 agent factory, reference-side hooks drive the B200 hooks by name, and
checkpoints move both ways."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
from install_reference import reference_path  # noqa: E402

try:
    _PATHS = reference_path()   # baseline/_ref (travels to the GPU box), else the build container's /root/reference
except RuntimeError:
    _PATHS = None
pytestmark = pytest.mark.skipif(_PATHS is None, reason="reference package not available (run tools/install_reference.py)")


@pytest.fixture(scope="module")
def reference():
    added = list(_PATHS)
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
    # the recurrent preset is registered next to it (python -m cusrl train -alg ppo-b200-lstm -m cusrl_b200.plugin)
    lstm = get_experiment("Synthetic-AnymalC-Rough-v0", plugin.RECURRENT_ALGORITHM_NAME).to_training_factory()
    assert type(lstm.agent_factory).__name__ == "RecurrentPpoAgentFactory" and lstm.agent_factory.actor_hidden_size == 256


def test_checkpoints_move_between_the_reference_agent_and_this_one(reference):
    """`agent.state_dict()` has the reference's on-disk layout -- module keys and shapes,
    hook state, and the optimizer in torch.optim.Adam's format with `param_names` -- so a checkpoint trained with the
    reference loads here (parameters + Adam moments land in the flat arenas) and one written here loads there."""
    import torch

    import cusrl_b200 as C
    from cusrl.preset.ppo import PpoAgentFactory as RefFactory

    kwargs = dict(num_steps_per_update=4, actor_hidden_dims=(16, 8), critic_hidden_dims=(16, 8), activation_fn="ELU",
                  lr=1e-3, orthogonal_init=False, desired_kl_divergence=0.015, device="cpu")
    spec = reference.EnvironmentSpec(num_instances=8, observation_dim=5, action_dim=2, autoreset=True, final_state_is_missing=True)
    torch.manual_seed(0)
    ref_agent = RefFactory(**kwargs)(spec)
    for _ in range(4):
        ref_agent.act(torch.randn(8, 5))
        ref_agent.step(torch.randn(8, 5), torch.randn(8, 1), torch.zeros(8, 1, dtype=torch.bool), torch.zeros(8, 1, dtype=torch.bool))
    ref_agent.update()  # 20 Adam steps: the optimizer state is populated
    ref_sd = ref_agent.state_dict()

    ours = C.PpoAgentFactory(**kwargs)(C.EnvironmentSpec(8, 5, 2, autoreset=True, final_state_is_missing=True))
    our_sd = ours.state_dict()
    for module in ("actor", "critic"):
        assert list(our_sd[module]) == list(ref_sd[module])
        assert all(our_sd[module][k].shape == ref_sd[module][k].shape for k in ref_sd[module])
    assert list(our_sd["hook"]) == list(ref_sd["hook"])
    assert list(our_sd) == list(ref_sd) and our_sd["grad_scaler"] == ref_sd["grad_scaler"] == {}   # disabled scaler, same entry
    assert our_sd["optimizer"]["param_groups"][0]["param_names"] == ref_sd["optimizer"]["param_groups"][0]["param_names"]
    assert set(ref_sd["optimizer"]["param_groups"][0]) <= set(our_sd["optimizer"]["param_groups"][0])

    # reference -> here
    ours.load_state_dict(ref_sd)
    for (name, p), (rname, rp) in zip(ours.named_parameters(), ref_agent.named_parameters()):
        assert name == rname and torch.equal(p.detach(), rp.detach())
    opt = ours.optimizer
    assert opt.step_count == 20 and opt.param_groups[0]["lr"] == ref_sd["optimizer"]["param_groups"][0]["lr"]
    flat = opt.flat_param
    for i, (p, off) in enumerate(zip(opt.arena.params, opt.arena.offsets)):
        assert p.data_ptr() == flat[off:].data_ptr()                                  # still a view of the arena
        ref_state = ref_sd["optimizer"]["state"][i]
        assert torch.equal(opt.exp_avg[off:off + p.numel()].view_as(p), ref_state["exp_avg"])
        assert torch.equal(opt.exp_avg_sq[off:off + p.numel()].view_as(p), ref_state["exp_avg_sq"])
    assert ours.hook["adaptive_lr_schedule"].state_dict() == ref_sd["hook"]["adaptive_lr_schedule"]

    # here -> reference: torch.optim.Adam accepts what FlatAdam writes
    back = ours.state_dict()
    fresh = RefFactory(**kwargs)(spec)
    fresh.load_state_dict(back)
    for (_, p), (_, rp) in zip(fresh.named_parameters(), ref_agent.named_parameters()):
        assert torch.equal(p.detach(), rp.detach())
    got = fresh.optimizer.state_dict()["state"]
    for i, st in ref_sd["optimizer"]["state"].items():
        assert float(got[i]["step"]) == float(st["step"]) == 20.0
        assert torch.equal(got[i]["exp_avg"], st["exp_avg"]) and torch.equal(got[i]["exp_avg_sq"], st["exp_avg_sq"])


def test_reference_side_schedule_hooks_drive_the_b200_hooks(reference):
    """the reference's control hooks address hooks BY NAME through ``agent.hook[name]``
    (cusrl/hook/control/schedule.py:12-77) and must keep working when mixed into the B200 hook list."""
    import cusrl_b200 as C
    from cusrl.hook.control.schedule import HookActivationSchedule, HookParameterSchedule

    factory = C.anymal_c_rough_ppo(num_steps_per_update=4, actor_hidden_dims=(16, 128), critic_hidden_dims=(16, 128),
                                   device="cpu").to_underlying()
    factory.register_hook(HookParameterSchedule("ppo_surrogate_loss", "clip_ratio", lambda it: 0.2 if it < 2 else 0.1))
    factory.register_hook(HookActivationSchedule("entropy_loss", lambda it: it < 3))
    agent = factory(C.EnvironmentSpec(8, 5, 2, autoreset=True, final_state_is_missing=True))
    names = [h.name for h in agent.hook]
    assert "ppo_surrogate_loss_clip_ratio_schedule" in names and "entropy_loss_activation_schedule" in names
    assert agent.hook["ppo_surrogate_loss"].clip_ratio == 0.2 and agent.hook["entropy_loss"].active
    agent.metrics.clear()
    agent.hook.apply_schedule(2)
    assert agent.hook["ppo_surrogate_loss"].clip_ratio == 0.1
    assert agent.metrics["ppo_surrogate_loss_clip_ratio"].mean.item() == pytest.approx(0.1)
    agent.hook.apply_schedule(3)
    assert not agent.hook["entropy_loss"].active
    with pytest.raises(ValueError, match="No hook named"):
        bad = C.anymal_c_rough_ppo(device="cpu").to_underlying()
        bad.register_hook(HookParameterSchedule("no_such_hook", "weight", lambda it: 1.0))
        bad(C.EnvironmentSpec(8, 5, 2, autoreset=True, final_state_is_missing=True))
    # non-hooks are still refused
    with pytest.raises(TypeError, match="Expected a Hook instance"):
        C.HookComposite([object()])


@pytest.mark.parametrize("recurrent", [False, True])
def test_export_goes_through_the_reference_exporter_and_matches_the_reference_policy(reference, tmp_path, recurrent):
    """SURVEY.md section 8(f) rank 4, export half: `agent.export(dir, target_format="jit")` hands a plain-torch twin of the
    B200 actor (shared parameters, reference module names) to the REFERENCE's own FlowGraph exporter
    (cusrl/nn/layer/export.py:130-171: actor.pt, actor_stateless.pt, actor.yml); the exported policy equals the reference's
    own actor loaded with the same state_dict.  (ONNX takes the same route; the `onnx` package is not in this image, which
    the reference's exporter itself imports.)"""
    import torch
    import yaml

    import cusrl_b200 as C

    obs_dim, act_dim = 19, 5
    if recurrent:
        ours = C.RecurrentPpoAgentFactory(num_steps_per_update=4, actor_hidden_size=64, critic_hidden_size=64, actor_num_layers=2,
                                          critic_num_layers=1, device="cpu")
        theirs = reference.preset.ppo.RecurrentPpoAgentFactory(num_steps_per_update=4, actor_hidden_size=64, critic_hidden_size=64,
                                                               actor_num_layers=2, critic_num_layers=1, device="cpu")
    else:
        ours = C.PpoAgentFactory(num_steps_per_update=4, actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128),
                                 activation_fn="ELU", device="cpu")
        theirs = reference.preset.ppo.PpoAgentFactory(num_steps_per_update=4, actor_hidden_dims=(64, 128),
                                                      critic_hidden_dims=(64, 128), activation_fn="ELU", device="cpu")
    torch.manual_seed(5)
    agent = ours(C.EnvironmentSpec(4, obs_dim, act_dim, autoreset=True, final_state_is_missing=True))
    from cusrl.template.environment import EnvironmentSpec as RefSpec

    ref_agent = theirs(RefSpec(num_instances=4, observation_dim=obs_dim, action_dim=act_dim, autoreset=True,
                               final_state_is_missing=True))
    ref_agent.actor.load_state_dict(agent.actor.state_dict())
    out = tmp_path / ("lstm" if recurrent else "mlp")
    agent.export(str(out), target_format="jit", verbose=False)
    assert (out / "actor.pt").exists() and (out / "actor_stateless.pt").exists() and (out / "actor.yml").exists()
    info = yaml.safe_load((out / "actor.yml").read_text())
    assert info["observation_dim"] == obs_dim and info["action_dim"] == act_dim and info["is_recurrent"] is recurrent
    assert [list(d)[0] for d in info["inputs"]][0] == "observation" and [list(d)[0] for d in info["outputs"]][0] == "action"
    stateless = torch.jit.load(str(out / "actor_stateless.pt"))
    obs = torch.randn(1, 1, obs_dim)
    with torch.no_grad():
        if recurrent:
            hidden, cell = torch.randn(1, 128) * 0.3, torch.randn(1, 128) * 0.3
            feed = {"observation": obs, "memory_in__hidden": hidden, "memory_in__cell": cell}
            got = stateless(*[feed[list(d)[0]] for d in info["inputs"]])
            want, want_mem = ref_agent.actor(obs, memory={"hidden": hidden, "cell": cell}, forward_type="act_deterministic")
            named = got if isinstance(got, dict) else dict(zip([list(d)[0] for d in info["outputs"]], got))
            torch.testing.assert_close(named["action"], want, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(named["memory_out__hidden"], want_mem["hidden"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(named["memory_out__cell"], want_mem["cell"], rtol=1e-5, atol=1e-6)
        else:
            got = stateless(obs)
            got = got[0] if isinstance(got, (tuple, list)) else got
            got = got["action"] if isinstance(got, dict) else got
            want, _ = ref_agent.actor(obs, forward_type="act_deterministic")
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="Unsupported export format"):
        agent.export(str(out), target_format="tflite")


def test_export_standalone_torchscript_without_the_reference(tmp_path, monkeypatch):
    """Without the reference package `target_format="jit"` traces the same twin directly; ONNX is refused, not approximated."""
    import builtins

    import torch

    import cusrl_b200 as C

    real_import = builtins.__import__

    def no_reference(name, *a, **k):
        if name == "cusrl" or name.startswith("cusrl."):
            raise ImportError(name)
        return real_import(name, *a, **k)

    monkeypatch.setattr(builtins, "__import__", no_reference)
    torch.manual_seed(6)
    agent = C.PpoAgentFactory(num_steps_per_update=4, actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), activation_fn="ELU",
                              device="cpu")(C.EnvironmentSpec(4, 19, 5, autoreset=True, final_state_is_missing=True))
    agent.export(str(tmp_path), target_format="jit", verbose=False)
    traced = torch.jit.load(str(tmp_path / "actor.pt"))
    obs = torch.randn(3, 19)
    lins = agent.actor.backbone.linears()
    h = obs
    for lin in lins:
        h = torch.nn.functional.elu(lin(h))
    torch.testing.assert_close(traced(obs), agent.actor.distribution.mean_head(h), rtol=1e-6, atol=1e-6)
    with pytest.raises(RuntimeError, match="ONNX export goes through the reference"):
        agent.export(str(tmp_path), target_format="onnx", verbose=False)


def test_module_initialization_matches_the_reference_hook_seed_for_seed(reference):
    """`ModuleInitialization` (cusrl/hook/control/initialization.py:66-125) initialises linear, recurrent, attention and
    convolution layers; this implementation must draw the same numbers from the same random stream on the same modules --
    in particular the recurrent preset's LSTM gets orthogonal weights (gain sqrt 2) and zero biases, not torch's default."""
    from types import SimpleNamespace

    import torch
    from torch import nn

    import cusrl_b200 as C

    def modules():
        torch.manual_seed(3)
        actor = nn.Module()
        actor.backbone = nn.LSTM(7, 8, 2)
        actor.distribution = nn.Module()
        actor.distribution.mean_head = nn.Linear(8, 3)
        critic = nn.Sequential(nn.GRU(7, 8, 1), nn.Linear(8, 1), nn.MultiheadAttention(8, 2), nn.Conv2d(2, 3, 3))
        return actor, critic

    results = []
    for cls in (reference.hook.ModuleInitialization, C.ModuleInitialization):
        actor, critic = modules()
        torch.manual_seed(11)
        hook = cls(scale=1.3, scale_dist=0.05)
        hook.agent = SimpleNamespace(actor=actor, critic=critic)
        hook.init()
        results.append(([p.detach().clone() for m in (actor, critic) for p in m.parameters()], torch.rand(4)))
    (ref_params, ref_next), (our_params, our_next) = results
    assert len(ref_params) == len(our_params) and torch.equal(ref_next, our_next)      # same position in the random stream
    for a, b in zip(ref_params, our_params):
        assert torch.equal(a, b)

    # and through the recurrent preset: both presets leave an LSTM with orthogonal weights and zero biases
    spec = C.EnvironmentSpec(4, 19, 5, autoreset=True, final_state_is_missing=True)
    agent = C.RecurrentPpoAgentFactory(device="cpu", actor_hidden_size=64, critic_hidden_size=64)(spec)
    lstm = agent.actor.backbone.rnn
    w_hh = lstm.weight_hh_l0.detach()                       # [4H, H]: orthonormal columns times the gain
    assert torch.allclose(w_hh.T @ w_hh, 2.0 * torch.eye(64), atol=1e-4)
    assert float(lstm.bias_ih_l0.abs().max()) == 0.0 and float(lstm.bias_hh_l1.abs().max()) == 0.0


@pytest.mark.parametrize("with_reference", [True, False])
def test_export_runs_the_hooks_export_callbacks(reference, tmp_path, monkeypatch, with_reference):
    """``ActorCritic.export`` calls ``hook.pre_export(graph)`` / ``hook.post_export(graph)`` (actor_critic.py:365,390): the
    B200 ``ObservationNormalization`` adds its running statistics in front of the actor (observation.py:248-255) and a
    REFERENCE-side user hook appends a node behind it -- through the reference's FlowGraph, and through the stand-in chain
    when the reference is not importable."""
    import torch
    import yaml
    from torch import nn

    import cusrl_b200 as C

    class Halve(nn.Module):
        def forward(self, input):
            return input * 0.5

    class HalveAction(reference.template.Hook):          # the reference's own Hook base class
        def post_export(self, graph):
            graph.add_node(Halve(), module_name="halve", input_names={"input": "action"}, output_names="action",
                           expose_outputs=False)

    torch.manual_seed(8)
    factory = C.PpoAgentFactory(num_steps_per_update=4, actor_hidden_dims=(32, 64), critic_hidden_dims=(32, 64), activation_fn="ELU",
                                normalize_observation=True, device="cpu").to_underlying()
    factory.register_hook(HalveAction())
    agent = factory(C.EnvironmentSpec(4, 19, 5, autoreset=True, final_state_is_missing=True))
    rms = agent.hook["observation_normalization"].observation_rms
    with torch.no_grad():
        rms.mean.copy_(torch.randn(19))
        rms.var.copy_(torch.rand(19) + 0.5)
        rms.std.copy_(rms.var.sqrt())
    obs = torch.randn(1, 1, 19)
    h = (obs - rms.mean) / rms.std
    clamp = getattr(rms, "clamp", None)
    if clamp is not None:
        h = h.clamp(-clamp, clamp)
    for lin in agent.actor.backbone.linears():
        h = torch.nn.functional.elu(lin(h))
    want = agent.actor.distribution.mean_head(h) * 0.5

    if with_reference:
        agent.export(str(tmp_path), target_format="jit", verbose=False)
        info = yaml.safe_load((tmp_path / "actor.yml").read_text())
        assert [list(d)[0] for d in info["inputs"]] == ["observation"]
        got = torch.jit.load(str(tmp_path / "actor_stateless.pt"))(obs)
        got = got[0] if isinstance(got, (tuple, list)) else got
        got = got["action"] if isinstance(got, dict) else got
    else:
        import builtins

        real_import = builtins.__import__

        def no_reference(name, *a, **k):
            if name == "cusrl" or name.startswith("cusrl."):
                raise ImportError(name)
            return real_import(name, *a, **k)

        monkeypatch.setattr(builtins, "__import__", no_reference)
        agent.export(str(tmp_path), target_format="jit", verbose=False)
        got = torch.jit.load(str(tmp_path / "actor.pt"))(obs)
    torch.testing.assert_close(got, want.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("options", [
    {}, {"final_missing": True}, {"symmetric": True}, {"max_count": 40}, {"subset": slice(0, 6)}, {"subset": [9, 2, 4, 0, 7, 5]},
    {"groups": True}, {"symmetric": True, "subset": slice(0, 6)}], ids=lambda o: "+".join(o) or "plain")
def test_observation_normalization_options_match_the_live_reference_hook(reference, options):
    """Every specification option `ObservationNormalization` reads (observation.py:59-255; the cases of
    cusrl_test/hook/mdp/test_observation_normalization.py: symmetry, observation-is-subset-of-state as slice and index list,
    statistic groups, excluded indices, a count window, missing final states) driven step by step next to the REFERENCE's own
    hook on the same inputs: normalised tensors and running statistics must be identical (CPU tensors: the host logic and
    torch arithmetic of this hook; the kernels behind the same code on CUDA tensors are covered by tests/test_obsnorm.py)."""
    from types import SimpleNamespace

    import torch

    import cusrl_b200 as C

    obs_dim, state_dim, N = 6, 10, 16

    def run(hook_cls, mirror_cls):
        spec = dict(final_state_is_missing=options.get("final_missing", False), mirror_observation=None, mirror_state=None,
                    observation_is_subset_of_state=options.get("subset"), observation_stat_groups=(), state_stat_groups=(),
                    observation_normalization_excluded_indices=None, state_normalization_excluded_indices=None)
        if options.get("symmetric"):
            spec["mirror_observation"] = mirror_cls([1, 0, 2, 4, 3, 5], [False, False, True, False, False, False])
            spec["mirror_state"] = mirror_cls([1, 0, 2, 4, 3, 5, 7, 6, 8, 9],
                                              [False, False, True, False, False, False, True, True, False, True])
        if options.get("groups"):
            spec.update(observation_stat_groups=(slice(0, 3),), state_stat_groups=((0, 3), slice(6, 10)),
                        observation_normalization_excluded_indices=slice(4, 6), state_normalization_excluded_indices=(4, 5))
        agent = SimpleNamespace(environment_spec=SimpleNamespace(**spec), observation_dim=obs_dim, state_dim=state_dim,
                                has_state=True, inference_mode=False, setup_module=lambda m: m, to_tensor=torch.as_tensor,
                                device=torch.device("cpu"), parallelism=N)
        hook = hook_cls(options.get("max_count"))
        hook.pre_init(agent)
        hook.init()
        g = torch.Generator().manual_seed(3)
        subset = options.get("subset")
        trace = []
        for _ in range(6):
            state = torch.randn(N, state_dim, generator=g) * 2 + 1
            observation = state[:, subset].clone() if subset is not None else torch.randn(N, obs_dim, generator=g) * 3 - 1
            transition = {"observation": observation, "state": state}
            hook.pre_act(transition)
            next_state = torch.randn(N, state_dim, generator=g) * 2 + 1
            next_observation = next_state[:, subset].clone() if subset is not None else torch.randn(N, obs_dim, generator=g) * 3 - 1
            step = {"next_observation": next_observation, "next_state": next_state, "done": torch.rand(N, 1, generator=g) < 0.3}
            hook.post_step(step)
            trace += [transition["observation"], transition["state"], step["next_observation"], step["next_state"],
                      hook.observation_rms.mean.clone(), hook.observation_rms.var.clone(), hook.state_rms.mean.clone(),
                      hook.state_rms.var.clone(), torch.tensor(float(hook.observation_rms.count)),
                      torch.tensor(float(hook.state_rms.count))]
        return trace

    theirs = run(reference.hook.ObservationNormalization, reference.hook.auxiliary.symmetry.MirrorDef)
    ours = run(C.ObservationNormalization, C.MirrorDef)
    assert len(theirs) == len(ours)
    for i, (a, b) in enumerate(zip(theirs, ours)):
        torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-6, msg=lambda m, i=i: f"trace entry {i}: {m}")


@pytest.mark.parametrize("kwargs", [
    dict(desired_kl_divergence=0.01), dict(desired_kl_divergence=0.015, max_kl_divergence=0.04),
    dict(desired_kl_divergence=0.01, warmup_iterations=5, initial_scale=0.2), dict(desired_kl_divergence=0.01, scale_all_params=True),
    dict(desired_kl_divergence=0.02, scale_factor=0.5, threshold=0.5)], ids=lambda k: "+".join(sorted(k)))
def test_adaptive_lr_schedule_follows_the_live_reference_hook(reference, kwargs):
    """`AdaptiveLRSchedule` (lr_schedule.py:19-239) with each of its options, next to the REFERENCE's own hook inside the
    reference's own agent, over 40 iterations of the same KL statistics: learning-rate scale, every param group's learning
    rate and the accept / reject decisions (`max_kl_divergence`) must be identical."""
    import random

    import cusrl_b200 as C

    rng = random.Random(1)
    kls = [0.01 * 2 ** rng.uniform(-3, 3) for _ in range(40)]

    def drive(theirs: bool):
        if theirs:
            factory = reference.preset.ppo.PpoAgentFactory(actor_hidden_dims=(16, 8), critic_hidden_dims=(16, 8),
                                                           desired_kl_divergence=None, device="cpu").to_underlying()
            factory.register_hook(reference.hook.AdaptiveLRSchedule(**kwargs))
            agent = factory(reference.EnvironmentSpec(19, 5, num_instances=4))
        else:
            factory = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), desired_kl_divergence=None,
                                        device="cpu").to_underlying()
            factory.register_hook(C.AdaptiveLRSchedule(**kwargs))
            agent = factory(C.EnvironmentSpec(19, 5, num_instances=4))       # the reference's call style
        hook = agent.hook["adaptive_lr_schedule"]
        trace = []
        for iteration, kl in enumerate(kls):
            agent.iteration = iteration
            hook.apply_schedule(iteration)
            hook.pre_update(agent.buffer)
            agent.metrics.clear()
            agent.record(kl_divergence=kl)
            hook.post_update()
            rejected = float(agent.metrics["update_rejected"].mean) if "update_rejected" in agent.metrics.keys() else None
            trace.append((hook._lr_scale, [group["lr"] for group in agent.optimizer.param_groups], rejected))
        return trace

    for step, (a, b) in enumerate(zip(drive(True), drive(False))):
        assert b[0] == pytest.approx(a[0], rel=1e-12), step
        assert b[1] == pytest.approx(a[1], rel=1e-12), step
        assert b[2] == a[2], step


@pytest.mark.parametrize("temporal", [False, True])
@pytest.mark.parametrize("T,N,epochs,mini_batches,shuffle", [(6, 8, 3, 4, True), (5, 7, 2, 3, True), (4, 9, 3, (2, 3, 5), True),
                                                              (6, 8, 2, 4, False), (3, 5, 1, 1, True)])
def test_sampler_index_streams_follow_the_live_reference_sampler(reference, temporal, T, N, epochs, mini_batches, shuffle):
    """Same seed -> the same minibatches as the REFERENCE's own samplers (mini_batch_sampler.py:52-114): permutation draws
    (one per epoch, `randperm(out=)` from the second on), the dropped remainder, per-epoch minibatch counts, no shuffling,
    metadata; flat `t*N + n` indices for the transition sampler, environment columns for the temporal one."""
    import torch

    import cusrl_b200 as C

    Ref = reference.TemporalMiniBatchSampler if temporal else reference.MiniBatchSampler
    Ours = C.TemporalMiniBatchSampler if temporal else C.MiniBatchSampler
    if (N if temporal else T * N) < (max(mini_batches) if isinstance(mini_batches, tuple) else mini_batches):
        pytest.skip("more minibatches than samples")
    marker = torch.arange(T * N, dtype=torch.float32).reshape(T, N, 1)           # value = flat index t*N + n
    ref_buffer = reference.template.Buffer(T, N, device="cpu")
    our_buffer = C.Buffer(T, N, device="cpu")
    for t in range(T):
        ref_buffer.push({"marker": marker[t]})
        our_buffer.push({"marker": marker[t]})
    torch.manual_seed(123)
    theirs = [(meta, batch["marker"]) for meta, batch in Ref(epochs, mini_batches, shuffle)(ref_buffer)]
    torch.manual_seed(123)
    # the index slices are views of ONE permutation buffer that `randperm(out=)` rewrites every epoch: copy them as they come
    ours = [(meta, idx.clone()) for meta, idx in Ours(epochs, mini_batches, shuffle).indices(our_buffer)]
    assert len(theirs) == len(ours)
    for (ref_meta, ref_rows), (meta, idx) in zip(theirs, ours):
        assert meta == ref_meta
        if temporal:      # the reference yields [T, n_mb, 1] sequences of the selected environment columns
            assert torch.equal(ref_rows[0, :, 0].long(), idx) and ref_rows.shape == (T, idx.numel(), 1)
        else:
            assert torch.equal(ref_rows[:, 0].long(), idx)
    # and both leave torch's generator at the same position
    torch.manual_seed(123)
    list(Ref(epochs, mini_batches, shuffle)(ref_buffer))
    after_theirs = torch.rand(3)
    torch.manual_seed(123)
    list(Ours(epochs, mini_batches, shuffle).indices(our_buffer))
    assert torch.equal(torch.rand(3), after_theirs)


@pytest.mark.parametrize("kind", ["ppo", "anymal", "lstm", "rnd", "state"])
def test_same_seed_gives_the_reference_agents_initial_parameters(reference, kind):
    """Drop-in at construction: with the same torch seed the agent built by this package has the SAME parameter names, the
    same initial values bit for bit (module creation order, default initialisers, `ModuleInitialization`'s orthogonal pass
    over linear and recurrent layers, the RND hook's Xavier initialisation, `init_distribution_std`) and leaves torch's
    generator at the same position as the reference's agent -- default PPO preset, the Anymal-C preset values
    (zoo/isaaclab/locomotion.py:48-59), the recurrent preset, PPO + RND, PPO with a critic state."""
    import torch

    import cusrl_b200 as C

    def build(theirs: bool):
        pkg = reference if theirs else C
        presets = reference.preset.ppo if theirs else C
        torch.manual_seed(42)
        if kind == "anymal":
            factory = presets.PpoAgentFactory(num_steps_per_update=24, actor_hidden_dims=(512, 256, 128), critic_hidden_dims=(512, 256, 128),
                                              activation_fn="ELU", lr=1e-3, sampler_epochs=5, sampler_mini_batches=4, orthogonal_init=False,
                                              entropy_loss_weight=0.005, desired_kl_divergence=0.015, device="cpu")
        elif kind == "lstm":
            factory = presets.RecurrentPpoAgentFactory(device="cpu")
        elif kind == "rnd":
            factory = presets.PpoAgentFactory(device="cpu").to_underlying()
            factory.register_hook(pkg.hook.RandomNetworkDistillation(pkg.Mlp.Factory([128, 128]), output_dim=16, reward_scale=0.1),
                                  before="value_computation")
        else:
            factory = presets.PpoAgentFactory(device="cpu", init_distribution_std=0.6 if kind == "state" else None)
        agent = factory(pkg.EnvironmentSpec(235, 12, num_instances=8, state_dim=40 if kind == "state" else None))
        return dict(agent.named_parameters()), torch.rand(3)

    (ref_params, ref_next), (our_params, our_next) = build(True), build(False)
    assert list(our_params) == list(ref_params)
    for name, value in ref_params.items():
        assert torch.equal(our_params[name].detach(), value.detach()), name
    assert torch.equal(our_next, ref_next)
