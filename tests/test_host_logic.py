"""CPU: host-side mirror of the reference's plugin surface (Buffer, samplers' index generation, hook
registry / ordering, factories).  Modelled on the reference's own tests: cusrl_test/template/test_buffer.py,
cusrl_test/sampler/test_mini_batch_sampler.py, cusrl_test/hook/on_policy/test_gae.py:34-46."""

from __future__ import annotations

import pytest
import torch

import cusrl_b200 as C
from cusrl_b200.template.buffer import padded_width


def make_buffer(T=6, N=8):
    buf = C.Buffer(T, N, device="cpu")
    for t in range(T):
        buf.push({
            "observation": torch.full((N, 3), float(t)),
            "flag": torch.zeros(N, 1, dtype=torch.bool),
            "action_dist": {"mean": torch.ones(N, 2) * t, "std": torch.ones(N, 2)},
            "skipped": None,
        })
    return buf


def test_buffer_push_wraps_and_preserves_dtype():
    buf = make_buffer()
    assert buf.full and buf.cursor == 0
    assert buf["flag"].dtype == torch.bool
    assert buf["observation"].shape == (6, 8, 3)
    assert set(buf.storage) == {"observation", "flag", "action_dist.mean", "action_dist.std"}
    assert buf["action_dist"]["mean"][4, 0, 0] == 4.0
    assert "skipped" not in buf


def test_buffer_shape_validation_and_schema():
    buf = C.Buffer(4, 3, device="cpu")
    with pytest.raises(ValueError, match="Parallelism mismatch"):
        buf.push({"x": torch.zeros(5, 2)})
    buf.push({"x": torch.zeros(3, 2)})
    with pytest.raises(ValueError, match="Schema mismatch"):
        buf.push({"x": {"a": torch.zeros(3, 2)}})
    with pytest.raises(ValueError, match="Capacity mismatch"):
        buf["y"] = torch.zeros(5, 3, 1)
    buf["y"] = torch.ones(4, 3, 1)
    assert buf["y"].sum() == 12
    del buf["y"]
    assert "y" not in buf and "y" not in buf.storage


def test_buffer_pads_wide_rows_to_16_bytes():
    assert padded_width(235, torch.float32) == 236
    assert padded_width(12, torch.float32) == 12     # < 64 bytes: left dense
    assert padded_width(512, torch.float32) == 512
    buf = C.Buffer(2, 4, device="cpu")
    buf.push({"observation": torch.randn(4, 235)})
    assert buf["observation"].shape == (2, 4, 235)
    assert buf.backing("observation").shape == (2, 4, 236)
    assert buf["observation"].stride() == (944, 236, 1)
    assert buf.backing("observation")[..., 235:].abs().sum() == 0


def test_sampler_requires_full_buffer():
    buf = C.Buffer(4, 2, device="cpu")
    buf.push({"x": torch.zeros(2, 1)})
    with pytest.raises(RuntimeError, match="requires a full buffer"):
        next(iter(C.MiniBatchSampler().indices(buf)))


def test_sampler_argument_validation():
    with pytest.raises(ValueError):
        C.MiniBatchSampler(num_epochs=0)
    with pytest.raises(ValueError):
        C.MiniBatchSampler(num_mini_batches=0)
    with pytest.raises(ValueError):
        C.MiniBatchSampler(num_epochs=2, num_mini_batches=[1, 2, 3])


def test_sampler_permutations_bit_exact_with_reference(golden):
    """Same seed, same device -> the same randperm call pattern gives the reference's exact index slices."""
    g = golden("sampler")
    buf = make_buffer(6, 8)
    torch.manual_seed(1234)
    ref = g.t("flat_indices")
    # slices are views of one index tensor that later epochs overwrite in place (as in the reference): compare on the fly
    for row, (meta, idx) in enumerate(C.MiniBatchSampler(num_epochs=3, num_mini_batches=4).indices(buf)):
        assert torch.equal(idx, ref[row])
        assert [meta["epoch_index"], meta["mini_batch_index"], meta["total_epochs"], meta["total_mini_batches"],
                int(meta["temporal"])] == g.np("flat_meta")[row].tolist()
    torch.manual_seed(1234)
    assert row == ref.shape[0] - 1
    for row, (meta, idx) in enumerate(C.TemporalMiniBatchSampler(num_epochs=3, num_mini_batches=2).indices(buf)):
        assert torch.equal(idx, g.t("temporal_indices")[row])
        assert meta["temporal"] is True


def test_sampler_coverage_and_remainder():
    buf = make_buffer(5, 7)  # 35 samples, 4 minibatches -> size 8, 3 dropped (mini_batch_sampler.py:66)
    for meta, idx in C.MiniBatchSampler(num_epochs=1, num_mini_batches=4).indices(buf):
        assert idx.numel() == 8
    seen = torch.cat([idx.clone() for _, idx in C.MiniBatchSampler(num_epochs=1, num_mini_batches=5).indices(buf)])
    assert sorted(seen.tolist()) == list(range(35))


def test_auto_sampler_picks_temporal_on_memory_fields():
    buf = make_buffer()
    assert not next(iter(C.AutoMiniBatchSampler().indices(buf)))[0]["temporal"]
    buf2 = C.Buffer(2, 4, device="cpu")
    for _ in range(2):
        buf2.push({"observation": torch.zeros(4, 3), "actor_memory": {"hidden": torch.zeros(4, 5)}})
    assert next(iter(C.AutoMiniBatchSampler().indices(buf2)))[0]["temporal"]


def test_hook_names_and_validation():
    assert C.GeneralizedAdvantageEstimation().name == "generalized_advantage_estimation"
    assert C.PpoSurrogateLoss().name == "ppo_surrogate_loss"
    for kwargs in ({"gamma": -0.1}, {"gamma": 1.0}, {"lamda": -0.1}, {"lamda": 1.1}, {"lamda_value": 1.1}):
        with pytest.raises(ValueError):
            C.GeneralizedAdvantageEstimation(**kwargs)
    with pytest.raises(ValueError):
        C.PpoSurrogateLoss(clip_ratio=0.0)
    with pytest.raises(ValueError):
        C.ValueLoss(weight=0.0)
    with pytest.raises(ValueError):
        C.EntropyLoss(weight=-1.0)
    hook = C.GeneralizedAdvantageEstimation()
    hook.update_attribute("gamma", 0.9)
    assert hook.gamma == 0.9
    with pytest.raises(ValueError, match="not mutable"):
        hook.update_attribute("recompute", True)


def test_factory_hook_order_and_registration():
    factory = C.anymal_c_rough_ppo(device="cpu").to_underlying()
    names = [h.name for h in factory.hooks]
    assert names == ["module_initialization", "value_computation", "generalized_advantage_estimation",
                     "advantage_normalization", "value_loss", "on_policy_preparation", "ppo_surrogate_loss",
                     "entropy_loss", "gradient_clipping", "on_policy_statistics", "adaptive_lr_schedule"]

    class Probe(C.Hook):
        pass

    factory.register_hook(Probe(), before="value_computation")
    assert factory.hooks[1].name == "probe" and factory.hooks.probe is factory.hooks[1]
    with pytest.raises(ValueError):
        factory.register_hook(Probe().name_("p2"), before="a", after="b")
    with pytest.raises(ValueError, match="No hook named"):
        factory.get_hook("missing")


def test_agent_construction_parameter_names_and_arena():
    spec = C.EnvironmentSpec(8, 235, 12, autoreset=True, final_state_is_missing=True)
    agent = C.anymal_c_rough_ppo(device="cpu")(spec)
    names = [n for n, _ in agent.named_parameters()]
    assert "actor.backbone.layers.4.weight" in names and "actor.distribution.std.param" in names
    assert "critic.value_head.bias" in names
    assert sum(p.numel() for p in agent.parameters()) == 571801  # reference count
    flat = agent.optimizer.flat_param
    p = dict(agent.named_parameters())["critic.value_head.weight"]
    assert p.data_ptr() >= flat.data_ptr() and p.grad is not None and p.grad.shape == p.shape
    # value_computation disabled bootstrapping because the env omits final states (value.py:38-40)
    assert agent.hook["value_computation"].bootstrap_truncated_states is False
    with pytest.raises(ValueError, match="compile"):
        C.anymal_c_rough_ppo(device="cpu", compile=True)(spec)
    with pytest.raises(ValueError, match="autocast"):
        C.anymal_c_rough_ppo(device="cpu", autocast=True)(spec)


def test_adaptive_lr_schedule_host_logic():
    hook = C.AdaptiveLRSchedule(0.015)
    # accumulated log error crosses +1 after a few large-KL iterations -> LR shrinks by exp(-clip(avg)*0.2)
    scales = [hook._compute_scale(0.03) for _ in range(2)]
    assert scales[0] is None and scales[1] == pytest.approx(torch.exp(torch.tensor(-0.2 * 0.6931471805599453)).item())
    assert hook._accumulated_log_error == 0.0 and hook._count == 0


def test_optimizer_state_dict_is_torch_adam_shaped_and_round_trips():
    """Checkpoint layout (reference template/agent.py:283-330 stores `optimizer.state_dict()` of torch.optim.Adam)."""
    spec = C.EnvironmentSpec(8, 235, 12, autoreset=True, final_state_is_missing=True)
    agent = C.anymal_c_rough_ppo(device="cpu")(spec)
    opt = agent.optimizer
    assert opt.state_dict()["state"] == {}                       # like torch: no state before the first step
    g = torch.Generator().manual_seed(0)
    opt.exp_avg.copy_(torch.randn(opt.exp_avg.shape, generator=g))
    opt.exp_avg_sq.copy_(torch.rand(opt.exp_avg_sq.shape, generator=g))
    opt.step_count = 7
    opt.param_groups[0]["lr"] = 5e-4
    sd = opt.state_dict()
    names = sd["param_groups"][0]["param_names"]
    assert names == [n for n, _ in agent.named_parameters()] and sd["param_groups"][0]["params"] == list(range(len(names)))
    assert sd["state"][0]["exp_avg"].shape == (512, 235) and float(sd["state"][3]["step"]) == 7.0
    # a stock torch.optim.Adam over equally shaped parameters accepts it
    stock = torch.optim.Adam([torch.nn.Parameter(torch.zeros_like(p)) for p in agent.parameters()], lr=1e-3)
    stock.load_state_dict({"state": sd["state"], "param_groups": [{k: v for k, v in sd["param_groups"][0].items() if k != "param_names"}]})
    assert stock.param_groups[0]["lr"] == 5e-4
    # and a second agent restores bit-identical moments, matching by name even when the checkpoint order differs
    other = C.anymal_c_rough_ppo(device="cpu")(spec).optimizer
    perm = list(reversed(range(len(names))))
    shuffled = {"state": {j: sd["state"][i] for j, i in enumerate(perm)},
                "param_groups": [{**sd["param_groups"][0], "param_names": [names[i] for i in perm], "params": list(range(len(names)))}]}
    other.load_state_dict(shuffled)
    assert other.step_count == 7 and other.param_groups[0]["lr"] == 5e-4
    for p, off in zip(other.arena.params, other.arena.offsets):
        n = p.numel()
        assert torch.equal(other.exp_avg[off:off + n], opt.exp_avg[off:off + n])
        assert torch.equal(other.exp_avg_sq[off:off + n], opt.exp_avg_sq[off:off + n])
    with pytest.raises(ValueError, match="parameters"):
        other.load_state_dict({"state": {}, "param_groups": [{"params": [0, 1], "param_names": ["a", "b"]}]})


def test_metrics_deferred_mode_collects_then_merges_like_eager():
    """template/graphs.py records metrics in deferred mode while capturing and merges them after every replay: the
    result must be the running mean the eager path produces (reference utils/metrics.py:11-96 semantics)."""
    from cusrl_b200.metrics import Metrics

    eager, deferred = Metrics(), Metrics()
    static = {"loss": torch.zeros(()), "ratio": torch.zeros(())}   # what a graph's static outputs look like
    deferred.begin_deferred()
    deferred.record(loss=static["loss"])
    deferred.record_mean("ratio", static["ratio"], 128)
    entries = deferred.end_deferred()
    assert len(deferred) == 0 and [(n, c) for n, _, c in entries] == [("loss", 1), ("ratio", 128)]
    for step in range(4):
        static["loss"].fill_(float(step))
        static["ratio"].fill_(1.0 + step)
        # record() took value.mean() at "capture" time: for a 0-dim tensor that is a new tensor, so refresh it the way a
        # replay would (the captured mean kernel rewrites the same output)
        entries[0][1].copy_(static["loss"])
        deferred.apply(entries)
        eager.record(loss=static["loss"].clone())
        eager.record_mean("ratio", static["ratio"].clone(), 128)
    assert deferred.summary() == eager.summary() == {"loss": 1.5, "ratio": 2.5}
    deferred.record(extra=torch.ones(3))                           # back to immediate merging
    assert deferred.summary()["extra"] == 1.0


def test_flat_adam_device_scalar_bookkeeping():
    spec = C.EnvironmentSpec(8, 235, 12, autoreset=True, final_state_is_missing=True)
    opt = C.anymal_c_rough_ppo(device="cpu")(spec).optimizer
    assert opt.step_dev is None
    opt.step_count = 5
    opt.use_device_scalars()
    assert int(opt.step_dev) == 5 and float(opt.lr_dev) == pytest.approx(1e-3)
    opt.param_groups[0]["lr"] = 2.5e-4                              # what AdaptiveLRSchedule does between iterations
    opt.sync_device_scalars()
    assert float(opt.lr_dev) == pytest.approx(2.5e-4) and opt._lr_uploaded == 2.5e-4
    sd = opt.state_dict()
    opt.step_count = 9
    opt.load_state_dict(sd)                                          # step 5 has no per-parameter state only if step == 0
    assert opt.step_count == 5 and int(opt.step_dev) == 5


def test_graph_runner_control_flow_with_mocked_capture():
    """The CUDA-graph train-step runner driven end to end on the CPU (kernels stubbed, torch.cuda.graph mocked): warm-up
    steps, one capture, replays, optimizer / metric bookkeeping.  Runs in a subprocess because the harness patches the
    binding."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "tools" / "host_overhead_cpu.py"), "--envs", "64", "--iters", "1", "--graphs"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "graph runner: 1 capture(s), 58 replays, optimizer step_count 60" in res.stdout, res.stdout


def test_gradient_clipping_groups_host_logic(monkeypatch):
    """Reference cusrl_test/hook/on_policy/test_gradient_clipping.py:29-50 on the flat arena: parameters are grouped by
    the longest matching name prefix, each group is clipped to its own limit, the pre-clip norm is recorded as
    `grad_norm/<prefix or default>`.  The three kernels the hook launches are replaced by their one-line definitions
    (include/cusrl_b200.h, K9) so the HOST logic -- prefix matching, arena ranges, record keys -- runs on the CPU; the
    kernels themselves are covered by tests/test_kernels_gpu.py."""
    from types import SimpleNamespace

    from cusrl_b200 import ops
    from cusrl_b200.metrics import Metrics
    from cusrl_b200.template.optimizer import FlatAdam

    def grad_sumsq_(grad, sumsq):
        return sumsq.add_(grad.double().square().sum())

    def clip_coef(sumsq, max_norm, norm, coef):
        norm.copy_(sumsq.sqrt().float())
        coef.copy_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))

    monkeypatch.setattr(ops, "grad_sumsq_", grad_sumsq_)
    monkeypatch.setattr(ops, "clip_coef", clip_coef)
    monkeypatch.setattr(ops, "scale_", lambda x, scale: x.mul_(scale))

    model = torch.nn.ModuleDict({"actor": torch.nn.Linear(2, 2), "critic": torch.nn.Linear(2, 1),
                                 "actor_extra": torch.nn.Linear(3, 1)})
    optimizer = FlatAdam(model.named_parameters(), lr=0.1)
    hook = C.GradientClipping(max_grad_norm=0.5, actor=0.25)
    hook.agent = SimpleNamespace(metrics=Metrics())
    optimizer.flat_grad.fill_(1.0)
    hook.pre_optim(optimizer)

    def norm(prefixes):
        return torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters() if n.split(".")[0] in prefixes]).norm().item()

    assert norm({"actor"}) == pytest.approx(0.25, abs=1e-5)                 # 6 ones: sqrt(6) clipped to 0.25
    assert norm({"critic", "actor_extra"}) == pytest.approx(0.5, abs=1e-5)  # "actor_extra" is NOT under prefix "actor"
    summary = hook.agent.metrics.summary()
    assert set(summary) == {"grad_norm/actor", "grad_norm/default"}
    assert summary["grad_norm/actor"] == pytest.approx(6 ** 0.5, rel=1e-6)
    assert summary["grad_norm/default"] == pytest.approx(7 ** 0.5, rel=1e-6)
    # longest prefix wins; unlimited groups are left alone
    hook = C.GradientClipping(max_grad_norm=None, **{"actor": 1.0, "actor.bias": None})
    assert list(hook.groups) == ["actor.bias", "actor"] and hook._match_prefix("actor.bias") == "actor.bias"
    assert hook._match_prefix("actor.weight") == "actor" and hook._match_prefix("actors.weight") == ""
    with pytest.raises(ValueError, match="'max_grad_norm' must be non-negative"):
        C.GradientClipping(max_grad_norm=-1.0)


def test_metric_is_exact_for_one_record_and_a_weighted_mean_for_many():
    """Reference utils/metrics.py:11-40 semantics: mean of the recorded means weighted by element count."""
    from cusrl_b200.metrics import Metric, Metrics

    m = Metric()
    assert m.count == 0 and m.mean.numel() == 0
    x = torch.tensor(0.1234567)
    m.update(x, 1_572_864)
    assert torch.equal(m.mean, x)                       # a metric recorded once per update is kept bit-exactly
    x.add_(1.0)
    assert float(m.mean) == pytest.approx(0.1234567)    # ... and does not alias the recorded tensor
    g = torch.Generator().manual_seed(0)
    values, counts = torch.randn(20, generator=g), [393216] * 19 + [17]
    m = Metric()
    for v, c in zip(values, counts):
        m.update(v, c)
    expected = float((values.double() * torch.tensor(counts, dtype=torch.float64)).sum() / sum(counts))
    assert m.count == sum(counts) and float(m.mean) == pytest.approx(expected, rel=1e-6)
    m.update(torch.tensor(5.0), 0)                      # empty records are ignored
    assert m.count == sum(counts)
    metrics = Metrics()
    metrics.record(a=torch.tensor([1.0, 3.0]), b=2.0, c=None, d=torch.tensor([]))
    metrics.record(a=torch.tensor([5.0]))
    assert metrics.summary("Agent") == {"Agent/a": 3.0, "Agent/b": 2.0}


def test_detached_gradients_are_rebound_to_the_arena():
    """`module.zero_grad()` (set_to_none) or a hook's `p.grad = None` detaches a parameter from the flat
    gradient arena; the gradients autograd then allocates must be folded back in, not silently dropped."""
    from cusrl_b200.template.optimizer import ParamArena

    lin = torch.nn.Linear(5, 3)
    arena = ParamArena(lin.named_parameters())
    assert arena.rebind_gradients() == 0
    lin.zero_grad()
    assert lin.weight.grad is None
    lin(torch.randn(2, 5)).sum().backward()
    fresh = lin.weight.grad.clone()
    assert lin.weight.grad.data_ptr() != arena.flat_grad.data_ptr()
    assert arena.rebind_gradients() == 2
    assert torch.equal(lin.weight.grad, fresh) and lin.weight.grad.data_ptr() == arena.flat_grad.data_ptr()
    assert torch.equal(arena.flat_grad[: fresh.numel()].view_as(fresh), fresh)


def test_trainer_refuses_environments_without_autoreset():
    import cusrl_b200 as C

    class Env:
        spec = C.EnvironmentSpec(4, 3, 2, autoreset=False)
        num_instances = 4

    class Agent:
        device = torch.device("cpu")

    with pytest.raises(ValueError, match="autoreset"):
        C.Trainer(Env(), Agent())


def test_sampler_may_skip_the_fp32_input_gather_only_when_provably_unused():
    """ActorCritic._configure_sampler: `sampler.pair_only` is set for the plain PPO preset (every hook one of this package's
    PPO hooks, MLP networks the f16x3 path covers) and stays empty as soon as anything else could read the fp32 minibatch
    observation: a user hook, observation normalisation, symmetry hooks, recurrent networks."""
    import cusrl_b200 as C

    spec = C.EnvironmentSpec(8, 19, 4, autoreset=True, final_state_is_missing=True)
    plain = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), device="cpu")(spec)
    assert plain.sampler.pair_only == frozenset(("observation", "state"))

    class Peek(C.Hook):
        pass

    factory = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), device="cpu").to_underlying()
    factory.register_hook(Peek())
    assert factory(spec).sampler.pair_only == frozenset()
    normalised = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), normalize_observation=True, device="cpu")
    assert normalised(spec).sampler.pair_only == frozenset()
    recurrent = C.RecurrentPpoAgentFactory(actor_hidden_size=64, critic_hidden_size=64, device="cpu")(spec)
    assert recurrent.sampler.pair_only == frozenset()
    sym = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), device="cpu").to_underlying()
    sym.register_hook(C.MirrorSymmetryLoss(0.1), after="ppo_surrogate_loss")
    sym_spec = C.EnvironmentSpec(8, 19, 4, autoreset=True, final_state_is_missing=True,
                                 mirror_observation=C.MirrorDef(list(range(19)), []), mirror_action=C.MirrorDef(list(range(4)), []))
    assert sym(sym_spec).sampler.pair_only == frozenset()


def test_adaptive_lr_schedule_rolls_back_an_update_that_exceeds_max_kl():
    """The reference's rollback test (cusrl_test/hook/on_policy/test_lr_schedule.py:38-63) against this implementation of
    the KL-adaptive schedule: an update whose KL exceeds `max_kl_divergence` is undone (parameters AND optimizer state back
    to the pre-update checkpoint) while the learning-rate scale decided from that KL is kept."""
    import math

    spec = C.EnvironmentSpec(8, 19, 4, autoreset=True, final_state_is_missing=True)
    factory = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), desired_kl_divergence=None,
                                device="cpu").to_underlying()
    factory.register_hook(C.AdaptiveLRSchedule(0.01, max_kl_divergence=0.02, scale_factor=0.2))
    agent = factory(spec)
    hook = agent.hook["adaptive_lr_schedule"]
    actor_param = next(agent.actor.parameters())
    original = actor_param.detach().clone()
    hook.pre_update(agent.buffer)
    with torch.no_grad():
        actor_param.add_(1.0)
    agent.metrics.clear()
    agent.record(kl_divergence=0.03)
    hook.post_update()
    assert torch.allclose(actor_param, original)
    assert hook._lr_scale == pytest.approx(math.exp(-0.2))          # log(0.03 / 0.01) > threshold 1 -> clip to 1 -> exp(-0.2)
    for base_lr, group in zip(hook._base_lrs, agent.optimizer.param_groups):
        scaled = any(n.startswith("actor.") for n in group["param_names"])
        assert group["lr"] == pytest.approx(base_lr * (hook._lr_scale if scaled else 1.0))
    assert agent.metrics["update_rejected"].mean.item() == pytest.approx(1.0)
    # an update inside the limit is kept
    hook.pre_update(agent.buffer)
    with torch.no_grad():
        actor_param.add_(0.5)
    agent.metrics.clear()
    agent.record(kl_divergence=0.012)
    hook.post_update()
    assert torch.allclose(actor_param, original + 0.5)
    assert agent.metrics["update_rejected"].mean.item() == pytest.approx(0.0)


def test_advantage_reduction_weighted_sum_mean_and_weight_updates():
    """The reference's own tests (cusrl_test/hook/on_policy/test_advantage.py:9-34) against this AdvantageReduction: the
    vector advantage of a multi-term reward is reduced to the scalar the surrogate needs."""
    from types import SimpleNamespace

    hook = C.AdvantageReduction(reduction="sum", weight=(1.0, 2.0))
    hook.agent = SimpleNamespace(to_tensor=lambda value: torch.as_tensor(value, dtype=torch.float32))
    hook.init()
    batch = {"advantage": torch.tensor([[1.0, 2.0], [3.0, 4.0]])}
    assert hook.objective({}, batch) is None
    assert torch.allclose(batch["advantage"], torch.tensor([[5.0], [11.0]]))
    hook.update_attribute("weight", (0.5, 0.5))
    batch = {"advantage": torch.tensor([[2.0, 6.0]])}
    hook.objective({}, batch)
    assert torch.allclose(batch["advantage"], torch.tensor([[4.0]]))
    hook.update_attribute("weight", None)
    batch = {"advantage": torch.tensor([[2.0, 6.0]])}
    hook.objective({}, batch)
    assert torch.allclose(batch["advantage"], torch.tensor([[8.0]]))

    mean = C.AdvantageReduction(reduction="mean")
    mean.agent = hook.agent
    mean.init()
    batch = {"advantage": torch.tensor([[1.0, 3.0]])}
    mean.objective({}, batch)
    assert torch.allclose(batch["advantage"], torch.tensor([[2.0]]))
    with pytest.raises(ValueError, match="Unsupported reduction"):
        C.AdvantageReduction(reduction="max")
    assert mean.name == "advantage_reduction" and mean.training_only


def test_module_initialization_rules_like_the_reference_tests():
    """cusrl_test/hook/control/test_module_initialization.py:19-45 against this hook: orthogonal rows with the requested
    gains, zero biases, the distribution head's own gain, and recurrent layers without biases."""
    import math
    from types import SimpleNamespace

    from torch import nn

    class Actor(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = nn.Linear(3, 2)
            self.distribution = nn.Module()
            self.distribution.mean_head = nn.Linear(2, 1)

    actor = Actor()
    critic = nn.Sequential(nn.Linear(3, 2), nn.ReLU(), nn.Linear(2, 1))
    hook = C.ModuleInitialization(scale=math.sqrt(2), scale_dist=0.1, zero_bias=True)
    hook.agent = SimpleNamespace(actor=actor, critic=critic)
    hook.init()
    assert torch.allclose(actor.backbone.bias, torch.zeros_like(actor.backbone.bias))
    assert torch.allclose(actor.distribution.mean_head.bias, torch.zeros_like(actor.distribution.mean_head.bias))
    assert torch.allclose(critic[0].bias, torch.zeros_like(critic[0].bias))
    assert actor.distribution.mean_head.weight.norm().item() == pytest.approx(0.1)
    assert critic[0].weight[0].norm().item() == pytest.approx(math.sqrt(2))
    for rnn_cls, gates in ((nn.RNN, 1), (nn.GRU, 3), (nn.LSTM, 4)):
        module = rnn_cls(input_size=3, hidden_size=4, num_layers=2, bias=False)
        C.ModuleInitialization()._init_module(module, scale=1.0, zero_bias=True)
        for layer in range(module.num_layers):
            w = getattr(module, f"weight_hh_l{layer}")
            assert w.shape == (4 * gates, 4) and not hasattr(module, f"bias_hh_l{layer}")
            assert torch.allclose(w.T @ w, torch.eye(4), atol=1e-5)           # orthonormal columns, gain 1


def test_ppo_hook_suite_options_like_the_reference_tests():
    """cusrl_test/preset/test_ppo.py:16-36: optional hooks appear exactly when switched on, in the reference's order."""
    hooks = C.ppo_hook_suite(normalize_observation=True, desired_kl_divergence=0.01, max_kl_divergence=0.02, empty_cuda_cache=True)
    kinds = [type(h) for h in hooks]
    assert C.ObservationNormalization in kinds and C.AdaptiveLRSchedule in kinds and C.EmptyCudaCache in kinds
    assert kinds[-1] is C.EmptyCudaCache and kinds[1] is C.ObservationNormalization and all(h is not None for h in hooks)
    assert hooks[-1].name == "empty_cuda_cache"
    hooks[-1].post_update()                                   # a no-op without a GPU, must not raise
    kinds = {type(h) for h in C.ppo_hook_suite(normalize_observation=False, desired_kl_divergence=None, empty_cuda_cache=False)}
    assert not kinds & {C.ObservationNormalization, C.AdaptiveLRSchedule, C.EmptyCudaCache}


def test_metrics_like_the_reference_tests():
    """cusrl_test/utils/test_metrics.py:7-38: count-weighted means, prefixed summary keys, empty / None values ignored,
    conversion errors name the metric."""
    from cusrl_b200.metrics import Metrics

    metrics = Metrics()
    metrics.record({"loss": torch.tensor([1.0, 3.0])}, accuracy=0.5, ignored=None)
    metrics.record(loss=torch.tensor([5.0, 7.0, 9.0]))
    assert len(metrics) == 2 and metrics["loss"].count == 5
    assert metrics.summary("train") == pytest.approx({"train/loss": 5.0, "train/accuracy": 0.5})
    metrics = Metrics()
    metrics.record(empty=torch.tensor([]), missing=None)
    assert len(metrics) == 0
    metrics.record(value=[1.0, 2.0])
    metrics.clear()
    assert list(metrics.items()) == []
    with pytest.raises(ValueError, match="bad_metric"):
        metrics.record(bad_metric=object())


def test_small_api_mirrors_resize_buffer_metrics_values_recurrent_empty_cache():
    """Members of the reference surface that are one-liners but must exist: ``ActorCritic.resize_buffer``
    (actor_critic.py:327-330), ``Metrics.values`` (utils/metrics.py:53-54), ``RecurrentPpoAgentFactory.empty_cuda_cache``
    (preset/ppo.py:241-243; off by default here, see preset.py) and the export callbacks on every hook (template/hook.py:344-356)."""
    from cusrl_b200.metrics import Metrics

    spec = C.EnvironmentSpec(8, 19, 4, autoreset=True, final_state_is_missing=True)
    agent = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), device="cpu")(spec)
    agent.buffer.push({"observation": torch.zeros(8, 19)})
    version = agent.buffer.layout_version
    agent.resize_buffer(agent.buffer_capacity)                       # same capacity: nothing happens
    assert agent.buffer.layout_version == version and "observation" in agent.buffer
    agent.resize_buffer(12)
    assert agent.buffer_capacity == 12 and agent.buffer.capacity == 12 and "observation" not in agent.buffer
    assert agent.buffer.layout_version > version
    metrics = Metrics()
    metrics.record(a=1.0, b=2.0)
    assert [m.count for m in metrics.values()] == [1, 1]
    names = [h.name for h in C.RecurrentPpoAgentFactory(device="cpu", empty_cuda_cache=True).to_underlying().hooks]
    assert names[-1] == "empty_cuda_cache"
    assert "empty_cuda_cache" not in [h.name for h in C.RecurrentPpoAgentFactory(device="cpu").to_underlying().hooks]
    calls = []

    class Probe(C.Hook):
        def pre_export(self, graph):
            calls.append(("pre", graph))

        def post_export(self, graph):
            calls.append(("post", graph))

    composite = C.HookComposite([Probe().active_(False), C.EmptyCudaCache()])      # inactive hooks are visited too
    composite.pre_export("g")
    composite.post_export("g")
    assert calls == [("pre", "g"), ("post", "g")]


def test_environment_spec_accepts_the_reference_call_styles():
    """cusrl/template/environment.py:118-176 and cusrl_test/template/test_hook.py:39-43: keyword construction with the
    reference's field names and defaults, the reference's two positional arguments, extra keywords / attributes, ``get``;
    and this package's original positional order, which every test and tool uses."""
    ref_style = C.EnvironmentSpec(35, 12, state_dim=42)
    assert (ref_style.observation_dim, ref_style.action_dim, ref_style.state_dim, ref_style.num_instances) == (35, 12, 42, 1)
    assert ref_style.reward_dim == 1 and ref_style.autoreset is False and ref_style.final_state_is_missing is False
    assert ref_style.device == torch.device("cpu") and ref_style.observation_stat_groups == () and ref_style.timestep is None
    assert ref_style.observation_normalization is None and ref_style.action_denormalization is None
    legacy = C.EnvironmentSpec(8, 19, 5, 11, 2, autoreset=True, final_state_is_missing=True)
    assert (legacy.num_instances, legacy.observation_dim, legacy.action_dim, legacy.state_dim, legacy.reward_dim) == (8, 19, 5, 11, 2)
    keyword = C.EnvironmentSpec(num_instances=4, observation_dim=3, action_dim=2, observation_stat_groups=[slice(0, 2)], foo=123)
    assert keyword.foo == 123 and keyword.get("foo") == 123 and keyword.get("bar", 7) == 7
    assert keyword.observation_stat_groups == (slice(0, 2),)
    keyword.baz = 5                                          # new attributes can be added later (EnvironmentSpecOverride)
    assert keyword.get("baz") == 5
    with pytest.raises(TypeError, match="observation_dim"):
        C.EnvironmentSpec(3)
    with pytest.raises(TypeError, match="multiple values"):
        C.EnvironmentSpec(3, 4, 5, observation_dim=2)


def test_actor_forward_type_router(monkeypatch):
    """The reference's actor routes on ``forward_type`` (nn/module/actor.py:70-99); deployment code (Player, export,
    InferenceWrapper) calls ``actor(obs, forward_type="act_deterministic")``.  Dispatch only: the kernels need a GPU."""
    actor = C.PpoAgentFactory(actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128), device="cpu")(
        C.EnvironmentSpec(4, 19, 5)).actor
    calls = []
    monkeypatch.setattr(type(actor), "explore", lambda self, obs, memory=None, deterministic=False, **kw: calls.append(("explore", deterministic)))
    monkeypatch.setattr(type(actor), "act", lambda self, obs, memory=None, deterministic=False, **kw: calls.append(("act", deterministic)))
    obs = torch.zeros(4, 19)
    actor(obs, forward_type="explore")
    actor(obs, forward_type="explore", deterministic=True)
    actor(obs, forward_type="act")
    actor(obs, forward_type="act_deterministic")
    assert calls == [("explore", False), ("explore", True), ("act", False), ("act", True)]
    with pytest.raises(ValueError, match="Unsupported 'forward_type'"):
        actor(obs, forward_type="sample")
    with pytest.raises(RuntimeError, match="CUDA"):        # the default route reaches the kernels, which refuse CPU tensors
        actor(obs)


def test_normal_dist_interface_like_the_reference_tests(monkeypatch):
    """cusrl_test/nn/module/test_distribution.py:7-24 against this NormalDist (identity bijector): shapes of parameters /
    sample / log-prob / entropy, KL of a distribution with itself, the deterministic wrapper; and the closed forms against
    torch.distributions.  The mean head is a kernel (CUDA only): it is replaced by torch's linear for this CPU test."""
    import cusrl_b200.nn.functional as F

    monkeypatch.setattr(F, "linear_head", lambda x, w, b: torch.nn.functional.linear(x, w, b))
    torch.manual_seed(0)
    dist = C.NormalDist(input_dim=4, output_dim=2, init_std=0.7)
    latent = torch.randn(3, 4)
    params = dist(latent)
    sample, logp = dist.sample_from_dist(params)
    assert params["mean"].shape == (3, 2) and params["std"].shape == (3, 2) and sample.shape == (3, 2) and logp.shape == (3, 1)
    assert torch.allclose(dist.deterministic()(latent), params["mean"]) and torch.allclose(dist.determine(latent), params["mean"])
    params2, (sample2, logp2) = dist.sample(latent)
    assert torch.equal(params2["mean"], params["mean"]) and sample2.shape == (3, 2) and logp2.shape == (3, 1)
    normal = torch.distributions.Normal(params["mean"], params["std"])
    assert torch.allclose(dist.compute_logp(params, sample), normal.log_prob(sample).sum(-1, keepdim=True), atol=1e-6)
    assert torch.allclose(dist.compute_entropy(params), normal.entropy().sum(-1, keepdim=True), atol=1e-6)
    assert torch.allclose(dist.compute_kl_div(params, params), torch.zeros(3, 1), atol=1e-6)
    other = {"mean": params["mean"] + 0.3, "std": params["std"] * 1.5}
    kl = torch.distributions.kl_divergence(normal, torch.distributions.Normal(other["mean"], other["std"])).sum(-1, keepdim=True)
    assert torch.allclose(dist.compute_kl_div(params, other), kl, atol=1e-6)
    with pytest.raises(ValueError, match="identity bijector"):
        C.NormalDist.Factory(bijector="exp")(4, 2)
