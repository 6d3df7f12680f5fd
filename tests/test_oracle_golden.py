"""CPU: pin oracle/ppo_path.py against fixtures generated from the LIVE reference
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import ppo_path as O

GAE_CASES = ["a", "b", "c", "d", "e"]


@pytest.mark.parametrize("tag", GAE_CASES)
def test_gae_matches_reference_bit_exact(golden, tag):
    g = golden("gae")
    gamma, lamda, lv = g.np(f"{tag}_hyper")
    adv, ret = O.advantage_and_return_ref(
        g.t(f"{tag}_reward"), g.t(f"{tag}_done"), g.t(f"{tag}_value"), g.t(f"{tag}_next_value"),
        float(gamma), float(lamda), None if lv < 0 else float(lv))
    assert torch.equal(adv, g.t(f"{tag}_advantage"))
    assert torch.equal(ret, g.t(f"{tag}_return"))


def test_gae_known_answer(golden):
    # cusrl_test/hook/on_policy/test_gae.py:8-16 -> [1.5, 1.0, 1.0]
    adv = O.gae_ref(torch.ones(3, 1, 1), torch.tensor([[[False]], [[True]], [[False]]]),
                    torch.zeros(3, 1, 1), torch.zeros(3, 1, 1), 0.5, 1.0)
    assert adv.flatten().tolist() == [1.5, 1.0, 1.0]
    assert torch.equal(adv, golden("gae").t("known_answer"))


def test_gae_hook_known_answer_lamda_value():
    # cusrl_test/hook/on_policy/test_gae.py:19-31 -> advantage [1.5, 1.0], return [1.5, 2.0]
    reward = torch.tensor([[[1.0]], [[2.0]]])
    done = torch.zeros(2, 1, 1, dtype=torch.bool)
    value = torch.tensor([[[0.5]], [[1.0]]])
    next_value = torch.tensor([[[1.0]], [[0.0]]])
    adv, ret = O.advantage_and_return_ref(reward, done, value, next_value, 0.5, 1.0, 0.0)
    assert torch.allclose(adv.flatten(), torch.tensor([1.5, 1.0]))
    assert torch.allclose(ret.flatten(), torch.tensor([1.5, 2.0]))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_advantage_normalisation(golden, tag):
    g = golden("advnorm")
    out = O.normalize_advantage_ref(g.t(f"{tag}_in"))
    ref = g.t(f"{tag}_out")
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    mask = ~torch.isnan(ref)
    assert torch.equal(out[mask], ref[mask])


def test_advantage_normalisation_zero_mean():
    # cusrl_test/hook/on_policy/test_advantage.py:37-48
    adv = torch.randn(8, 16, 2, generator=torch.Generator().manual_seed(0)) * 4 + 3
    out = O.normalize_advantage_ref(adv)
    assert torch.allclose(out.mean(dim=(0, 1)), torch.zeros(2), atol=1e-5)


def test_merge_mean_var(golden):
    g = golden("advnorm")
    mean, var = O.merge_mean_var_ref(g.t("merge_means"), g.t("merge_vars"))
    assert torch.equal(mean, g.t("merge_mean")) and torch.equal(var, g.t("merge_var"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_next_value(golden, tag):
    g = golden("next_value")
    nv = O.next_value_ref(g.t(f"{tag}_value"), g.t(f"{tag}_terminated"), g.t(f"{tag}_truncated"),
                          g.t(f"{tag}_boot"), float(g.np(f"{tag}_termination_value")))
    assert torch.equal(nv, g.t(f"{tag}_next_value"))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_objective(golden, tag):
    g = golden("objective")
    clip, w_s, w_e, w_v, vclip = g.np(f"{tag}_hyper")
    out = O.ppo_objective_ref(
        g.t(f"{tag}_mean"), g.t(f"{tag}_std_param"), g.t(f"{tag}_action"), g.t(f"{tag}_logp_old"),
        g.t(f"{tag}_advantage"), g.t(f"{tag}_return"), g.t(f"{tag}_value_old"), g.t(f"{tag}_curr_value"),
        float(clip), float(w_s), float(w_e), float(w_v), None if vclip < 0 else float(vclip))
    tol = dict(rtol=1e-6, atol=1e-7)
    assert torch.allclose(out.logp, g.t(f"{tag}_logp"), **tol)
    assert torch.allclose(out.entropy, g.t(f"{tag}_entropy"), **tol)
    assert torch.allclose(out.prob_ratio, g.t(f"{tag}_prob_ratio"), **tol)
    losses = torch.stack([out.value_loss, out.surrogate_loss, out.entropy_loss])
    assert torch.allclose(losses, g.t(f"{tag}_losses"), **tol)
    assert torch.allclose(out.d_mean, g.t(f"{tag}_d_mean"), **tol)
    assert torch.allclose(out.d_std, g.t(f"{tag}_d_std"), rtol=1e-5, atol=1e-7)
    assert torch.allclose(out.d_value, g.t(f"{tag}_d_value"), **tol)
    std = g.t(f"{tag}_std_param").repeat(g.t(f"{tag}_mean").shape[0], 1)
    kl, iwa, sd = O.policy_stats_ref(g.t(f"{tag}_mean_old"), g.t(f"{tag}_std_old"), g.t(f"{tag}_mean"), std,
                                     g.t(f"{tag}_action"), g.t(f"{tag}_logp_old"), g.t(f"{tag}_advantage"))
    assert torch.allclose(torch.stack([kl, iwa, sd]), g.t(f"{tag}_stats"), **tol)


def test_objective_known_answers():
    # cusrl_test/hook/on_policy/test_ppo.py:8-14 -> 0.2 ; :28-32 -> entropy loss -1.0 for entropy 2.0, weight 0.5
    loss = O.surrogate_loss_ref(torch.tensor([[1.0], [-2.0]]), torch.tensor([[1.5], [0.5]]), 0.2)
    assert loss.item() == pytest.approx(0.2)
    assert (-(torch.full((4, 1), 2.0)).mean() * 0.5).item() == pytest.approx(-1.0)


def test_sampler_slices(golden):
    g = golden("sampler")
    flat = g.t("flat_indices")          # [3 epochs * 4 minibatches, 12]
    obs = g.t("obs")
    for e in range(3):
        perm = flat[4 * e : 4 * e + 4].reshape(-1)
        assert sorted(perm.tolist()) == list(range(48))  # coverage: every sample once per epoch
        for k, idx in enumerate(O.minibatch_slices_ref(perm, 4)):
            assert torch.equal(idx, flat[4 * e + k])
            assert torch.equal(O.gather_ref(obs, idx), g.t("flat_obs")[4 * e + k])
    temporal = g.t("temporal_indices")  # [3 epochs * 2 minibatches, 4 env columns]
    for row, idx in enumerate(temporal):
        assert torch.equal(O.gather_ref(obs, idx, temporal=True), g.t("temporal_obs")[row])
    assert g.np("flat_meta")[5].tolist() == [1, 1, 3, 4, 0]
    assert g.np("temporal_meta")[3].tolist() == [1, 1, 3, 2, 1]


def _iteration_setup(golden):
    g = golden("iteration")
    params = {k[len("param0/"):]: g.t(k) for k in g.keys() if k.startswith("param0/")}
    buf = {k[len("buffer/"):]: g.t(k) for k in g.keys() if k.startswith("buffer/")}
    cfg = O.PpoConfig(obs_dim=19, act_dim=5, hidden=(64, 32, 128), num_steps=6)
    return g, cfg, params, buf


def test_rollout_matches_reference(golden):
    g, cfg, params, buf = _iteration_setup(golden)
    agent = O.OraclePpo(cfg, params)
    for t in range(cfg.num_steps):
        tr = agent.act(buf["observation"][t], g.t("noise")[t])
        assert torch.allclose(tr["action_dist.mean"], buf["action_dist.mean"][t], rtol=1e-6, atol=1e-6)
        assert torch.allclose(tr["action"], buf["action"][t], rtol=1e-5, atol=1e-5)
        assert torch.allclose(tr["action_logp"], buf["action_logp"][t], rtol=1e-5, atol=1e-5)
        assert torch.allclose(tr["value"], buf["value"][t], rtol=1e-6, atol=1e-6)


def test_full_iteration_matches_reference(golden):
    g, cfg, params, buf = _iteration_setup(golden)
    agent = O.OraclePpo(cfg, params)
    perms = list(g.t("perms"))
    logs = agent.update(buf, perms)
    assert torch.equal(buf["next_value"], g.t("post/next_value"))
    assert torch.allclose(buf["return"], g.t("post/return"), rtol=0, atol=0)
    assert torch.allclose(buf["advantage"], g.t("post/advantage"), rtol=1e-6, atol=1e-6)
    ref_losses = g.np("minibatch_losses")
    got = np.array([[d["value_loss"], d["surrogate_loss"], d["entropy_loss"]] for d in logs])
    assert got.shape == ref_losses.shape == (20, 3)
    np.testing.assert_allclose(got, ref_losses, rtol=2e-5, atol=1e-6)
    for k in g.keys():
        if k.startswith("param1/"):
            name = k[len("param1/"):]
            assert torch.allclose(agent.params[name].detach(), g.t(k), rtol=1e-4, atol=2e-6), name
    names = g.np("metric_names").tolist()
    vals = dict(zip(names, g.np("metric_values").tolist()))
    assert agent.metrics["kl_divergence"] == pytest.approx(vals["Agent/kl_divergence"], rel=1e-3, abs=1e-7)
    assert agent.metrics["action_std"] == pytest.approx(vals["Agent/action_std"], rel=1e-5)
    assert cfg.lr * agent.lr_scale == pytest.approx(float(g.np("lr_after")), rel=1e-9)


def test_full_iteration_at_config2_shape_matches_reference(golden):
    """The oracle at BASELINE.json config 2 (4096 envs x 24 steps, obs 235, act 12, MLP 512-256-128) against the live
    reference's own iteration (golden `iteration_anymal.npz`; inputs regenerated from the shared seeded recipes)."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    import recipes as R

    g = golden("iteration_anymal")
    N, T = (int(v) for v in g.np("shape"))
    cfg = O.PpoConfig()
    names = g.np("param_names").tolist()
    params = {}
    shapes = {k: tuple(v.shape) for k, v in O.init_mlp_params_ref(cfg.obs_dim, cfg.act_dim, cfg.hidden).items()}
    for name in names:
        params[name] = torch.ones(shapes[name]) if name.endswith("std.param") else R.seeded_parameter(name, shapes[name], seed=1)
    torch.set_num_threads(8)
    agent = O.OraclePpo(cfg, params)
    stream, noise = R.anymal_stream(T, N, seed=2), R.noise_stream(T, N, R.ACT, seed=3)
    leaves = {k: [] for k in ("observation", "action", "action_logp", "action_dist.mean", "action_dist.std", "value")}
    for t in range(T):
        tr = agent.act(stream["obs"][t], noise[t])
        for k in leaves:
            leaves[k].append(tr[k])
    buf = {k: torch.stack(v) for k, v in leaves.items()}
    buf.update(next_observation=stream["obs"][1:], reward=stream["reward"].clone(), terminated=stream["terminated"],
               truncated=stream["truncated"], done=stream["terminated"] | stream["truncated"])

    def check(prefix, tensor, rtol, atol):
        fp = R.fingerprint(tensor)
        assert torch.allclose(fp["sample"], g.t(f"{prefix}/sample"), rtol=rtol, atol=atol), prefix
        ref_abs = float(g.np(f"{prefix}/abs_sum")[0])
        assert abs(float(fp["abs_sum"]) - ref_abs) <= 10 * rtol * ref_abs + atol, prefix

    for key in ("action", "action_logp", "action_dist.mean", "value"):
        check(f"rollout/{key}", buf[key], 1e-5, 1e-5)
    seeded = R.SeededRandperm(seed=4)
    perms = [seeded(T * N) for _ in range(cfg.epochs)]
    logs = agent.update(buf, perms)
    for key in ("next_value", "return", "advantage"):
        check(f"post/{key}", buf[key], 1e-5, 1e-5)
    got = np.array([[d["value_loss"], d["surrogate_loss"], d["entropy_loss"]] for d in logs])
    np.testing.assert_allclose(got, g.np("minibatch_losses"), rtol=5e-5, atol=2e-6)
    vals = dict(zip(g.np("metric_names").tolist(), g.np("metric_values").tolist()))
    assert agent.metrics["kl_divergence"] == pytest.approx(vals["Agent/kl_divergence"], rel=2e-3, abs=1e-7)
    assert agent.metrics["action_std"] == pytest.approx(vals["Agent/action_std"], rel=1e-5)
    assert cfg.lr * agent.lr_scale == pytest.approx(float(g.np("lr_after")), rel=1e-9)
    torch.set_num_threads(1)
