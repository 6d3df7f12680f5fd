"""GPU parity tests (run with -m gpu on a B200): every kernel, called through the C ABI, against
(a) golden fixtures generated from the live reference and (b) the CPU oracle on seeded inputs,
plus size-independent properties at BASELINE.json's full size (65536 envs x 24 steps).

Tolerances: bit-exact for GAE / next_value / gather / permutation work; 1e-5 relative fp32 (the
north-star bound) for reductions and transcendental arithmetic."""

from __future__ import annotations

import pytest
import torch

from oracle import ppo_path as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from cusrl_b200 import build

    build.build()
    from cusrl_b200 import ops as _ops

    return _ops


def rel_close(a: torch.Tensor, b: torch.Tensor, rtol=1e-5, atol=1e-6):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    assert bool((err <= bound).all()), f"max err {err.max().item():.3e}, max rel {(err / (b.abs() + 1e-12)).max().item():.3e}"


# ------------------------------------------------------------------------------------------- K1/K3
@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_gae_golden_bit_exact(ops, golden, tag):
    g = golden("gae")
    gamma, lamda, lv = (float(x) for x in g.np(f"{tag}_hyper"))
    adv, ret = ops.gae(g.t(f"{tag}_reward", DEV), g.t(f"{tag}_done", DEV), g.t(f"{tag}_value", DEV),
                       g.t(f"{tag}_next_value", DEV), gamma, lamda, None if lv < 0 else lv)
    assert torch.equal(adv.cpu(), g.t(f"{tag}_advantage"))
    assert torch.equal(ret.cpu(), g.t(f"{tag}_return"))


def test_gae_known_answer(ops):
    # cusrl_test/hook/on_policy/test_gae.py:8-16
    adv, _ = ops.gae(torch.ones(3, 1, 1, device=DEV), torch.tensor([[[False]], [[True]], [[False]]], device=DEV),
                     torch.zeros(3, 1, 1, device=DEV), torch.zeros(3, 1, 1, device=DEV), 0.5, 1.0)
    assert adv.flatten().tolist() == [1.5, 1.0, 1.0]


def test_gae_validates_like_reference(ops):
    x = torch.zeros(2, 4, 1, device=DEV)
    d = torch.zeros(2, 4, 1, dtype=torch.bool, device=DEV)
    for kw in ({"gamma": -0.1}, {"gamma": 1.0}, {"lamda": -0.1}, {"lamda": 1.1}, {"lamda_value": 1.1}):
        args = {"gamma": 0.99, "lamda": 0.95, "lamda_value": None} | kw
        with pytest.raises(ValueError):
            ops.gae(x, d, x, x, **args)


@pytest.mark.parametrize("T,N,Dv", [(24, 4096, 1), (24, 4099, 1), (24, 1024, 3), (7, 33, 1), (100, 130, 2), (1, 5, 1)])
@pytest.mark.parametrize("lamda_value", [None, 0.7])
def test_gae_vs_oracle_bit_exact(ops, T, N, Dv, lamda_value):
    g = torch.Generator().manual_seed(T * 1000 + N + Dv)
    reward, value, nv = (torch.randn(T, N, Dv, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, generator=g) < 0.05
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, lamda_value)
    adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95, lamda_value)
    assert torch.equal(adv.cpu(), ref_adv)
    assert torch.equal(ret.cpu(), ref_ret)


@pytest.mark.parametrize("vec,threads", [(1, 128), (2, 64), (2, 32), (4, 128), (1, 32), (1, 96)])
def test_gae_all_vector_widths(ops, vec, threads):
    from cusrl_b200 import _lib

    lib = _lib.load()
    T, N = 24, 2048
    g = torch.Generator().manual_seed(5)
    reward, value, nv = (torch.randn(T, N, 1, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, generator=g) < 0.05
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)
    assert lib.cusrl_b200_gae_set_config(vec, threads) == 0
    try:
        adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95)
        assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)
    finally:
        lib.cusrl_b200_gae_set_config(1, 64)


@pytest.mark.parametrize("schedule", [0, 1])
@pytest.mark.parametrize("T", [8, 12, 16, 24, 32, 7, 33])
@pytest.mark.parametrize("N,Dv,threads", [(4096, 1, 64), (1001, 1, 32), (130, 3, 128)])
def test_gae_both_schedules_bit_exact(ops, schedule, T, N, Dv, threads):
    """Chunked kernel (run-time T) and exact-length kernel (T in {8,12,16,24,32}; others fall back): same bits as the
    oracle for both lamda settings, with and without the return, ragged N and vector rewards."""
    from cusrl_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(T * 131 + N + Dv)
    reward, value, nv = (torch.randn(T, N, Dv, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, generator=g) < 0.08
    assert lib.cusrl_b200_gae_set_schedule(schedule) == 0 and lib.cusrl_b200_gae_set_config(1, threads) == 0
    try:
        for lamda_value in (None, 0.6):
            ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, lamda_value)
            adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95, lamda_value)
            assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)
        adv_only, none = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95, compute_return=False)
        assert none is None and torch.equal(adv_only.cpu(), O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)[0])
    finally:
        lib.cusrl_b200_gae_set_schedule(ops.GAE_DEFAULT_SCHEDULE)
        lib.cusrl_b200_gae_set_config(1, 64)
    assert lib.cusrl_b200_gae_set_schedule(2) != 0


@pytest.mark.parametrize("tag", ["a", "b"])
def test_next_value_golden(ops, golden, tag):
    g = golden("next_value")
    nv = ops.next_value(g.t(f"{tag}_value", DEV), g.t(f"{tag}_terminated", DEV), g.t(f"{tag}_truncated", DEV),
                        g.t(f"{tag}_boot", DEV), float(g.np(f"{tag}_termination_value")))
    assert torch.equal(nv.cpu(), g.t(f"{tag}_next_value"))


@pytest.mark.parametrize("T,N,Dv", [(24, 4096, 1), (24, 1001, 1), (5, 64, 2)])
def test_gae_fused_equals_separate(ops, T, N, Dv):
    g = torch.Generator().manual_seed(N)
    reward, value = torch.randn(T, N, Dv, generator=g), torch.randn(T, N, Dv, generator=g)
    boot = torch.randn(N, Dv, generator=g)
    term = torch.rand(T, N, 1, generator=g) < 0.03
    trunc = (torch.rand(T, N, 1, generator=g) < 0.03) & ~term
    nv_ref = O.next_value_ref(value, term, trunc, boot, 0.0)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, term | trunc, value, nv_ref, 0.99, 0.95, None)
    nv_out = torch.empty(T, N, Dv, device=DEV)
    adv, ret = ops.gae_fused(reward.to(DEV), term.to(DEV), trunc.to(DEV), value.to(DEV), boot.to(DEV), 0.99, 0.95,
                             next_value_out=nv_out)
    assert torch.equal(nv_out.cpu(), nv_ref)
    assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)


@pytest.mark.parametrize("T,N", [(24, 65536), (24, 1001), (8, 77), (32, 4096), (16, 449), (12, 128)])
@pytest.mark.parametrize("lamda_value", [None, 0.9])
@pytest.mark.parametrize("threads", [128, 256, 448])
def test_gae_chain_equals_the_three_stages(ops, T, N, lamda_value, threads):
    """K3 + K1 + K2 statistics in one launch (ops.gae_chain) against the oracle's three separate stages: next_value,
    advantage and return BIT-exact (every CTA size), mean / unbiased variance to 1e-6 relative, and the normalised
    advantages that follow from them to the north-star 1e-5."""
    from cusrl_b200 import _lib

    g = torch.Generator().manual_seed(N + T)
    reward, value = torch.randn(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g)
    boot = torch.randn(N, 1, generator=g)
    term = torch.rand(T, N, 1, generator=g) < 0.03
    trunc = (torch.rand(T, N, 1, generator=g) < 0.03) & ~term
    nv_ref = O.next_value_ref(value, term, trunc, boot, 0.25)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, term | trunc, value, nv_ref, 0.99, 0.95, lamda_value)
    nv, adv, ret = (torch.full((T, N, 1), float("nan"), device=DEV) for _ in range(3))
    assert ops.gae_chain_supported(T, 1) and not ops.gae_chain_supported(T, 2) and not ops.gae_chain_supported(7, 1)
    try:
        assert _lib.load().cusrl_b200_gae_set_chain_threads(threads) == 0
        mean_var = ops.gae_chain(reward.to(DEV), term.to(DEV), trunc.to(DEV), value.to(DEV), boot.to(DEV), 0.99, 0.95,
                                 lamda_value, 0.25, nv, adv, ret)
    finally:
        _lib.load().cusrl_b200_gae_set_chain_threads(448)
    assert torch.equal(nv.cpu(), nv_ref)
    assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)
    var, mean = torch.var_mean(ref_adv.double(), dim=(0, 1))
    assert mean_var[0].item() == pytest.approx(mean.item(), rel=1e-6, abs=1e-7)
    assert mean_var[1].item() == pytest.approx(var.item(), rel=1e-6)
    ops.advantage_normalize_(adv, mean_var, 1e-8)
    rel_close(adv, O.normalize_advantage_ref(ref_adv), rtol=1e-5, atol=2e-6)


def test_pre_update_chain_fuses_across_adjacent_hooks_only():
    """ValueComputation defers to the GAE hook only when that hook runs NEXT; a hook in between gets the unfused stages
    (it may read buffer["next_value"]).  Both orders must give identical leaves."""
    from cusrl_b200 import build

    build.build()
    import cusrl_b200 as C

    class Peek(C.Hook):
        def pre_update(self, buffer):
            self.seen = buffer["next_value"].clone()

    results = []
    for with_peek in (False, True):
        torch.manual_seed(3)
        factory = C.anymal_c_rough_ppo(num_steps_per_update=8, actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128),
                                       device=DEV).to_underlying()
        peek = Peek()
        if with_peek:
            factory.register_hook(peek, after="value_computation")
        env = C.SyntheticEnvironment(300, 19, 5, device=DEV, seed=4)
        agent = factory.from_environment(env)
        obs, _, _ = env.reset()
        for _ in range(8):
            agent.act(obs)
            obs, _, reward, term, trunc, _ = env.step(None)
            agent.step(obs, reward, term, trunc)
        launches = {}
        real_chain, real_next_value = C.ops.gae_chain, C.ops.next_value
        C.ops.gae_chain = lambda *a, **k: (launches.__setitem__("chain", 1), real_chain(*a, **k))[1]
        C.ops.next_value = lambda *a, **k: (launches.__setitem__("separate", 1), real_next_value(*a, **k))[1]
        try:
            agent.hook.pre_update(agent.buffer)
        finally:
            C.ops.gae_chain, C.ops.next_value = real_chain, real_next_value
        assert launches == ({"separate": 1} if with_peek else {"chain": 1})
        if with_peek:
            assert torch.equal(peek.seen, agent.buffer["next_value"])
        results.append({k: agent.buffer[k].clone() for k in ("next_value", "advantage", "return")})
    for k in results[0]:
        if k == "advantage":   # normalised with statistics reduced in a different order: equal to fp32 rounding
            assert torch.allclose(results[0][k], results[1][k], rtol=1e-6, atol=1e-6)
        else:
            assert torch.equal(results[0][k], results[1][k]), k


def test_gae_full_size_properties(ops):
    """65536 x 24 (BASELINE.json): checked through size-independent properties + a sampled oracle."""
    T, N = 24, 65536
    g = torch.Generator(device=DEV).manual_seed(0)
    reward, value, nv = (torch.randn(T, N, 1, device=DEV, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, device=DEV, generator=g) < 0.011
    adv, ret = ops.gae(reward, done, value, nv, 0.99, 0.95)
    # return - value == advantage (gae.py:99-101), last step is the one-step TD error (gae.py:17)
    assert torch.equal(ret, value + adv)
    assert torch.equal(adv[-1], reward[-1] + nv[-1] * 0.99 - value[-1])
    # a done at step t cuts the recursion: adv[t] is the plain delta there
    delta = reward + nv * 0.99 - value
    assert torch.equal(adv[:-1][done[:-1]], delta[:-1][done[:-1]])
    # linearity in (reward, value, next_value) for fixed done: scaling inputs by 2 scales outputs by 2 exactly
    adv2, _ = ops.gae(reward * 2, done, value * 2, nv * 2, 0.99, 0.95)
    assert torch.equal(adv2, adv * 2)
    # sampled columns against the oracle, bit-exact
    cols = torch.randint(0, N, (512,), generator=torch.Generator().manual_seed(1))
    ref_adv, ref_ret = O.advantage_and_return_ref(reward[:, cols].cpu(), done[:, cols].cpu(), value[:, cols].cpu(),
                                                  nv[:, cols].cpu(), 0.99, 0.95)
    assert torch.equal(adv[:, cols].cpu(), ref_adv) and torch.equal(ret[:, cols].cpu(), ref_ret)


# ---------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("tag", ["a", "b"])
def test_advnorm_golden(ops, golden, tag):
    g = golden("advnorm")
    adv = g.t(f"{tag}_in", DEV).clone()
    ops.advantage_normalize_(adv, ops.advantage_stats(adv))
    rel_close(adv, g.t(f"{tag}_out"), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("shape", [(24, 4096, 1), (24, 65536, 1), (24, 1023, 1), (8, 300, 3), (3, 1, 1)])
def test_advnorm_vs_oracle(ops, shape):
    g = torch.Generator().manual_seed(sum(shape))
    adv = torch.randn(*shape, generator=g) * 2.5 + 0.7
    ref = O.normalize_advantage_ref(adv)
    dev = adv.to(DEV)
    mv = ops.advantage_stats(dev)
    var, mean = torch.var_mean(adv, dim=(0, 1))
    rel_close(mv[: shape[-1]], mean, rtol=1e-5, atol=1e-6)
    rel_close(mv[shape[-1]:], var, rtol=1e-5, atol=1e-7)
    ops.advantage_normalize_(dev, mv)
    rel_close(dev, ref, rtol=1e-5, atol=2e-6)
    # post-normalisation mean ~ 0 per channel (cusrl_test/hook/on_policy/test_advantage.py:37-48)
    assert torch.allclose(dev.mean(dim=(0, 1)).cpu(), torch.zeros(shape[-1]), atol=1e-5)


def test_merge_mean_var_golden(ops, golden):
    g = golden("advnorm")
    gathered = torch.cat([g.t("merge_means"), g.t("merge_vars")], dim=1).to(DEV)
    out = ops.merge_mean_var(gathered)
    rel_close(out[:2], g.t("merge_mean"))
    rel_close(out[2:], g.t("merge_var"))


# ---------------------------------------------------------------------------------------------- K8
def test_gather_matches_reference_sampler(ops, golden):
    g = golden("sampler")
    obs = g.t("obs", DEV)                       # [T, N, 3]
    flat = obs.flatten(0, 1)
    for row, idx in enumerate(g.t("flat_indices", DEV)):
        dst = torch.empty(idx.numel(), 3, device=DEV)
        ops.gather_rows([(flat, dst)], idx)
        assert torch.equal(dst.cpu(), g.t("flat_obs")[row])


@pytest.mark.parametrize("E,B", [(98304, 24576), (1000, 333)])
def test_gather_multi_field_padded_bit_exact(ops, E, B):
    g = torch.Generator().manual_seed(E)
    obs_store = torch.zeros(E, 240, device=DEV)             # padded storage, public view [E, 235]
    obs_store[:, :235] = torch.randn(E, 235, generator=g).to(DEV)
    obs = obs_store[:, :235]
    action = torch.randn(E, 12, generator=g).to(DEV)
    logp = torch.randn(E, 1, generator=g).to(DEV)
    done = (torch.rand(E, 1, generator=g) < 0.3).to(DEV)
    idx = torch.randperm(E, generator=g)[:B].to(DEV)
    d_obs = torch.full((B, 240), 7.0, device=DEV)
    d_action, d_logp = torch.empty(B, 12, device=DEV), torch.empty(B, 1, device=DEV)
    d_done = torch.empty(B, 1, dtype=torch.bool, device=DEV)
    ops.gather_rows([(obs, d_obs), (action, d_action), (logp, d_logp), (done, d_done)], idx)
    assert torch.equal(d_obs[:, :235], obs[idx])
    assert torch.equal(d_obs[:, 235:], torch.zeros(B, 5, device=DEV))   # padding written as zeros
    assert torch.equal(d_action, action[idx]) and torch.equal(d_logp, logp[idx]) and torch.equal(d_done, done[idx])


# ---------------------------------------------------------------------------------------------- K4
def _run_loss(ops, g, tag):
    clip, w_s, w_e, w_v, vclip = (float(x) for x in g.np(f"{tag}_hyper"))
    return ops.ppo_loss(
        g.t(f"{tag}_mean", DEV), g.t(f"{tag}_std_param", DEV), g.t(f"{tag}_action", DEV), g.t(f"{tag}_logp_old", DEV),
        g.t(f"{tag}_advantage", DEV), g.t(f"{tag}_return", DEV), g.t(f"{tag}_value_old", DEV),
        g.t(f"{tag}_curr_value", DEV), clip, w_s, w_e, w_v, None if vclip < 0 else vclip)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_ppo_loss_golden(ops, golden, tag):
    g = golden("objective")
    out = _run_loss(ops, g, tag)
    rel_close(out["logp"], g.t(f"{tag}_logp"))
    rel_close(out["entropy"], g.t(f"{tag}_entropy"))
    rel_close(out["logp_ratio"], g.t(f"{tag}_logp_ratio"), atol=1e-5)
    rel_close(out["prob_ratio"], g.t(f"{tag}_prob_ratio"), rtol=2e-5)
    rel_close(out["losses"], g.t(f"{tag}_losses"))
    rel_close(out["d_mean"], g.t(f"{tag}_d_mean"), rtol=2e-5, atol=1e-8)
    rel_close(out["d_std_surr"] + out["d_std_ent"], g.t(f"{tag}_d_std"), rtol=2e-5, atol=1e-7)
    rel_close(out["d_value"], g.t(f"{tag}_d_value"), rtol=1e-5, atol=1e-9)


def test_ppo_loss_known_answer(ops):
    # cusrl_test/hook/on_policy/test_ppo.py:8-14: A=[1,-2], ratio=[1.5,0.5], clip 0.2 -> 0.2
    ratio = torch.tensor([[1.5], [0.5]])
    mean = torch.zeros(2, 1, device=DEV)
    std = torch.ones(1, device=DEV)
    action = torch.zeros(2, 1, device=DEV)
    logp_now = -O.LOG_SQRT_2PI
    logp_old = (logp_now - ratio.log()).to(DEV)
    out = ops.ppo_loss(mean, std, action, logp_old, torch.tensor([[1.0], [-2.0]], device=DEV), None, None, None,
                       0.2, 1.0, 0.5, 0.5)
    assert out["losses"][1].item() == pytest.approx(0.2, rel=1e-6)
    # entropy loss = -mean(entropy) * weight (test_ppo.py:28-32)
    ent = 0.5 + O.LOG_SQRT_2PI
    assert out["losses"][2].item() == pytest.approx(-ent * 0.5, rel=1e-6)
    assert out["losses"][0].item() == 0.0


@pytest.mark.parametrize("B,A,vclip", [(24576, 12, None), (98304, 12, 0.2), (4097, 7, None), (130, 32, None)])
def test_ppo_loss_vs_oracle(ops, B, A, vclip):
    g = torch.Generator().manual_seed(B + A)
    mean = torch.randn(B, A, generator=g)
    std = torch.rand(A, generator=g) * 0.8 + 0.4
    action = mean + torch.randn(B, A, generator=g) * std
    logp_old = O.normal_log_prob_ref(mean, std.repeat(B, 1), action) + torch.randn(B, 1, generator=g) * 0.2
    adv = torch.randn(B, 1, generator=g)
    ret = torch.randn(B, 1, generator=g)
    v_old = ret + torch.randn(B, 1, generator=g) * 0.5
    v = v_old + torch.randn(B, 1, generator=g) * 0.3
    ref = O.ppo_objective_ref(mean, std, action, logp_old, adv, ret, v_old, v, 0.2, 1.0, 0.005, 0.5, vclip)
    out = ops.ppo_loss(mean.to(DEV), std.to(DEV), action.to(DEV), logp_old.to(DEV), adv.to(DEV), ret.to(DEV),
                       v_old.to(DEV), v.to(DEV), 0.2, 1.0, 0.005, 0.5, vclip)
    rel_close(out["losses"], torch.stack([ref.value_loss, ref.surrogate_loss, ref.entropy_loss]), rtol=1e-5, atol=1e-7)
    rel_close(out["logp"], ref.logp, rtol=1e-5, atol=1e-5)
    rel_close(out["prob_ratio"], ref.prob_ratio, rtol=3e-5)
    rel_close(out["d_mean"], ref.d_mean, rtol=3e-5, atol=1e-9)
    rel_close(out["d_std_surr"], ref.d_std_surr, rtol=1e-4, atol=1e-6)
    rel_close(out["d_std_ent"], ref.d_std_ent, rtol=1e-5)
    rel_close(out["d_value"], ref.d_value, rtol=1e-5, atol=1e-10)
    rel_close(out["metrics"][0], ref.logp_ratio.abs().mean(), rtol=1e-5)
    rel_close(out["metrics"][2], v.sum(-1).mean(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("tag", ["a", "c"])
def test_policy_stats_golden(ops, golden, tag):
    g = golden("objective")
    B = g.t(f"{tag}_mean").shape[0]
    out = ops.policy_stats(g.t(f"{tag}_mean_old", DEV), g.t(f"{tag}_std_old", DEV), g.t(f"{tag}_mean", DEV),
                           g.t(f"{tag}_std_param", DEV), g.t(f"{tag}_action", DEV), g.t(f"{tag}_logp_old", DEV),
                           g.t(f"{tag}_advantage", DEV))
    assert B > 0
    rel_close(out, g.t(f"{tag}_stats"), rtol=2e-5, atol=1e-7)


def test_scale(ops):
    x = torch.randn(1000, device=DEV)
    ref = x * 0.25
    ops.scale_(x, torch.tensor([0.25], device=DEV))
    assert torch.equal(x, ref)


# ---------------------------------------------------------------------------------------------- K9
@pytest.mark.parametrize("n", [571801, 1000, 3])
def test_clip_and_adam_vs_torch(ops, n):
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * s for s in (3.0, 0.5, 0.01)]
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p = p0.clone().to(DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sumsq = torch.zeros(1, dtype=torch.float64, device=DEV)
    norm, coef = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    for step, gr in enumerate(grads, start=1):
        ref_p.grad = gr.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        gd = gr.to(DEV)
        sumsq.zero_()
        ops.grad_sumsq_(gd, sumsq)
        ops.clip_coef(sumsq, 1.0, norm, coef)
        ops.adam_step_(p, gd, m, v, step, 1e-3, coef=coef)
        rel_close(norm, ref_norm.reshape(1), rtol=1e-5)  # torch's fp32 norm vs our fp64 accumulation
        rel_close(p, ref_p.detach(), rtol=1e-5, atol=1e-7)
    # oracle restatement agrees with torch too
    po, mo, vo = O.adam_step_ref(p0, O.clip_grad_norm_ref([grads[0]], 1.0)[1][0], torch.zeros(n), torch.zeros(n), 1, 1e-3)
    assert po.shape == p0.shape
