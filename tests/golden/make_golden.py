"""Generate the golden fixtures in this directory from the LIVE reference (chengruiz/cusrl).

Run in the build container only (the reference is mounted read-only at /root/reference and does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own hooks / sampler / agent (with two tiny stand-in modules for the
missing pure-Python deps ``gymnasium`` and ``objprint``, see ``_shims/``), feeds them seeded inputs and
stores inputs + outputs as ``.npz``.  The fixtures pin ``oracle/ppo_path.py`` (tests/test_oracle_golden.py)
and are what the ``-m gpu`` parity tests compare the CUDA path against.  Nothing here is product code.
"""

from __future__ import annotations

import os
import sys
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
REFERENCE = Path(os.environ.get("CUSRL_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE / "_shims"))
sys.path.insert(0, str(REFERENCE))

import contextlib  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cusrl  # noqa: E402
from cusrl.hook.on_policy.advantage import AdvantageNormalization  # noqa: E402
from cusrl.hook.on_policy.gae import GeneralizedAdvantageEstimation, _generalized_advantage_estimation  # noqa: E402
from cusrl.hook.on_policy.ppo import _ppo_surrogate_loss  # noqa: E402
from cusrl.hook.on_policy.value import ValueComputation, _clipped_value_loss  # noqa: E402
from cusrl.nn.module.distribution import NormalDist  # noqa: E402
from cusrl.sampler.mini_batch_sampler import MiniBatchSampler, TemporalMiniBatchSampler  # noqa: E402
from cusrl.template.buffer import Buffer  # noqa: E402

torch.set_num_threads(1)


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name: str, **arrays):
    np.savez_compressed(HERE / f"{name}.npz", **{k: npy(v) for k, v in arrays.items()})
    print(f"wrote {name}.npz: " + ", ".join(f"{k}{tuple(npy(v).shape)}" for k, v in arrays.items()))


# ------------------------------------------------------------------------------------------------
def make_gae():
    cases = {}
    g = torch.Generator().manual_seed(0)
    for tag, (T, N, Dv, p_done, gamma, lamda, lamda_value) in {
        "a": (24, 64, 1, 0.05, 0.99, 0.95, None),
        "b": (24, 64, 1, 0.05, 0.99, 0.95, 0.9),
        "c": (5, 7, 3, 0.3, 0.9, 0.8, None),
        "d": (1, 4, 1, 0.5, 0.99, 0.95, 0.5),
        "e": (33, 10, 2, 0.1, 0.995, 1.0, 0.0),
    }.items():
        reward = torch.randn(T, N, Dv, generator=g)
        value = torch.randn(T, N, Dv, generator=g)
        next_value = torch.randn(T, N, Dv, generator=g)
        done = torch.rand(T, N, 1, generator=g) < p_done
        hook = GeneralizedAdvantageEstimation(gamma=gamma, lamda=lamda, lamda_value=lamda_value)
        data = {"reward": reward, "done": done, "value": value, "next_value": next_value}
        hook._compute_advantage_and_return(data)
        cases.update({
            f"{tag}_reward": reward, f"{tag}_value": value, f"{tag}_next_value": next_value, f"{tag}_done": done,
            f"{tag}_hyper": np.array([gamma, lamda, -1.0 if lamda_value is None else lamda_value], dtype=np.float64),
            f"{tag}_advantage": data["advantage"], f"{tag}_return": data["return"],
        })
    # the reference's own known-answer case (cusrl_test/hook/on_policy/test_gae.py:8-16)
    ka = _generalized_advantage_estimation(
        reward=torch.ones(3, 1, 1), done=torch.tensor([[[False]], [[True]], [[False]]]),
        value=torch.zeros(3, 1, 1), next_value=torch.zeros(3, 1, 1), gamma=0.5, lamda=1.0)
    cases["known_answer"] = ka
    save("gae", **cases)


def make_advnorm():
    g = torch.Generator().manual_seed(1)
    out = {}
    hook = AdvantageNormalization()
    for tag, shape in {"a": (24, 64, 1), "b": (6, 10, 3), "c": (2, 1, 1)}.items():
        adv = torch.randn(*shape, generator=g) * 3.0 + 1.5
        out[f"{tag}_in"] = adv.clone()
        hook.normalize_(adv)
        out[f"{tag}_out"] = adv
    # cross-rank merge arithmetic of utils/distributed.py:175-183, evaluated on stacked per-rank stats
    means = torch.randn(4, 2, generator=g)
    variances = torch.rand(4, 2, generator=g) + 0.5
    mean = torch.mean(means, dim=0)
    var = torch.mean(variances + (means - mean).square(), dim=0)
    out.update(merge_means=means, merge_vars=variances, merge_mean=mean, merge_var=var)
    save("advnorm", **out)


def make_next_value():
    g = torch.Generator().manual_seed(2)
    out = {}
    for tag, (T, N, Dv) in {"a": (24, 32, 1), "b": (4, 6, 2)}.items():
        value = torch.randn(T, N, Dv, generator=g)
        boot = torch.randn(N, Dv, generator=g)
        terminated = torch.rand(T, N, 1, generator=g) < 0.1
        truncated = (torch.rand(T, N, 1, generator=g) < 0.1) & ~terminated
        next_obs = torch.randn(T, N, 3, generator=g)
        buffer = Buffer(T, N, device="cpu")
        buffer["value"] = value
        buffer["terminated"] = terminated
        buffer["truncated"] = truncated
        buffer["next_observation"] = next_obs
        hook = ValueComputation(termination_value=0.25 if tag == "b" else 0.0)
        critic = SimpleNamespace(evaluate=lambda state, memory=None: boot)
        hook.agent = SimpleNamespace(
            critic=critic, autocast=contextlib.nullcontext,
            environment_spec=SimpleNamespace(final_state_is_missing=True))
        hook.init()  # -> bootstrap_truncated_states = False (IsaacLab-style env, value.py:38-40)
        hook.pre_update(buffer)
        out.update({f"{tag}_value": value, f"{tag}_boot": boot, f"{tag}_terminated": terminated,
                    f"{tag}_truncated": truncated, f"{tag}_termination_value": np.float32(hook.termination_value),
                    f"{tag}_next_value": buffer["next_value"]})
    save("next_value", **out)


def make_objective():
    g = torch.Generator().manual_seed(3)
    out = {}
    for tag, (B, A, Dv, value_clip) in {"a": (256, 12, 1, None), "b": (100, 12, 1, 0.2), "c": (37, 5, 2, None)}.items():
        dist = NormalDist(8, A)
        with torch.no_grad():
            dist.std.param.copy_(torch.rand(A, generator=g) * 0.8 + 0.4)
        mean = torch.randn(B, A, generator=g).requires_grad_(True)
        action = (mean.detach() + torch.randn(B, A, generator=g) * 0.7)
        std = dist.std(mean)  # StddevVector.forward: param.repeat(B, 1)
        params = {"mean": mean, "std": std}
        logp = dist.compute_logp(params, action)
        entropy = dist.compute_entropy(params)
        logp_old = (logp.detach() + torch.randn(B, 1, generator=g) * 0.3)
        advantage = torch.randn(B, 1, generator=g)
        if Dv != 1:
            advantage = advantage  # surrogate requires [B,1]; value dims are independent
        ret = torch.randn(B, Dv, generator=g)
        value_old = ret + torch.randn(B, Dv, generator=g) * 0.5
        curr_value = (value_old + torch.randn(B, Dv, generator=g) * 0.3).requires_grad_(True)
        logp_ratio = logp - logp_old
        prob_ratio = logp_ratio.exp()
        w_v, w_s, w_e, clip = 0.5, 1.0, 0.005, 0.2
        l_v = (torch.nn.functional.mse_loss(ret, curr_value) if value_clip is None
               else _clipped_value_loss(value_old, curr_value, ret, value_clip)) * w_v
        l_s = _ppo_surrogate_loss(advantage, prob_ratio, clip) * w_s
        l_e = -entropy.mean() * w_e
        loss = sum({"value_loss": l_v, "surrogate_loss": l_s, "entropy_loss": l_e}.values())
        loss.backward()
        out.update({
            f"{tag}_mean": mean, f"{tag}_std_param": dist.std.param, f"{tag}_action": action,
            f"{tag}_logp_old": logp_old, f"{tag}_advantage": advantage, f"{tag}_return": ret,
            f"{tag}_value_old": value_old, f"{tag}_curr_value": curr_value,
            f"{tag}_hyper": np.array([clip, w_s, w_e, w_v, -1.0 if value_clip is None else value_clip]),
            f"{tag}_logp": logp, f"{tag}_entropy": entropy, f"{tag}_logp_ratio": logp_ratio,
            f"{tag}_prob_ratio": prob_ratio,
            f"{tag}_losses": torch.stack([l_v, l_s, l_e]), f"{tag}_total": loss,
            f"{tag}_d_mean": mean.grad, f"{tag}_d_std": dist.std.param.grad, f"{tag}_d_value": curr_value.grad,
        })
        # KL / statistics (hook/on_policy/stats.py:29-40)
        mean_old = mean.detach() + torch.randn(B, A, generator=g) * 0.1
        std_old = std.detach() * (1.0 + torch.rand(B, A, generator=g) * 0.1)
        kl = dist.compute_kl_div({"mean": mean_old, "std": std_old}, {"mean": mean.detach(), "std": std.detach()})
        iwa = advantage * (dist.compute_logp({"mean": mean.detach(), "std": std.detach()}, action) - logp_old).exp()
        out.update({f"{tag}_mean_old": mean_old, f"{tag}_std_old": std_old,
                    f"{tag}_stats": torch.stack([kl.mean(), iwa.mean(), std.detach().mean()])})
    # reference known-answer tests (cusrl_test/hook/on_policy/test_ppo.py:8-14)
    out["known_surrogate"] = _ppo_surrogate_loss(torch.tensor([[1.0], [-2.0]]), torch.tensor([[1.5], [0.5]]), 0.2)
    save("objective", **out)


def make_sampler():
    out = {}
    T, N = 6, 8
    buffer = Buffer(T, N, device="cpu")
    buffer["observation"] = torch.arange(T * N * 3, dtype=torch.float32).reshape(T, N, 3)
    buffer["flag"] = (torch.arange(T * N).reshape(T, N, 1) % 3) == 0
    buffer.full = True
    for cls, tag in ((MiniBatchSampler, "flat"), (TemporalMiniBatchSampler, "temporal")):
        torch.manual_seed(1234)
        sampler = cls(num_epochs=3, num_mini_batches=4 if tag == "flat" else 2)
        idx_rows, obs_rows, meta_rows = [], [], []
        # capture the index slices through the public _sample hook
        captured = []
        orig = sampler._sample

        def spy(name, data, indices, _orig=orig):
            if name == "observation":
                captured.append(indices.clone())
            return _orig(name, data, indices)

        sampler._sample = spy
        for meta, batch in sampler(buffer):
            obs_rows.append(batch["observation"].clone())
            meta_rows.append([meta["epoch_index"], meta["mini_batch_index"], meta["total_epochs"],
                              meta["total_mini_batches"], int(meta["temporal"])])
        out[f"{tag}_indices"] = torch.stack(captured)
        out[f"{tag}_obs"] = torch.stack(obs_rows)
        out[f"{tag}_meta"] = np.array(meta_rows)
    out["obs"] = buffer["observation"]
    save("sampler", **out)


def make_iteration():
    """A whole reference PPO iteration (rollout + update) on CPU at a small size."""
    from cusrl.template.environment import EnvironmentSpec

    obs_dim, act_dim, N, T = 19, 5, 16, 6
    hidden = (64, 32, 128)  # latent 128: the B200 head kernels need a latent width that is a multiple of 128
    torch.manual_seed(7)
    factory = cusrl.preset.ppo.PpoAgentFactory(
        num_steps_per_update=T, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation_fn="ELU", lr=1e-3,
        sampler_epochs=5, sampler_mini_batches=4, orthogonal_init=False, entropy_loss_weight=0.005,
        desired_kl_divergence=0.015, device="cpu")
    spec = EnvironmentSpec(num_instances=N, observation_dim=obs_dim, action_dim=act_dim, reward_dim=1,
                           autoreset=True, final_state_is_missing=True)
    agent = factory(spec)
    init_state = {net: {k: v.clone() for k, v in agent.state_dict()[net].items()} for net in ("actor", "critic")}
    out = {}
    for net in ("actor", "critic"):
        for k, v in init_state[net].items():
            out[f"param0/{net}.{k}"] = v
    g = torch.Generator().manual_seed(11)
    obs = torch.randn(N, obs_dim, generator=g)
    noises = []
    for t in range(T):
        # make action sampling reproducible outside the reference: capture the noise rsample draws
        torch.manual_seed(1000 + t)
        action = agent.act(obs)
        noises.append((action - agent.transition["action_dist"]["mean"]) / agent.transition["action_dist"]["std"])
        next_obs = torch.randn(N, obs_dim, generator=g)
        reward = torch.randn(N, 1, generator=g)
        terminated = torch.rand(N, 1, generator=g) < 0.15
        truncated = (torch.rand(N, 1, generator=g) < 0.1) & ~terminated
        ready = agent.step(next_obs, reward, terminated, truncated)
        obs = next_obs
    assert ready
    for key, leaf in agent.buffer.storage.items():
        out[f"buffer/{key}"] = leaf.clone()
    out["noise"] = torch.stack(noises)

    perms, minibatch_logs = [], []
    real_randperm = torch.randperm

    def spy_randperm(*args, **kwargs):
        r = real_randperm(*args, **kwargs)
        perms.append(r.clone())
        return r

    real_record = agent.record

    def spy_record(metrics=None, /, **kwargs):
        if "surrogate_loss" in kwargs:
            minibatch_logs.append([float(kwargs["value_loss"]), float(kwargs["surrogate_loss"]), float(kwargs["entropy_loss"])])
        return real_record(metrics, **kwargs)

    torch.manual_seed(99)
    torch.randperm = spy_randperm
    agent.record = spy_record
    try:
        metrics = agent.update()
    finally:
        torch.randperm = real_randperm
    out["perms"] = torch.stack(perms[:5])          # 5 training epochs (mini_batch_sampler.py:56,68)
    out["stats_perm"] = perms[5]                   # OnPolicyStatistics' sampler (stats.py:32)
    out["minibatch_losses"] = np.array(minibatch_logs, dtype=np.float64)
    for key in ("advantage", "return", "next_value"):
        out[f"post/{key}"] = agent.buffer.storage[key].clone()
    final_state = agent.state_dict()
    for net in ("actor", "critic"):
        for k, v in final_state[net].items():
            out[f"param1/{net}.{k}"] = v
    out["metric_names"] = np.array(sorted(metrics))
    out["metric_values"] = np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64)
    out["lr_after"] = np.float64(agent.optimizer.param_groups[0]["lr"])
    save("iteration", **out)


def make_rnd():
    """RandomNetworkDistillation.pre_update / objective of the reference (hook/auxiliary/rnd.py:68-81)."""
    from cusrl.hook.auxiliary.rnd import RandomNetworkDistillation

    torch.manual_seed(21)
    T, N, D = 6, 16, 19
    hook = RandomNetworkDistillation(cusrl.Mlp.Factory([64, 64]), output_dim=16, reward_scale=0.1)
    recorded = {}
    hook.agent = SimpleNamespace(state_dim=D, setup_module=lambda m: m, record=lambda **kw: recorded.update(kw))
    hook.init()
    g = torch.Generator().manual_seed(22)
    buffer = Buffer(T, N, device="cpu")
    buffer["next_observation"] = torch.randn(T, N, D, generator=g)
    reward0 = torch.randn(T, N, 1, generator=g)
    buffer["reward"] = reward0.clone()
    hook.pre_update(buffer)
    out = {"next_observation": buffer["next_observation"], "reward_before": reward0, "reward_after": buffer["reward"],
           "rnd_reward": recorded["rnd_reward"]}
    for net in ("target", "predictor"):
        for k, v in getattr(hook, net).state_dict().items():
            out[f"{net}/{k}"] = v
    batch = {"next_observation": buffer["next_observation"].flatten(0, 1)[:40]}
    loss = hook.objective({}, batch)["rnd_loss"]
    loss.backward()
    out["rnd_loss"] = loss
    for k, p in hook.predictor.named_parameters():
        out[f"grad/{k}"] = p.grad
    save("rnd", **out)


def make_lstm():
    """Reference Rnn (LSTM) forward on a done-segmented sequence + gradients (nn/module/rnn.py:264-299)."""
    torch.manual_seed(31)
    T, N, I, H, L = 7, 12, 19, 32, 2
    rnn = cusrl.Rnn.Factory("LSTM", hidden_size=H, num_layers=L)(I)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(T, N, I, generator=g)
    done = torch.rand(T, N, 1, generator=g) < 0.2
    memory = {"hidden": torch.randn(N, L * H, generator=g) * 0.5, "cell": torch.randn(N, L * H, generator=g) * 0.5}
    out, _ = rnn(x, memory={k: v.clone() for k, v in memory.items()}, done=done)
    gout = torch.randn(T, N, H, generator=g)
    (out * gout).sum().backward()
    res = {"x": x, "done": done, "hidden0": memory["hidden"], "cell0": memory["cell"], "out": out, "gout": gout}
    for k, p_ in rnn.named_parameters():
        res[f"param/{k}"] = p_.detach()
        res[f"grad/{k}"] = p_.grad
    # single-step (rollout) call without done: returns the next memory
    step_out, step_mem = rnn(x[0], memory={k: v.clone() for k, v in memory.items()}, sequential=False)
    res.update(step_out=step_out, step_hidden=step_mem["hidden"], step_cell=step_mem["cell"])
    save("lstm", **res)


# ------------------------------------------------------------------------------------------------
# Fixtures at BASELINE.json's shapes: inputs are regenerated from seeds on both sides (recipes.py); only what the
# reference computed from them is stored (scalar series, checksums, strided samples).
def make_iteration_anymal(N: int = 4096, T: int = 24):
    """A whole reference PPO iteration (24 x act/step + update) at BASELINE.json config 2: 4096 envs x 24 steps, obs 235,
    act 12, MLP 512-256-128 ELU, the Anymal-C-rough preset values (zoo/isaaclab/locomotion.py:48-59), on CPU."""
    import recipes as R
    from cusrl.template.environment import EnvironmentSpec
    import torch.distributions.normal as normal_mod

    factory = cusrl.preset.ppo.PpoAgentFactory(
        num_steps_per_update=T, actor_hidden_dims=(512, 256, 128), critic_hidden_dims=(512, 256, 128),
        activation_fn="ELU", lr=1e-3, sampler_epochs=5, sampler_mini_batches=4, orthogonal_init=False,
        entropy_loss_weight=0.005, desired_kl_divergence=0.015, device="cpu")
    spec = EnvironmentSpec(num_instances=N, observation_dim=R.OBS, action_dim=R.ACT, reward_dim=1, autoreset=True,
                           final_state_is_missing=True)
    torch.set_num_threads(8)
    agent = factory(spec)
    R.set_seeded_parameters(agent.named_parameters(), seed=1)
    stream = R.anymal_stream(T, N, seed=2)
    noise = R.noise_stream(T, N, R.ACT, seed=3)
    step = {"t": 0}
    real_normal = normal_mod._standard_normal
    normal_mod._standard_normal = lambda shape, dtype, device: noise[step["t"]].to(dtype=dtype, device=device).reshape(shape)
    try:
        for t in range(T):
            step["t"] = t
            agent.act(stream["obs"][t])
            ready = agent.step(stream["obs"][t + 1], stream["reward"][t], stream["terminated"][t], stream["truncated"][t])
    finally:
        normal_mod._standard_normal = real_normal
    assert ready
    out = {"shape": np.array([N, T])}
    for key in ("action", "action_logp", "action_dist.mean", "value"):
        out.update(R.flatten_fingerprints(f"rollout/{key}", R.fingerprint(agent.buffer.storage[key])))

    minibatch_logs = []
    real_record = agent.record

    def spy_record(metrics=None, /, **kwargs):
        if "surrogate_loss" in kwargs:
            minibatch_logs.append([float(kwargs["value_loss"]), float(kwargs["surrogate_loss"]), float(kwargs["entropy_loss"])])
        return real_record(metrics, **kwargs)

    real_randperm = torch.randperm
    seeded = R.SeededRandperm(seed=4)
    torch.randperm = seeded
    agent.record = spy_record
    try:
        metrics = agent.update()
    finally:
        torch.randperm = real_randperm
    assert seeded.calls == 6
    out["minibatch_losses"] = np.array(minibatch_logs, dtype=np.float64)
    for key in ("advantage", "return", "next_value"):
        out.update(R.flatten_fingerprints(f"post/{key}", R.fingerprint(agent.buffer.storage[key])))
    names = []
    for name, p in agent.named_parameters():
        names.append(name)
        out.update(R.flatten_fingerprints(f"param1/{name}", R.fingerprint(p)))
    out["param_names"] = np.array(names)
    out["metric_names"] = np.array(sorted(metrics))
    out["metric_values"] = np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64)
    out["lr_after"] = np.float64(agent.optimizer.param_groups[0]["lr"])
    torch.set_num_threads(1)
    save("iteration_anymal", **out)


def make_lstm_anymal(T: int = 24, N: int = 1024, I: int = 235, H: int = 256, L: int = 2):
    """Reference Rnn (LSTM 2 x 256 on 235 inputs, the config-3 backbone) on one temporal minibatch [24, 1024, 235]."""
    import recipes as R

    torch.set_num_threads(8)
    rnn = cusrl.Rnn.Factory("LSTM", hidden_size=H, num_layers=L)(I)
    R.set_seeded_parameters(rnn.named_parameters(), seed=5)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(T, N, I, generator=g)
    done = torch.rand(T, N, 1, generator=g) < 0.011
    memory = {"hidden": torch.randn(N, L * H, generator=g) * 0.5, "cell": torch.randn(N, L * H, generator=g) * 0.5}
    gout = torch.randn(T, N, H, generator=g) / (T * N)
    out, _ = rnn(x, memory={k: v.clone() for k, v in memory.items()}, done=done)
    (out * gout).sum().backward()
    res = {"shape": np.array([T, N, I, H, L])}
    res.update(R.flatten_fingerprints("out", R.fingerprint(out, 4096)))
    names = []
    for k, p_ in rnn.named_parameters():
        names.append(k)
        res.update(R.flatten_fingerprints(f"grad/{k}", R.fingerprint(p_.grad, 1024)))
    res["param_names"] = np.array(names)
    torch.set_num_threads(1)
    save("lstm_anymal", **res)


def make_rnd_anymal(T: int = 24, N: int = 16384, D: int = 235):
    """RandomNetworkDistillation at BASELINE.json config 4: 16384 envs x 24 steps, nets 235 -> 128 -> 128 -> 16
    (cusrl_test/hook/auxiliary/test_rnd.py:13-17), one minibatch of T*N/4 samples for the objective."""
    import recipes as R
    from cusrl.hook.auxiliary.rnd import RandomNetworkDistillation

    torch.set_num_threads(8)
    hook = RandomNetworkDistillation(cusrl.Mlp.Factory([128, 128]), output_dim=16, reward_scale=0.1)
    recorded = {}
    hook.agent = SimpleNamespace(state_dim=D, setup_module=lambda m: m, record=lambda **kw: recorded.update(kw))
    hook.init()
    for net in ("target", "predictor"):
        R.set_seeded_parameters(((f"{net}.{k}", p) for k, p in getattr(hook, net).named_parameters()), seed=8)
    g = torch.Generator().manual_seed(9)
    buffer = Buffer(T, N, device="cpu")
    buffer["next_observation"] = torch.randn(T, N, D, generator=g)
    buffer["reward"] = torch.randn(T, N, 1, generator=g)
    hook.pre_update(buffer)
    out = {"shape": np.array([T, N, D])}
    out.update(R.flatten_fingerprints("reward_after", R.fingerprint(buffer["reward"], 4096)))
    out["rnd_reward_mean"] = recorded["rnd_reward"].double().mean().reshape(1)
    B = T * N // 4
    batch = {"next_observation": buffer["next_observation"].flatten(0, 1)[:B]}
    loss = hook.objective({}, batch)["rnd_loss"]
    loss.backward()
    out["rnd_loss"] = loss.detach().double().reshape(1)
    names = []
    for k, p in hook.predictor.named_parameters():
        names.append(k)
        out.update(R.flatten_fingerprints(f"grad/{k}", R.fingerprint(p.grad, 1024)))
    out["param_names"] = np.array(names)
    torch.set_num_threads(1)
    save("rnd_anymal", **out)


def make_symmetry():
    """One whole reference PPO iteration per symmetry hook (hook/auxiliary/symmetry.py:98-360 + SymmetricArchitecture) at a
    small shape, inputs regenerated from seeds on both sides (recipes.py): rollout outputs, the leaves the hook adds, the
    per-minibatch objectives in order, post-update parameters, metrics, LR."""
    import recipes as R
    from cusrl.hook.auxiliary.symmetry import (MirrorDef, MirrorSymmetryLoss, SymmetricArchitecture,
                                               SymmetricDataAugmentation, TransitionMirroring)
    from cusrl.template.environment import EnvironmentSpec
    import torch.distributions.normal as normal_mod

    sh = R.SYMMETRY_SHAPE
    N, T, obs_dim, act_dim, hidden = sh["N"], sh["T"], sh["obs"], sh["act"], sh["hidden"]
    out = {}
    for variant in R.SYMMETRY_VARIANTS:
        factory = cusrl.preset.ppo.PpoAgentFactory(
            num_steps_per_update=T, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation_fn="ELU", lr=1e-3,
            sampler_epochs=3, sampler_mini_batches=2, orthogonal_init=False, entropy_loss_weight=0.005,
            desired_kl_divergence=0.015, device="cpu").to_underlying()
        if variant == "augmentation":
            factory.register_hook(SymmetricDataAugmentation(), before="value_loss")
        elif variant == "mirror_loss":
            factory.register_hook(MirrorSymmetryLoss(0.5, symmetrize_action_std=True), after="ppo_surrogate_loss")
        elif variant == "transition_mirroring":
            factory.register_hook(TransitionMirroring(), index=0)
        else:
            factory.register_hook(SymmetricArchitecture(), after="module_initialization")
        spec = EnvironmentSpec(num_instances=N, observation_dim=obs_dim, action_dim=act_dim, reward_dim=1, autoreset=True,
                               final_state_is_missing=True,
                               mirror_observation=MirrorDef(*R.mirror_tables(obs_dim, seed=21)),
                               mirror_action=MirrorDef(*R.mirror_tables(act_dim, seed=22)))
        agent = factory(spec)
        R.set_seeded_parameters(agent.named_parameters(), seed=23)
        stream = R.anymal_stream(T, N, seed=24, obs_dim=obs_dim, p_term=0.1, p_trunc=0.05)
        noise = R.noise_stream(T, N, act_dim, seed=25)
        step = {"t": 0}
        real_normal = normal_mod._standard_normal
        normal_mod._standard_normal = lambda shape, dtype, device: noise[step["t"]].to(dtype=dtype, device=device).reshape(shape)
        actions = []
        try:
            for t in range(T):
                step["t"] = t
                actions.append(agent.act(stream["obs"][t]).clone())
                ready = agent.step(stream["obs"][t + 1], stream["reward"][t], stream["terminated"][t], stream["truncated"][t])
        finally:
            normal_mod._standard_normal = real_normal
        assert ready
        pre = f"{variant}/"
        out[pre + "returned_action"] = torch.stack(actions)
        for key, leaf in agent.buffer.storage.items():
            if key in ("observation", "action", "action_logp", "action_dist.mean", "value", "augmented_observation",
                       "augmented_action"):
                out[pre + f"buffer/{key}"] = leaf.clone()

        logs, names = [], []
        real_record = agent.record

        def spy_record(metrics=None, /, **kwargs):
            if "surrogate_loss" in kwargs:
                if not names:
                    names.extend(kwargs)
                logs.append([float(kwargs[k]) for k in names])
            return real_record(metrics, **kwargs)

        real_randperm = torch.randperm
        torch.randperm = R.SeededRandperm(seed=26)
        agent.record = spy_record
        try:
            metrics = agent.update()
        finally:
            torch.randperm = real_randperm
        out[pre + "objective_names"] = np.array(names)
        out[pre + "minibatch_losses"] = np.array(logs, dtype=np.float64)
        pnames = []
        for name, p in agent.named_parameters():
            pnames.append(name)
            out[pre + f"param1/{name}"] = p.detach().clone()
        out[pre + "param_names"] = np.array(pnames)
        out[pre + "metric_names"] = np.array(sorted(metrics))
        out[pre + "metric_values"] = np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64)
        out[pre + "lr_after"] = np.float64(agent.optimizer.param_groups[0]["lr"])
    save("symmetry", **out)


def make_obsnorm():
    """ObservationNormalization of the reference (hook/mdp/observation.py:161-215) over a few environment steps: the
    normalised observations it hands to the agent and its running statistics after every step."""
    from cusrl.hook.mdp.observation import ObservationNormalization

    out = {}
    for tag, (final_missing, state_dim) in {"a": (True, None), "b": (False, 7)}.items():
        N, C, steps = 64, 19, 5
        g = torch.Generator().manual_seed(41 if tag == "a" else 42)
        hook = ObservationNormalization()
        spec = SimpleNamespace(final_state_is_missing=final_missing, mirror_observation=None, mirror_state=None,
                               observation_is_subset_of_state=None, observation_stat_groups=(), state_stat_groups=(),
                               observation_normalization_excluded_indices=None, state_normalization_excluded_indices=None)
        hook.agent = SimpleNamespace(environment_spec=spec, observation_dim=C, state_dim=state_dim, has_state=state_dim is not None,
                                     inference_mode=False, setup_module=lambda m: m, to_tensor=torch.as_tensor)
        hook.init()
        obs = torch.randn(N, C, generator=g) * 3.0 + 1.0
        state = None if state_dim is None else torch.randn(N, state_dim, generator=g) * 0.5 - 2.0
        for t in range(steps):
            tr = {"observation": obs.clone()}
            if state is not None:
                tr["state"] = state.clone()
            hook.pre_act(tr)
            out[f"{tag}_obs_in_{t}"], out[f"{tag}_obs_norm_{t}"] = obs, tr["observation"]
            if state is not None:
                out[f"{tag}_state_in_{t}"], out[f"{tag}_state_norm_{t}"] = state, tr["state"]
            next_obs = torch.randn(N, C, generator=g) * (3.0 + t) + 1.0 - t
            next_state = None if state_dim is None else torch.randn(N, state_dim, generator=g) * 0.5 - 2.0 + 0.3 * t
            done = torch.rand(N, 1, generator=g) < 0.2
            tr2 = {"next_observation": next_obs.clone(), "done": done}
            if next_state is not None:
                tr2["next_state"] = next_state.clone()
            hook.post_step(tr2)
            out[f"{tag}_next_obs_in_{t}"], out[f"{tag}_next_obs_norm_{t}"], out[f"{tag}_done_{t}"] = next_obs, tr2["next_observation"], done
            if next_state is not None:
                out[f"{tag}_next_state_in_{t}"], out[f"{tag}_next_state_norm_{t}"] = next_state, tr2["next_state"]
            out[f"{tag}_mean_{t}"], out[f"{tag}_var_{t}"] = hook.observation_rms.mean.clone(), hook.observation_rms.var.clone()
            out[f"{tag}_count_{t}"] = np.int64(hook.observation_rms.count)
            obs, state = next_obs, next_state
        if state_dim is not None:
            out[f"{tag}_state_mean"], out[f"{tag}_state_var"] = hook.state_rms.mean.clone(), hook.state_rms.var.clone()
            out[f"{tag}_state_count"] = np.int64(hook.state_rms.count)
    save("obsnorm", **out)


if __name__ == "__main__":
    only = set(sys.argv[1:])
    big = {"iteration_anymal": make_iteration_anymal, "lstm_anymal": make_lstm_anymal, "rnd_anymal": make_rnd_anymal,
           "obsnorm": make_obsnorm, "symmetry": make_symmetry}
    if only:
        for name in only:
            (big.get(name) or globals()[f"make_{name}"])()
        sys.exit(0)
    for fn in big.values():
        fn()
    make_lstm()
    make_rnd()
    make_gae()
    make_advnorm()
    make_next_value()
    make_objective()
    make_sampler()
    make_iteration()
