"""GPU: the B200 LSTM backbone (K7) against the reference's Rnn (nn.LSTM + split/pad/scatter) golden and the oracle."""

from __future__ import annotations

import pytest
import torch

from oracle import ppo_path as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from cusrl_b200 import build

    build.build()
    import cusrl_b200

    return cusrl_b200


def _module(C, g, I=19, H=32, L=2):
    from cusrl_b200.nn.recurrent import Rnn

    rnn = Rnn.Factory("LSTM", hidden_size=H, num_layers=L)(I).to("cuda")
    with torch.no_grad():
        for k, p in rnn.named_parameters():
            p.copy_(g.t(f"param/{k}", "cuda"))
    return rnn


def test_lstm_sequence_with_done_matches_reference(C, golden):
    g = golden("lstm")
    rnn = _module(C, g)
    memory = {"hidden": g.t("hidden0", "cuda"), "cell": g.t("cell0", "cuda")}
    out, mem = rnn(g.t("x", "cuda"), memory=memory, done=g.t("done", "cuda"))
    assert mem is None
    assert torch.allclose(out.cpu(), g.t("out"), rtol=1e-5, atol=2e-6)   # reference test_rnn.py tolerance: atol 1e-5
    (out * g.t("gout", "cuda")).sum().backward()
    for k, p in rnn.named_parameters():
        ref = g.t(f"grad/{k}")
        assert torch.allclose(p.grad.cpu(), ref, rtol=1e-4, atol=2e-5 * max(ref.abs().max().item(), 1.0)), k


def test_lstm_single_step_memory(C, golden):
    g = golden("lstm")
    rnn = _module(C, g)
    memory = {"hidden": g.t("hidden0", "cuda"), "cell": g.t("cell0", "cuda")}
    with torch.no_grad():
        out, mem = rnn(g.t("x", "cuda")[0], memory=memory, sequential=False)
    assert torch.allclose(out.cpu(), g.t("step_out"), rtol=1e-5, atol=2e-6)
    assert torch.allclose(mem["hidden"].cpu(), g.t("step_hidden"), rtol=1e-5, atol=2e-6)
    assert torch.allclose(mem["cell"].cpu(), g.t("step_cell"), rtol=1e-5, atol=2e-6)


def test_lstm_step_by_step_equals_sequence(C):
    """cusrl_test/nn/module/test_rnn.py:88-121,145-164: stepping with resets == sequence with done."""
    from cusrl_b200.nn.recurrent import Rnn

    torch.manual_seed(0)
    T, N, I, H = 9, 33, 20, 64
    rnn = Rnn.Factory("LSTM", hidden_size=H, num_layers=2)(I).to("cuda")
    x = torch.randn(T, N, I, device="cuda")
    done = torch.rand(T, N, 1, device="cuda") < 0.25
    with torch.no_grad():
        seq_out, _ = rnn(x, memory=None, done=done)
        mem, outs = None, []
        for t in range(T):
            o, mem = rnn(x[t], memory=mem, sequential=False)
            rnn.reset_memory(mem, done[t])
            outs.append(o)
    assert torch.allclose(seq_out, torch.stack(outs), rtol=1e-5, atol=1e-5)
    # and against the oracle restatement
    ws = [(getattr(rnn.rnn, f"weight_ih_l{l}").detach().cpu(), getattr(rnn.rnn, f"weight_hh_l{l}").detach().cpu(),
           getattr(rnn.rnn, f"bias_ih_l{l}").detach().cpu(), getattr(rnn.rnn, f"bias_hh_l{l}").detach().cpu()) for l in range(2)]
    ref, _, _ = O.lstm_sequence_ref(x.cpu(), done.cpu(), torch.zeros(2, N, H), torch.zeros(2, N, H), ws)
    assert torch.allclose(seq_out.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_recurrent_ppo_trainer_and_train_rollout_consistency(C, monkeypatch):
    """Recurrent PPO end to end (BASELINE config 3 shape at a small env count) + the reference's train/rollout
    consistency oracle (cusrl_test/_helpers.py:76-94): on the first minibatch, BEFORE any optimizer step, the action
    mean recomputed from the stored memories equals the mean stored during the rollout to 1e-4."""
    # the kernels' numerics in this comparison were validated on a B200 with the LSTM at torch's default initialisation:
    # keep those weights (ModuleInitialization's orthogonal initialisation of recurrent layers is host-side torch code,
    # covered on the CPU by tests/test_plugin_reference.py against the reference's hook, seed for seed)
    monkeypatch.setattr(C.ModuleInitialization, "_init_rnn", lambda *args, **kwargs: None)
    env = C.SyntheticEnvironment(256, device="cuda", seed=9, p_term=0.05, p_trunc=0.01)
    factory = C.RecurrentPpoAgentFactory(device="cuda", entropy_loss_weight=0.005, desired_kl_divergence=0.015).to_underlying()

    class Consistency(C.Hook):
        checked = False

        def objective(self, metadata, batch):
            if not Consistency.checked:
                mean = batch["curr_action_dist"]["mean"]
                stored = self.agent.buffer["action_dist"]["mean"]
                cols = self.last_indices
                assert torch.allclose(mean.detach(), stored[:, cols], rtol=1e-4, atol=1e-4)
                Consistency.checked = True

    probe = Consistency()
    factory.register_hook(probe, after="on_policy_preparation")
    agent = factory.from_environment(env)
    impl_indices = []
    orig = C.TemporalMiniBatchSampler.indices

    def spy(self, buffer):
        for meta, idx in orig(self, buffer):
            probe.last_indices = idx.clone()
            yield meta, idx

    C.TemporalMiniBatchSampler.indices = spy
    try:
        history = C.Trainer(env, agent, num_iterations=2).run_training_loop()
    finally:
        C.TemporalMiniBatchSampler.indices = orig
    assert Consistency.checked
    names = [n for n, _ in agent.named_parameters()]
    assert "actor.backbone.rnn.weight_hh_l1" in names and "critic.backbone.rnn.bias_ih_l0" in names
    assert sum(p.numel() for p in agent.actor.parameters()) == 1034264     # reference count
    assert sum(p.numel() for p in agent.critic.parameters()) == 1031425
    for key in ("Agent/value_loss", "Agent/surrogate_loss", "Agent/kl_divergence", "Agent/grad_norm/default"):
        import numpy as np

        assert np.isfinite(history[-1][key]), key


@pytest.mark.parametrize("T,Nb,I,H,L", [(24, 1024, 235, 256, 2), (6, 200, 19, 64, 1), (5, 2500, 40, 128, 2), (1, 4096, 235, 256, 2),
                                        (3, 128, 16, 192, 1)])
def test_sequence_resident_kernel_equals_per_step_path(C, monkeypatch, T, Nb, I, H, L):
    """csrc/lstm_seq.cu (one launch per layer, recurrence on the SMs) against the per-step recurrence (one GEMM + one cell
    kernel per step): outputs, final state, and every parameter gradient computed from the tensors it saves.  Shapes cover
    the config-3 minibatch, a ragged last row tile, more row tiles than resident CTA groups, the single-step rollout call."""
    from cusrl_b200 import ops
    from cusrl_b200.nn.recurrent import Rnn

    torch.manual_seed(T * 1000 + Nb)
    rnn = Rnn.Factory("LSTM", hidden_size=H, num_layers=L)(I).to("cuda")
    x = torch.randn(T, Nb, I, device="cuda")
    done = (torch.rand(T, Nb, 1, device="cuda") < 0.1) if T > 1 else None
    memory = {"hidden": torch.randn(Nb, L * H, device="cuda").tanh() * 0.9, "cell": torch.randn(Nb, L * H, device="cuda")}
    gout = torch.randn(T, Nb, H, device="cuda") / (T * Nb) ** 0.5
    results = {}
    for mode in (True, False):
        monkeypatch.setattr(ops, "LSTM_SEQ", mode)
        assert ops.lstm_seq_supported(H) is mode
        rnn.zero_grad(set_to_none=True)
        ops.invalidate_weight_cache()
        if T > 1:
            out, mem = rnn(x, memory=memory, done=done)
            assert mem is None
            out2, mem2 = rnn(x, memory=memory)   # no episode cuts: the final state is returned
        else:
            out, mem2 = rnn(x[0], memory=memory, sequential=False)
            out2 = out
        (out * gout.reshape(out.shape)).sum().backward()
        results[mode] = (out.detach().clone(), out2.detach().clone(), mem2["hidden"].clone(), mem2["cell"].clone(),
                         {k: p.grad.clone() for k, p in rnn.named_parameters()})
        if mode:
            # the no-autograd path (rollout, statistics, bootstrap values): same kernel, nothing saved, the flat memory read
            # and written in place -> identical numbers
            with torch.no_grad():
                if T > 1:
                    out_i, mem_i = rnn(x, memory=memory, done=done)
                    assert mem_i is None and torch.equal(out_i, out.detach())
                    out2_i, mem2_i = rnn(x, memory=memory)
                else:
                    out2_i, mem2_i = rnn(x[0], memory=memory, sequential=False)
            assert torch.equal(out2_i, out2.detach())
            assert torch.equal(mem2_i["hidden"], mem2["hidden"]) and torch.equal(mem2_i["cell"], mem2["cell"])
    a, b = results[True], results[False]
    for i in range(4):
        assert torch.allclose(a[i], b[i], rtol=1e-5, atol=2e-6), (i, (a[i] - b[i]).abs().max().item())
    for k in a[4]:
        scale = max(b[4][k].abs().max().item(), 1e-6)
        assert torch.allclose(a[4][k], b[4][k], rtol=1e-4, atol=2e-5 * scale), (k, (a[4][k] - b[4][k]).abs().max().item(), scale)
