"""CPU: the oracle's restatements (oracle/ppo_path.py) next to the LIVE reference's own functions on randomised inputs --
beyond the fixed golden vectors of tests/test_oracle_golden.py: many shapes (vector rewards, one step, one column), both
lambdas, both bootstrap branches, clipped and unclipped value losses, ties at the clipping boundary.  Uses the reference
package that travels as baseline/_ref (tools/install_reference.py); skipped when it is not there."""

from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from install_reference import reference_path  # noqa: E402

try:
    _PATHS = reference_path()
except RuntimeError:
    _PATHS = None
pytestmark = pytest.mark.skipif(_PATHS is None, reason="reference package not available (run tools/install_reference.py)")

from oracle import ppo_path as O  # noqa: E402


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


SHAPES = [(24, 64, 1), (24, 33, 1), (7, 5, 3), (1, 9, 1), (2, 1, 2), (50, 3, 1)]


def _rollout(T, N, Dv, seed, p_done=0.15):
    g = torch.Generator().manual_seed(seed)
    reward = torch.randn(T, N, Dv, generator=g) * 3
    value = torch.randn(T, N, Dv, generator=g) * 2
    next_value = torch.randn(T, N, Dv, generator=g) * 2
    terminated = torch.rand(T, N, 1, generator=g) < p_done
    truncated = (torch.rand(T, N, 1, generator=g) < p_done) & ~terminated
    return reward, value, next_value, terminated, truncated


@pytest.mark.parametrize("T,N,Dv", SHAPES)
@pytest.mark.parametrize("gamma,lamda", [(0.99, 0.95), (0.9, 1.0), (0.0, 0.5), (0.999, 0.0)])
def test_gae_scan_is_bit_identical_to_the_reference_function(reference, T, N, Dv, gamma, lamda):
    from cusrl.hook.on_policy.gae import _generalized_advantage_estimation

    for seed in range(3):
        reward, value, next_value, terminated, truncated = _rollout(T, N, Dv, seed)
        done = terminated | truncated
        theirs = _generalized_advantage_estimation(reward, done, value, next_value, gamma, lamda)
        assert torch.equal(O.gae_ref(reward, done, value, next_value, gamma, lamda), theirs)


@pytest.mark.parametrize("T,N,Dv", SHAPES)
@pytest.mark.parametrize("lamda_value", [None, 0.7, 1.0])
def test_advantage_and_return_match_the_reference_hook(reference, T, N, Dv, lamda_value):
    reward, value, next_value, terminated, truncated = _rollout(T, N, Dv, 5)
    done = terminated | truncated
    hook = reference.hook.GeneralizedAdvantageEstimation(gamma=0.97, lamda=0.9, lamda_value=lamda_value)
    data = {"reward": reward, "value": value, "next_value": next_value, "done": done}
    hook._compute_advantage_and_return(data)
    adv, ret = O.advantage_and_return_ref(reward, done, value, next_value, 0.97, 0.9, lamda_value)
    assert torch.equal(adv, data["advantage"]) and torch.equal(ret, data["return"])


@pytest.mark.parametrize("T,N,Dv", SHAPES)
@pytest.mark.parametrize("bootstrap", [False, True])
def test_next_value_matches_the_reference_hook(reference, T, N, Dv, bootstrap):
    """value.py:56-82 with a critic stand-in that is a fixed function of the state, so both bootstrap branches are defined."""
    reward, value, _, terminated, truncated = _rollout(T, N, Dv, 9, p_done=0.3)
    g = torch.Generator().manual_seed(1)
    next_state = torch.randn(T, N, 4, generator=g)

    def value_of(state):     # elementwise, so that evaluating a gathered subset and gathering the evaluation agree bit for bit
        return torch.stack([state[..., 0] * (0.5 + k) + state[..., 1] - state[..., 3] * state[..., 2] for k in range(Dv)], dim=-1)

    critic = SimpleNamespace(evaluate=lambda state, memory=None: value_of(state))
    hook = reference.hook.ValueComputation(termination_value=-1.5, bootstrap_truncated_states=bootstrap)
    from contextlib import nullcontext

    hook.agent = SimpleNamespace(critic=critic, autocast=nullcontext)
    buffer = {"value": value.clone(), "next_state": next_state, "terminated": terminated, "truncated": truncated}
    buffer_obj = type("B", (dict,), {"get": dict.get})(buffer)
    hook.pre_update(buffer_obj)
    ours = O.next_value_ref(value, terminated, truncated, value_of(next_state[-1]), termination_value=-1.5,
                            trunc_value=value_of(next_state) if bootstrap else None)
    assert torch.equal(ours, buffer_obj["next_value"])


@pytest.mark.parametrize("shape", [(24, 64, 1), (3, 1, 1), (8, 30, 3), (2, 2, 1)])
def test_advantage_normalisation_matches_the_reference_hook(reference, shape):
    g = torch.Generator().manual_seed(2)
    advantage = torch.randn(*shape, generator=g) * 4 + 1.5
    hook = reference.hook.AdvantageNormalization()
    theirs = advantage.clone()
    hook.normalize_(theirs)
    torch.testing.assert_close(O.normalize_advantage_ref(advantage), theirs, rtol=0, atol=0)


@pytest.mark.parametrize("B,A", [(64, 12), (1, 1), (33, 5)])
@pytest.mark.parametrize("loss_clip", [None, 0.2])
def test_objective_terms_match_the_reference_functions(reference, B, A, loss_clip):
    from cusrl.hook.on_policy.ppo import _ppo_surrogate_loss
    from cusrl.hook.on_policy.value import _clipped_value_loss

    g = torch.Generator().manual_seed(4)
    advantage = torch.randn(B, 1, generator=g)
    ratio = torch.exp(torch.randn(B, 1, generator=g) * 0.3)
    ratio[::5] = 1.2            # exactly on the clipping boundary: the tie-breaking of min / clamp matters for the gradient
    ratio[1::7] = 0.8
    ratio.requires_grad_(True)
    theirs = _ppo_surrogate_loss(advantage, ratio, 0.2)
    grad_theirs, = torch.autograd.grad(theirs, ratio)
    ratio2 = ratio.detach().clone().requires_grad_(True)
    ours = O.surrogate_loss_ref(advantage, ratio2, 0.2)
    grad_ours, = torch.autograd.grad(ours, ratio2)
    assert torch.equal(ours, theirs) and torch.equal(grad_ours, grad_theirs)

    value_old, curr, ret = (torch.randn(B, 1, generator=g) for _ in range(3))
    if loss_clip is None:
        expected = torch.nn.functional.mse_loss(ret, curr)
    else:
        expected = _clipped_value_loss(value_old, curr, ret, loss_clip)
    assert torch.equal(O.value_loss_ref(value_old, curr, ret, loss_clip), expected)

    mean, sample = torch.randn(B, A, generator=g), torch.randn(B, A, generator=g)
    std = torch.rand(B, A, generator=g) + 0.3
    dist = reference.NormalDist(4, A)
    params = {"mean": mean, "std": std}
    assert torch.equal(O.normal_log_prob_ref(mean, std, sample), dist.compute_logp(params, sample))
    assert torch.equal(O.normal_entropy_ref(std), dist.compute_entropy(params))
    other = {"mean": mean + 0.1, "std": std * 1.3}
    assert torch.equal(O.normal_kl_ref(mean, std, other["mean"], other["std"]), dist.compute_kl_div(params, other))


@pytest.mark.parametrize("T,N,I,H,L,p_done", [(6, 5, 7, 8, 1, 0.2), (9, 4, 3, 12, 2, 0.4), (5, 6, 10, 4, 3, 0.0), (4, 3, 5, 8, 2, 1.0),
                                              (1, 7, 6, 8, 2, 0.5), (12, 2, 235, 16, 2, 0.1)])
def test_lstm_in_line_reset_form_matches_the_reference_rnn(reference, T, N, I, H, L, p_done):
    """`lstm_sequence_ref` (the in-line state reset where `done`) next to the REFERENCE's `Rnn` (nn.LSTM behind its split-at-
    done / pad / scatter machinery, rnn.py:264-299, recurrent.py:160-272) on random sequences: no episode end, every step an
    episode end, one step, three layers, the Anymal observation width; initial memory given per step like a minibatch's."""
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(0)
    rnn = reference.Rnn.Factory("LSTM", hidden_size=H, num_layers=L)(I)
    x = torch.randn(T, N, I, generator=g)
    done = torch.rand(T, N, 1, generator=g) < p_done
    # a stored per-step memory [T, N, L*H] of which only step 0 is consumed (recurrent.py:202-212)
    hidden = torch.randn(T, N, L * H, generator=g) * 0.5
    cell = torch.randn(T, N, L * H, generator=g) * 0.5
    with torch.no_grad():
        theirs, _ = rnn(x, memory={"hidden": hidden, "cell": cell}, done=done)
    lstm = next(m for m in rnn.modules() if isinstance(m, torch.nn.LSTM))
    weights = [tuple(getattr(lstm, f"{name}_l{layer}").detach() for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
               for layer in range(L)]
    to_lnh = lambda m: m[0].reshape(N, L, H).transpose(0, 1).contiguous()  # noqa: E731   "n (k c) -> k n c"
    ours, _, _ = O.lstm_sequence_ref(x, done, to_lnh(hidden), to_lnh(cell), weights)
    torch.testing.assert_close(ours, theirs, rtol=1e-5, atol=1e-5)       # the reference's own tolerance (test_rnn.py:163)
