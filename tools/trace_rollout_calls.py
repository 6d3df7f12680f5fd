"""Symbolic trace of the C-ABI calls the fused rollout step issues (template/rollout.py), with every kernel launch STUBBED
OUT and CPU tensors standing in for device memory: each pointer argument is translated to ``<tensor name>+<byte offset>``
(buffer leaves, parameters, tensor-core operand copies, activation buffers, staging tensors), so that the host logic --
which slot, which source, which operand copy every launch is handed -- can be checked, and diffed between two versions of
the code, on a box without a GPU.

    python tools/trace_rollout_calls.py [--state-dim 48] [--out trace.txt] [--check]

`--check` asserts the addressing rules of the fused step (used by tests/test_rollout_host.py)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--state-dim", type=int, default=0)
ap.add_argument("--out", default=None)
ap.add_argument("--check", action="store_true")
ap.add_argument("--recurrent", action="store_true", help="the recurrent preset (FusedRecurrentRollout); trace only, no --check")
args = ap.parse_args()

real = _lib.load()
QUERIES = {"cusrl_b200_abi_version", "cusrl_b200_last_error", "cusrl_b200_sm_count"}
LOG: list[tuple[str, tuple]] = []


class Stub:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if name in QUERIES or name.endswith("_bytes") or name.endswith("_supported") or ("_set_" in name and "reset" not in name):
            return fn

        def launch(*a):
            LOG.append((name, a))
            return 0

        return launch


stub = Stub()
stub.__dict__["cusrl_b200_sm_count"] = lambda: 148
_lib._lib = stub
_lib.load = lambda: stub
ops._stream = lambda: 0
ops._require_cuda = lambda t, name: None

import cusrl_b200 as C  # noqa: E402
from cusrl_b200.template.rollout import FusedRollout  # noqa: E402

FusedRollout.REQUIRE_CUDA = False
if hasattr(C.Rnn, "REQUIRE_CUDA"):
    C.Rnn.REQUIRE_CUDA = False     # drive the recurrent inference path (one sequence-kernel launch per layer) on CPU tensors
state_dim = args.state_dim or None
N, T, OBS, ACT = 64, 6, 235, 12
torch.manual_seed(0)
spec = C.EnvironmentSpec(N, OBS, ACT, state_dim=state_dim, autoreset=True, final_state_is_missing=True)
if args.recurrent:
    agent = C.RecurrentPpoAgentFactory(num_steps_per_update=T, actor_hidden_size=128, critic_hidden_size=128, actor_num_layers=2,
                                       critic_num_layers=1, device=torch.device("cpu"))(spec)
else:
    agent = C.anymal_c_rough_ppo(num_steps_per_update=T, device=torch.device("cpu"))(spec)
agent.cuda_graphs = False
g = torch.Generator().manual_seed(1)
obs = [torch.randn(N, OBS, generator=g) for _ in range(T + 1)]
state = [torch.randn(N, state_dim, generator=g) if state_dim else None for _ in range(T + 1)]
reward = [torch.randn(N, 1, generator=g) for _ in range(T)]
terminated = [torch.rand(N, 1, generator=g) < 0.1 for _ in range(T)]
truncated = [torch.rand(N, 1, generator=g) < 0.1 for _ in range(T)]
ITERATIONS = 3
for it in range(ITERATIONS):
    for t in range(T):
        LOG.append(("--act", (it, t)))
        hand_back = it != 1      # iteration 1: act() receives a fresh array, not the one step() got as next_observation
        # the Trainer's loop: the observation of step 0 of a later iteration is the last next_observation of the one before
        o, st = (obs[T], state[T]) if (it > 0 and t == 0) else (obs[t], state[t])
        agent.act(o if hand_back else o.clone(), st if (hand_back or st is None) else st.clone())
        LOG.append(("--step", (it, t)))
        agent.step(obs[t + 1], reward[t], terminated[t], truncated[t], state[t + 1])
    n = len(LOG)
    agent.update()       # re-splits nothing here (stubbed), but bumps the weights' epoch like a real optimizer step
    del LOG[n:]          # the update's own launches are not under test
# a caller that moves the cursor between step() and act(): the hand-over must still read the slot that step() wrote
LOG.append(("--extra", ()))
for t in range(3):
    agent.act(obs[t], state[t])
    agent.step(obs[t + 1], reward[t], terminated[t], truncated[t], state[t + 1])
agent.buffer.reset_cursor()
LOG.append(("--moved", ()))
agent.act(obs[3], state[3])

# ---- pointer -> name+offset --------------------------------------------------------------------------------------------------
registry: list[tuple[int, int, str]] = []


def register(name: str, t) -> None:
    if isinstance(t, torch.Tensor) and t.numel():
        registry.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), name))


for key, leaf in agent.buffer._backing.items():
    register("buf." + key, leaf)
for name, p in agent.named_parameters():
    register("param." + name, p)
rollout = agent._fused_rollout
for key, t in rollout._stage.items():
    register("stage." + key, t)
for i, net in enumerate(rollout._nets or ()):
    for j, a in enumerate(net.acts):
        register(f"net{i}.act{j}", a)
names = {p.data_ptr(): n for n, p in agent.named_parameters()}
for key, (_stamp, wp, _owner) in ops._weight_cache.items():
    for kind, t in wp.items():
        register(f"wp.{names.get(key, key)}.{kind}", t)


def symbol(x):
    if isinstance(x, int) and x > (1 << 20):
        for lo, hi, name in registry:
            if lo <= x < hi:
                return f"{name}+{x - lo}"
        return "ptr?"
    return x


lines = [(name, [symbol(x) for x in a]) for name, a in LOG]
if args.out:
    Path(args.out).write_text("".join(name + " " + " ".join(map(str, a)) + "\n" for name, a in lines))
print(f"fast steps {rollout.fast_steps}, {len(LOG)} log lines")

if args.check and args.recurrent:
    raise SystemExit("--check covers the feed-forward step; use --recurrent --out to diff two versions of the code")
if args.check:
    assert rollout.fast_steps == ITERATIONS * T - 1 + 3, rollout.fast_steps     # only the allocating first step is generic (+ the 3 of the moved-cursor scenario)
    pitch_obs = agent.buffer.backing("observation").stride(1) * 4            # padded row pitch in bytes
    assert pitch_obs == 944
    per_step = {}
    key = None
    for name, a in lines:
        if name.startswith("--"):
            key = (name[2:], *a)
            per_step[key] = []
        else:
            per_step[key].append((name, a))
    for it in range(ITERATIONS):
        for t in range(T):
            if it == 0 and t == 0:
                continue
            act, step = per_step[("act", it, t)], per_step[("step", it, t)]
            kernels = [n.replace("cusrl_b200_", "") for n, _ in act if "weight_prep" not in n]
            wide = 2 if state_dim else 1
            assert kernels == (["copy_rows_padded_f32"] * wide + ["linear_fwd_tf32"] * 3 + ["head_fwd_f32", "sample_logp_f32"]
                               + ["linear_fwd_tf32"] * 3 + ["head_fwd_f32"]), kernels
            calls = [(n, a) for n, a in act if "weight_prep" not in n]
            # the only re-splits of the tensor-core operand copies: the first act after an update (the critic's three trunk
            # layers; the actor's were refreshed by OnPolicyStatistics' pass over the buffer after the last optimizer step)
            assert sum("weight_prep" in n for n, _ in act) == (3 if (t == 0 and it > 0) else 0)
            off = lambda leaf, width_bytes: f"buf.{leaf}+{t * N * width_bytes}"  # noqa: E731
            copy = calls[0][1]
            # handed back = the very array the previous FUSED step() received (t == 0: obs[0] never was a next_observation;
            # step 1 of the run: the step before it was the generic, allocating one)
            handed_back = it != 1 and (t > 0 or it > 0) and (it, t) != (0, 1)
            prev = f"buf.next_observation+{((t - 1) % T) * N * pitch_obs}"
            assert copy[0] == (prev if handed_back else "stage.observation+0"), (it, t, copy)
            assert copy[2] == off("observation", pitch_obs) and copy[3] == 236 and copy[4:6] == [N, OBS]
            first_actor, first_critic = calls[wide][1], calls[wide + 5][1]
            assert first_actor[0] == off("observation", pitch_obs)
            critic_leaf, critic_pitch = ("state", state_dim * 4) if state_dim else ("observation", pitch_obs)
            assert first_critic[0] == off(critic_leaf, critic_pitch), (first_critic[0], off(critic_leaf, critic_pitch))
            assert calls[wide + 3][1][4] == off("action_dist.mean", ACT * 4)           # mean head -> its slot
            assert calls[wide + 8][1][4] == off("value", 4)                             # value head -> its slot
            sample = calls[wide + 4][1]
            assert sample[0] == off("action_dist.mean", ACT * 4) and sample[1] == "param.actor.distribution.std.param+0"
            assert sample[6:9] == [off("action_dist.std", ACT * 4), off("action", ACT * 4), off("action_logp", 4)]
            assert len(step) == 1 and step[0][0] == "cusrl_b200_rollout_store_step_f32"
            s = step[0][1]
            assert s[0] == "stage.next_observation+0" and s[2] == off("next_observation", pitch_obs) and s[4] == OBS
            if state_dim:
                assert s[5] == "stage.next_state+0" and s[7] == off("next_state", state_dim * 4) and s[9] == state_dim
            else:
                assert s[5:10] == [None, 0, None, 0, 0]
            assert s[10:13] == ["stage.reward+0", off("reward", 4), 1]
            assert s[13:15] == ["stage.terminated+0", "stage.truncated+0"]
            assert s[15:19] == [off("terminated", 1), off("truncated", 1), off("done", 1), N]
    moved = [a for n, a in per_step[("moved",)] if "copy_rows_padded" in n][0]
    assert moved[0] == f"buf.next_observation+{2 * N * pitch_obs}" and moved[2] == "buf.observation+0", moved
    print("addressing rules hold")
