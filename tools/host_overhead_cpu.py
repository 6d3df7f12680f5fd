"""Host-side (Python + dispatcher) cost of one PPO iteration with every cusrl_b200 kernel launch STUBBED OUT, on the CPU
box: what remains is exactly the per-launch overhead the GPU cannot hide when a rank's share is small (8 x B200: 8192
envs per rank, where the GPU work of an iteration is ~23 ms).  Results are garbage by construction; only time is read.

    python tools/host_overhead_cpu.py [--envs 256] [--iters 5] [--profile]
"""
import argparse
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=256)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--recurrent", action="store_true")
ap.add_argument("--count-ops", action="store_true", help="count ATen operators dispatched per iteration (deterministic)")
ap.add_argument("--graphs", action="store_true",
                help="drive the CUDA-graph runner with torch.cuda.graph mocked (capture = run the host code, replay = "
                     "nothing): checks its control flow and shows the host time a replayed step costs")
ap.add_argument("--rollout-only", action="store_true", help="time only the 24 x (act, step) of an iteration")
args = ap.parse_args()

real = _lib.load()
QUERIES = {"cusrl_b200_abi_version", "cusrl_b200_last_error", "cusrl_b200_sm_count"}


class Stub:
    calls = 0

    def __getattr__(self, name):
        fn = getattr(real, name)
        if name in QUERIES or name.endswith("_bytes") or name.endswith("_supported") or ("_set_" in name and "reset" not in name):
            return fn

        def launch(*a):
            Stub.calls += 1
            return 0

        return launch


stub = Stub()
_lib._lib = stub
_lib.load = lambda: stub
ops._stream = lambda: 0


ops._require_cuda = lambda t, name: None
real_sm = real.cusrl_b200_sm_count
stub.__dict__["cusrl_b200_sm_count"] = lambda: 148

import cusrl_b200 as C  # noqa: E402
from cusrl_b200.template.rollout import FusedRollout  # noqa: E402

FusedRollout.REQUIRE_CUDA = False
C.Rnn.REQUIRE_CUDA = False
from bench import RolloutData, run_iteration  # noqa: E402

dev = torch.device("cpu")
torch.manual_seed(0)
env = C.SyntheticEnvironment(args.envs, device=dev, seed=0)
factory = C.RecurrentPpoAgentFactory(device=dev) if args.recurrent else C.anymal_c_rough_ppo(device=dev)
agent = factory.from_environment(env)
data = RolloutData(24, args.envs, dev, seed=1, pinned_host=False)
if args.graphs:
    import contextlib

    from cusrl_b200.template.graphs import TrainStepGraphs

    class FakeGraph:
        def pool(self):
            return None

        def replay(self):
            pass

    torch.cuda.CUDAGraph = FakeGraph
    torch.cuda.graph = lambda g, pool=None, capture_error_mode=None: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.is_current_stream_capturing = lambda: False
    runner = TrainStepGraphs(agent)
    agent._train_step = runner
for _ in range(2):
    run_iteration(agent, data)
n0 = Stub.calls
pr = cProfile.Profile() if args.profile else None
torch.set_num_threads(1)  # tiny CPU tensors: intra-op threads only add noise
if args.count_ops:
    # deterministic proxy for host cost: ATen operators dispatched per iteration from the calling thread (every one is a
    # kernel launch or an allocation on the GPU); wall-clock on a shared CPU box is too noisy to compare small changes
    from collections import Counter

    from torch.utils._python_dispatch import TorchDispatchMode

    class OpCounter(TorchDispatchMode):
        def __init__(self):
            super().__init__()
            self.ops = Counter()

        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            self.ops[str(func.overloadpacket)] += 1
            return func(*args, **(kwargs or {}))

    n0 = Stub.calls
    with OpCounter() as counter:
        run_iteration(agent, data)
    total = sum(counter.ops.values())
    print(f"ATen ops dispatched in one iteration (calling thread): {total}; stubbed C-ABI calls: {Stub.calls - n0}")
    print("  " + ", ".join(f"{k.replace('aten.', '')} {v}" for k, v in counter.ops.most_common(14)))
if pr:
    pr.enable()
times = []
for _ in range(args.iters):
    t0 = time.perf_counter()
    if args.rollout_only:
        for t in range(data.T):
            agent.act(data.obs[t])
            agent.step(data.obs[t + 1], data.reward[t], data.terminated[t], data.truncated[t])
        times.append(time.perf_counter() - t0)
        if pr:
            pr.disable()
        agent.update()
        if pr:
            pr.enable()
        continue
    run_iteration(agent, data)
    times.append(time.perf_counter() - t0)
if pr:
    pr.disable()
dt = sum(times) / len(times)
print(f"host time per iteration: {dt * 1e3:.2f} ms (best {min(times) * 1e3:.2f} ms) with {(Stub.calls - n0) // args.iters} stubbed C-ABI calls "
      f"({args.envs} envs, CPU tensors, kernels not executed)")
if args.graphs:
    print(f"graph runner: {runner.captures} capture(s), {runner.replays} replays, optimizer step_count "
          f"{agent.optimizer.step_count}, step_dev {int(agent.optimizer.step_dev.item())} (mock capture executes add_)")
if pr:
    for key in ("tottime", "cumulative"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(40)
        print("\n".join(line[:160] for line in s.getvalue().splitlines()[4:]))
