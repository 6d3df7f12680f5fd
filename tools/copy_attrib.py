"""Which Python lines launch the torch copy / fill kernels inside agent.update()? (torch.profiler with stacks)"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusrl_b200 as C
from bench import RolloutData, run_iteration
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.manual_seed(42)
envs = 65536
env = C.SyntheticEnvironment(envs, device=dev, seed=42)
agent = C.anymal_c_rough_ppo(device=dev).from_environment(env)
data = RolloutData(24, envs, dev, seed=1000, pinned_host=False)
run_iteration(agent, data)
for t in range(data.T):
    agent.act(data.obs[t]); agent.step(data.obs[t + 1], data.reward[t], data.terminated[t], data.truncated[t])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    agent.update()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_stack_n=12):
    if e.key in ("aten::copy_", "aten::fill_", "aten::zero_", "aten::zeros", "aten::contiguous", "aten::clone", "aten::mul", "aten::add") and e.device_time_total > 50:
        rows.append((e.device_time_total, e.count, e.key, [s for s in e.stack if "cusrl_b200" in s or "bench" in s][:4]))
for t, n, k, st in sorted(rows, key=lambda r: -r[0])[:14]:
    print(f"{t/1e3:7.2f} ms  x{n:4d}  {k}")
    for s in st:
        print("      ", s)
