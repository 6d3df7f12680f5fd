"""Kernel-time table of one PPO iteration via torch.profiler (CUPTI), much cheaper than an ncu launch list.

    python tools/iter_profile.py [--envs 65536] [--iters 1]
"""
import argparse, sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusrl_b200 as C
from bench import RolloutData, make_b200_agent, run_iteration

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--config", default="mlp", choices=["mlp", "lstm", "rnd"])
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(42)
env = C.SyntheticEnvironment(args.envs, device=dev, seed=42)
agent = make_b200_agent(C, args.config, dev, env)
data = RolloutData(24, args.envs, dev, seed=1000, pinned_host=False)
for _ in range(2):
    run_iteration(agent, data)
torch.cuda.synchronize()
# phase split with events
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for t in range(data.T):
    agent.act(data.obs[t]); agent.step(data.obs[t + 1], data.reward[t], data.terminated[t], data.truncated[t])
ev[1].record()
agent.update()
ev[2].record()
torch.cuda.synchronize()
print(json.dumps({"rollout_ms": ev[0].elapsed_time(ev[1]), "update_ms": ev[1].elapsed_time(ev[2])}))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.iters):
        run_iteration(agent, data)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(r[1] for r in rows)
print(f"total device time {tot/1e3:.1f} ms over {args.iters} iteration(s)")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, t, n in sorted(rows, key=lambda r: -r[1])[:32]:
    print(f"| `{k[:230]}` | {n} | {t/1e3:.2f} | {100*t/tot:.1f}% | {t/n:.1f} |")
