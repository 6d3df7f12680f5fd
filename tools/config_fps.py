"""env-steps/s of BASELINE.json's single-GPU parity configurations (configs[1..3]) driven exactly like bench.py's timed
region (24 x (agent.act + agent.step) + agent.update on synthetic HBM-resident rollouts, CUDA events).  These are not
bench lines -- bench.py reports the 65536-env configuration -- they document that the LSTM and RND paths run at size.

    python tools/config_fps.py [--steps 3] [--warmup 2]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusrl_b200 as C  # noqa: E402
from bench import RolloutData, time_iterations  # noqa: E402
from cusrl_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--no-graphs", action="store_true", help="eager train step instead of CUDA-graph replay (agent.cuda_graphs = False)")
ap.add_argument("--no-fused-rollout", action="store_true", help="generic act / step flow instead of template/rollout.py")
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)


def mlp(envs):
    return C.anymal_c_rough_ppo(device=dev)


def lstm(envs):
    return C.RecurrentPpoAgentFactory(device=dev)


def rnd(envs):
    factory = C.anymal_c_rough_ppo(device=dev).to_underlying()
    factory.register_hook(C.RandomNetworkDistillation(C.Mlp.Factory([128, 128]), output_dim=16, reward_scale=0.1),
                          before="value_computation")
    return factory


for name, envs, make in (("mlp_ppo_4096", 4096, mlp), ("lstm_ppo_4096", 4096, lstm), ("mlp_ppo_rnd_16384", 16384, rnd)):
    torch.manual_seed(42)
    env = C.SyntheticEnvironment(envs, device=dev, seed=42)
    if args.only and args.only not in name:
        continue
    agent = make(envs).from_environment(env)
    agent.cuda_graphs = not args.no_graphs
    agent.fused_rollout = not args.no_fused_rollout
    data = RolloutData(24, envs, dev, seed=1000, pinned_host=False)
    n0 = ops.launch_count()
    seconds, metrics = time_iterations(agent, data, args.steps, args.warmup, False)
    launches = (ops.launch_count() - n0) // (args.steps + args.warmup)
    print(json.dumps({"config": name, "cuda_graphs": not args.no_graphs, "fused_rollout": not args.no_fused_rollout, "envs": envs, "rollout_steps": 24, "env_steps_per_s": round(args.steps * 24 * envs / seconds, 1),
                      "ms_per_iteration": round(seconds / args.steps * 1e3, 3), "gpu_launches_per_iteration": int(launches),
                      "metrics": {k: round(v, 6) for k, v in metrics.items() if k.startswith("Agent/")}}), flush=True)
    del agent, data, env
    torch.cuda.empty_cache()
