"""cProfile of the host side of PPO iterations at a tiny env count (GPU time negligible -> what remains is host overhead)."""
import cProfile, pstats, sys, io
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cusrl_b200 as C
from bench import RolloutData, make_b200_agent, run_iteration
dev = torch.device("cuda", 0)
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
config = sys.argv[2] if len(sys.argv) > 2 else "mlp"
env = C.SyntheticEnvironment(envs, device=dev, seed=42)
agent = make_b200_agent(C, config, dev, env)
data = RolloutData(24, envs, dev, seed=1000, pinned_host=False)
for _ in range(3):
    run_iteration(agent, data)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    run_iteration(agent, data)
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45 if key == "tottime" else 60)
    print("\n".join(l[:170] for l in s.getvalue().splitlines()[4:]))
