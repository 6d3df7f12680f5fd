"""Worker for tests/test_distributed_cpu.py: run under torchrun with 2 ranks on CPU (Gloo)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import cusrl_b200 as C  # noqa: E402  (reads RANK / WORLD_SIZE from the torchrun environment)
from cusrl_b200 import distributed as D  # noqa: E402
from cusrl_b200.runtime import CONFIG  # noqa: E402
from oracle import ppo_path as O  # noqa: E402


def main(out_path: str):
    CONFIG.device = "cpu"
    rank, world = D.rank(), D.world_size()
    assert D.enabled() and world == 2
    res = {}
    # C3: scalar mean
    x = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
    D.reduce_mean_(x)
    res["mean"] = x.tolist()
    # C2: advantage statistics merge vs the oracle restatement of utils/distributed.py:175-183
    g = torch.Generator().manual_seed(100 + rank)
    adv = torch.randn(24, 16, 2, generator=g) * (1 + rank) + rank
    var, mean = torch.var_mean(adv, dim=(0, 1))
    mv = torch.cat([mean, var])
    D.reduce_mean_var_(mv)
    stats = []
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        a = torch.randn(24, 16, 2, generator=gr) * (1 + r) + r
        v, m = torch.var_mean(a, dim=(0, 1))
        stats.append((m, v))
    om, ov = O.merge_mean_var_ref(torch.stack([s[0] for s in stats]), torch.stack([s[1] for s in stats]))
    res["mean_var_ok"] = bool(torch.allclose(mv, torch.cat([om, ov]), rtol=1e-6, atol=1e-7))
    # C4 + C1: flat arena broadcast and in-place gradient mean
    torch.manual_seed(rank)  # different initial weights per rank on purpose
    spec = C.EnvironmentSpec(4, 19, 5, autoreset=True, final_state_is_missing=True)
    agent = C.anymal_c_rough_ppo(device="cpu", actor_hidden_dims=(64, 128), critic_hidden_dims=(64, 128))(spec)
    flat = agent.optimizer.flat_param
    gathered = D.gather_stack(flat)
    res["params_equal_after_broadcast"] = bool(torch.equal(gathered[0], gathered[1]))
    agent.optimizer.flat_grad.fill_(float(rank + 1))
    D.reduce_gradients(agent.optimizer)
    res["grad_mean"] = float(agent.optimizer.flat_grad[0])
    res["grad_uniform"] = bool((agent.optimizer.flat_grad == agent.optimizer.flat_grad[0]).all())
    # C5
    res["avg_dict"] = D.average_dict({"a": float(rank), "only0": 5.0} if rank == 0 else {"a": float(rank)})
    D.barrier()
    if rank == 0:
        Path(out_path).write_text(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1])
