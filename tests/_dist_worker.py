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
    # RunningMeanStd across ranks (nn/layer/rms.py:157-196): every rank streams DIFFERENT batches of different sizes; with
    # per-step synchronisation, and with the deferred protocol (local updates, one `synchronize()` per update), both ranks
    # must end with the statistics of the pooled data (population variance, count = all samples)
    from cusrl_b200.nn.rms import RunningMeanStd

    def batches(r):
        gr = torch.Generator().manual_seed(500 + r)
        return [torch.randn(5 + 3 * r + k, 6, generator=gr) * (1 + r) + 2 * r - k for k in range(4)]

    pooled = torch.cat([b for r in range(world) for b in batches(r)])
    pooled_var, pooled_mean = torch.var_mean(pooled, dim=0, correction=0)
    for mode in ("every_step", "deferred"):
        rms = RunningMeanStd(6)
        for batch in batches(rank):
            rms.update(batch, synchronize=(mode == "every_step"))
        if mode == "deferred":
            res["rms_deferred_is_local_before_sync"] = rms.count == sum(b.shape[0] for b in batches(rank))
            rms.synchronize()
        both = D.gather_stack(torch.cat([rms.mean, rms.var]))
        res[f"rms_{mode}"] = bool(torch.allclose(rms.mean, pooled_mean, rtol=1e-5, atol=1e-5)
                                  and torch.allclose(rms.var, pooled_var, rtol=1e-4, atol=1e-5)
                                  and rms.count == pooled.shape[0] and torch.allclose(both[0], both[1], rtol=0, atol=1e-6)
                                  and torch.allclose(rms.std, torch.sqrt(rms.var + rms.epsilon)))
    # C5
    res["avg_dict"] = D.average_dict({"a": float(rank), "only0": 5.0} if rank == 0 else {"a": float(rank)})
    D.barrier()
    if rank == 0:
        Path(out_path).write_text(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1])
