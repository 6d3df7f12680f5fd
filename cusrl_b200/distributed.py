"""Cross-rank helpers for env-axis data parallelism (the only parallelism the reference has).

Same names and semantics as the reference's ``cusrl/utils/distributed.py`` (reduce_gradients :145-172,
reduce_mean_ :101-110, reduce_mean_var_ :175-183, broadcast_parameters :58-63, average_dict :35-48),
re-designed around a persistent flat gradient arena: the allreduce runs in place on the buffer the
weight-gradient kernels wrote, with no ``cat`` and no per-parameter copy-back."""

from __future__ import annotations

from collections.abc import Iterable

import torch

from .runtime import CONFIG, configure_distributed

__all__ = [
    "average_dict",
    "barrier",
    "broadcast_parameters",
    "enabled",
    "is_main_process",
    "rank",
    "reduce_gradients",
    "reduce_mean_",
    "reduce_mean_var_",
    "world_size",
]


def enabled() -> bool:
    return CONFIG.distributed


def rank() -> int:
    return CONFIG.rank


def world_size() -> int:
    return CONFIG.world_size


def is_main_process() -> bool:
    return CONFIG.rank == 0


def barrier() -> None:
    if configure_distributed():
        torch.distributed.barrier()


def reduce_mean_(tensor: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks (NCCL: one AVG allreduce; Gloo: SUM then divide)."""
    if not configure_distributed():
        return tensor
    if torch.distributed.get_backend() == torch.distributed.Backend.GLOO:
        torch.distributed.all_reduce(tensor, op=torch.distributed.ReduceOp.SUM)
        return tensor.div_(CONFIG.world_size)
    torch.distributed.all_reduce(tensor, op=torch.distributed.ReduceOp.AVG)
    return tensor


def reduce_gradients(optimizer) -> None:
    """Average gradients over ranks, once per minibatch, BEFORE clipping (actor_critic.py:312-315).

    For :class:`cusrl_b200.template.optimizer.FlatAdam` the whole gradient is one contiguous arena and is
    reduced in place with a single collective; any other optimizer falls back to per-parameter reduction
    of the same mean."""
    if not configure_distributed():
        return
    flat = getattr(optimizer, "flat_grad", None)
    if flat is not None:
        reduce_mean_(flat)
        return
    for group in optimizer.param_groups:
        for param in group["params"]:
            if param.grad is not None:
                reduce_mean_(param.grad)


def gather_stack(tensor: torch.Tensor) -> torch.Tensor:
    """[W, *shape] stack of `tensor` from every rank."""
    if not configure_distributed():
        return tensor.unsqueeze(0)
    if torch.distributed.get_backend() == torch.distributed.Backend.GLOO:
        parts = [torch.empty_like(tensor) for _ in range(CONFIG.world_size)]
        torch.distributed.all_gather(parts, tensor)
        return torch.stack(parts, dim=0)
    out = tensor.new_empty(CONFIG.world_size, *tensor.shape)
    torch.distributed.all_gather_into_tensor(out, tensor)
    return out


def merge_mean_var(all_mean_var: torch.Tensor) -> torch.Tensor:
    """Equal-weight merge of stacked per-rank [mean | var] rows (distributed.py:175-183) -> [2*Dv].
    CUDA tensors go through the K2 merge kernel; CPU tensors (Gloo host-logic tests) use torch."""
    if all_mean_var.is_cuda:
        from . import ops

        return ops.merge_mean_var(all_mean_var.contiguous())
    means, variances = all_mean_var.chunk(2, -1)
    mean = means.mean(dim=0)
    var = (variances + (means - mean).square()).mean(dim=0)
    return torch.cat((mean, var))


def reduce_mean_var_(mean_var: torch.Tensor) -> torch.Tensor:
    """In place: replace this rank's [mean | var] by the cross-rank merge (8 bytes per channel on the wire)."""
    if not configure_distributed():
        return mean_var
    mean_var.copy_(merge_mean_var(gather_stack(mean_var)))
    return mean_var


def broadcast_parameters(parameters: Iterable[torch.Tensor]) -> None:
    """Rank 0's values to everyone at construction (actor_critic.py:224).  A flat arena is one broadcast."""
    if not configure_distributed():
        return
    for param in parameters:
        torch.distributed.broadcast(param.data if isinstance(param, torch.nn.Parameter) else param, src=0)


def average_dict(info: dict[str, float]) -> dict[str, float]:
    """Average a python dict of floats over ranks (trainer.py:387); keys missing on a rank are skipped."""
    if not configure_distributed():
        return info
    gathered: list[dict[str, float] | None] = [None] * CONFIG.world_size
    torch.distributed.all_gather_object(gathered, info)
    out: dict[str, float] = {}
    for key in {k for d in gathered for k in d}:
        vals = [d[key] for d in gathered if d.get(key) is not None]
        if vals:
            out[key] = float(sum(vals) / len(vals))
    return out
