"""Running mean / standard deviation of a data stream, on the B200 kernels.

Mirrors the reference's ``RunningMeanStd`` (cusrl/nn/layer/rms.py:14-246): same constructor arguments, buffers (``mean``,
``var``, ``std``), ``count``, methods and cross-rank protocol (``synchronize`` / ``_synchronized_state``), same state-dict
layout (``get_extra_state`` = count).  The per-step arithmetic -- column statistics of the ``[N, C]`` batch, the parallel
merge into the running statistics, the normalisation -- runs as three kernel launches (csrc/rms_kernels.cu); the optional
statistic groups / excluded indices are tiny C-element index operations and stay in torch.
"""

from __future__ import annotations

from collections.abc import Iterable
from typing import Any

import torch
from torch import Tensor, nn

from .. import distributed, ops

__all__ = ["RunningMeanStd", "mean_var_count", "synchronize_mean_var_count"]


def mean_var_count(input: Tensor, *, uncentered: bool = False) -> tuple[Tensor, Tensor, int]:
    """(mean, population variance, count) over all but the last dim (nn/utils/normalization.py:15-49)."""
    if input.ndim < 2:
        raise ValueError("Input tensor must be at least 2-dimensional")
    x = input.reshape(-1, input.shape[-1])
    count = int(x.shape[0])
    C = x.shape[1]
    if count == 0:
        return x.new_zeros(C), x.new_ones(C), 0
    if uncentered or not x.is_cuda:
        if uncentered:
            var = x.square().mean(dim=0)
            return torch.zeros_like(var), var, count
        var, mean = torch.var_mean(x, dim=0, correction=0)
        return mean, var, count
    mean_var = ops.column_stats(x if x.stride(-1) == 1 else x.contiguous())
    return mean_var[:C], mean_var[C:], count


def synchronize_mean_var_count(mean: Tensor, var: Tensor, count: int) -> tuple[Tensor, Tensor, int]:
    """Count-weighted cross-rank merge with ONE all-gather (normalization.py:52-75)."""
    if not distributed.enabled():
        return mean, var, count
    count_tensor = torch.tensor([count], dtype=mean.dtype, device=mean.device)
    stacked = distributed.gather_stack(torch.cat((mean, var, count_tensor), dim=0))
    dim = mean.size(0)
    all_means, all_vars, all_counts = stacked[:, :dim], stacked[:, dim : 2 * dim], stacked[:, [2 * dim]]
    total_count = int(all_counts.sum().item())
    if total_count == 0:
        return mean, var, 0
    weights = all_counts / (total_count + 1e-8)
    total_mean = (all_means * weights).sum(dim=0)
    delta = all_means - total_mean
    total_var = torch.sum((all_vars + delta.square()) * weights, dim=0)
    return total_mean, total_var, total_count


class RunningMeanStd(nn.Module):
    def __init__(self, num_channels: int, *, groups: Iterable = (), excluded_indices=None, clamp: float | None = 10.0,
                 max_count: int | None = None, epsilon: float = 1e-8):
        if clamp is not None and clamp <= 0:
            raise ValueError("'clamp' must be None or a positive value")
        if max_count is not None and max_count <= 0:
            raise ValueError("'max_count' must be None or a positive value")
        self.groups = tuple(groups)
        self.excluded_indices = excluded_indices
        self.clamp, self.max_count, self.epsilon = clamp, max_count, epsilon
        dummy = torch.zeros(num_channels, dtype=torch.int64)
        for indices in self.groups:
            dummy[indices,] += 1
        if torch.any(dummy > 1):
            raise ValueError("Indices in 'groups' must not overlap")
        if excluded_indices is not None:
            mask = torch.zeros(num_channels, dtype=torch.bool)
            mask[excluded_indices,] = True
            if torch.any(dummy[mask] > 0):
                raise ValueError("'excluded_indices' must not overlap with 'groups'")
        super().__init__()
        self.register_buffer("mean", torch.zeros(num_channels))
        self.register_buffer("var", torch.ones(num_channels))
        self.register_buffer("std", torch.ones(num_channels))
        self.count: int = 0
        self._is_synchronized = True
        self._synchronized_state: tuple[Tensor, Tensor, int] | None = None

    def clear(self) -> None:
        self.mean.fill_(0.0)
        self.var.fill_(1.0)
        self.std.fill_(1.0)
        self.count = 0
        self._is_synchronized = False
        self._synchronized_state = None

    def update(self, input: Tensor, *, uncentered: bool = False, synchronize: bool = True) -> None:
        self.update_from_stats(*mean_var_count(input, uncentered=uncentered), synchronize=synchronize)

    @torch.no_grad()
    def update_from_stats(self, batch_mean: Tensor, batch_var: Tensor, batch_count: int, *, synchronize: bool = True) -> None:
        if synchronize:
            self.synchronize()
            batch_mean, batch_var, batch_count = synchronize_mean_var_count(batch_mean, batch_var, batch_count)
        if batch_count == 0:
            return
        if self.excluded_indices is not None or self.groups:
            batch_mean, batch_var = batch_mean.clone(), batch_var.clone()
            self._process_mean_var(batch_mean, batch_var)
        self._merge(self.mean, self.var, self.count, batch_mean, batch_var, batch_count, refresh_std=True)
        self.count += batch_count
        self._is_synchronized = synchronize
        if self._is_synchronized:
            if self.max_count is not None and self.count > self.max_count:
                self.count = self.max_count
            self._synchronized_state = (self.mean.clone(), self.var.clone(), self.count)

    def _merge(self, mean, var, w_old, new_mean, new_var, w_new, refresh_std: bool = False) -> None:
        """merge_mean_var_ (normalization.py:78-93); the running statistics go through the kernel (which also refreshes std)."""
        if w_old + w_new <= 0:
            raise ValueError(f"Weight sum must be positive; got {w_old + w_new}")
        if mean.is_cuda and refresh_std:
            ops.rms_merge_(mean, var, self.std, new_mean.contiguous(), new_var.contiguous(), w_old, w_new, self.epsilon)
            return
        w_sum = w_old + w_new
        wo, wn = w_old / w_sum, w_new / w_sum
        delta = new_mean - mean
        mean.add_(delta * wn)
        var.add_((new_var - var) * wn + delta.square() * (wo * wn))
        if refresh_std:
            self.std.copy_(torch.sqrt(self.var + self.epsilon))

    def synchronize(self) -> None:
        """rms.py:169-196: merge what this rank accumulated since the last synchronisation into the shared state."""
        if self._is_synchronized or not distributed.enabled():
            return
        if self._synchronized_state is None:
            total_mean, total_var, total_count = synchronize_mean_var_count(self.mean, self.var, self.count)
        else:
            sync_mean, sync_var, sync_count = self._synchronized_state
            self._merge(self.mean, self.var, self.count, sync_mean, sync_var, -sync_count)
            patch = synchronize_mean_var_count(self.mean, self.var, self.count - sync_count)
            self._merge(sync_mean, sync_var, sync_count, *patch)
            total_mean, total_var, total_count = sync_mean, sync_var, sync_count + patch[2]
        self.mean.copy_(total_mean)
        self.var.copy_(total_var)
        self.std.copy_(torch.sqrt(total_var + self.epsilon))
        self.count = total_count
        if self.max_count is not None and self.count > self.max_count:
            self.count = self.max_count
        self._is_synchronized = True
        self._synchronized_state = (total_mean, total_var, self.count)

    def forward(self, input: Tensor) -> Tensor:
        return self.normalize(input)

    def normalize(self, input: Tensor, out: Tensor | None = None, zero_padding: bool = False) -> Tensor:
        if input.is_cuda and input.dtype == torch.float32:
            x = input.reshape(-1, input.shape[-1])
            x = x if x.stride(-1) == 1 else x.contiguous()
            if out is None:
                return ops.rms_normalize(x, self.mean, self.std, self.clamp).reshape(input.shape)
            ops.rms_normalize(x, self.mean, self.std, self.clamp, out=out.reshape(-1, out.shape[-1]) if out.dim() != 2 else out,
                              zero_padding=zero_padding)
            return out
        output = (input - self.mean) / self.std
        if self.clamp is not None:
            output = output.clamp(-self.clamp, self.clamp)
        return output.type_as(input)

    def normalize_(self, input: Tensor) -> Tensor:
        if input.is_cuda and input.dtype == torch.float32 and input.stride(-1) == 1 and input.dim() == 2:
            ops.rms_normalize(input, self.mean, self.std, self.clamp, out=input)
            return input
        input.sub_(self.mean).div_(self.std)
        if self.clamp is not None:
            input.clamp_(-self.clamp, self.clamp)
        return input

    def unnormalize(self, input: Tensor) -> Tensor:
        return (input * self.std + self.mean).type_as(input)

    def unnormalize_(self, input: Tensor) -> Tensor:
        return input.mul_(self.std).add_(self.mean)

    def _process_mean_var(self, batch_mean: Tensor, batch_var: Tensor) -> None:
        if self.excluded_indices is not None:
            batch_mean[self.excluded_indices,] = 0.0
            batch_var[self.excluded_indices,] = 1.0
        for indices in self.groups:
            group_mean = batch_mean[indices,].mean()
            group_squared_mean = batch_mean[indices,].square().mean()
            group_var = batch_var[indices,].mean() - group_mean.square() + group_squared_mean
            batch_mean[indices,] = group_mean
            batch_var[indices,] = group_var

    def get_extra_state(self) -> Any:
        return torch.tensor(self.count, dtype=torch.int64)

    def set_extra_state(self, state: Any) -> None:
        count = int(state.item() if isinstance(state, Tensor) else state)
        if count < 0:
            raise ValueError("'count' must be non-negative")
        self.count = count
        self._is_synchronized = True
        self._synchronized_state = (self.mean.clone(), self.var.clone(), self.count)
