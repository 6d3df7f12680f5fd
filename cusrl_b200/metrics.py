"""Running means of recorded tensors, kept on the device, flushed with ONE device-to-host copy.

Same contract as the reference's ``Metrics`` (cusrl/utils/metrics.py:11-96: per-key running mean weighted
by numel) but ``summary`` stacks every mean and reads them back in a single transfer instead of one
``.item()`` (= one host sync) per key (reference agent.py:227-231)."""

from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import torch

__all__ = ["Metric", "Metrics"]


class Metric:
    __slots__ = ("mean", "count")

    def __init__(self):
        self.mean: torch.Tensor = torch.tensor([])
        self.count: int = 0

    @torch.no_grad()
    def update(self, mean: torch.Tensor, count: int) -> None:
        if count == 0:
            return
        if self.count == 0:
            self.mean, self.count = mean.clone(), count
            return
        total = self.count + count
        self.mean.mul_(self.count / total).add_(mean.to(self.mean.device) * (count / total))
        self.count = total


class Metrics:
    def __init__(self):
        self._data: dict[str, Metric] = {}
        # while a CUDA graph is being captured the merge into the running means is postponed: its weights
        # (count / total) are host numbers that differ from replay to replay (template/graphs.py)
        self._deferred: list[tuple[str, torch.Tensor, int]] | None = None

    def begin_deferred(self) -> None:
        """Until :meth:`end_deferred`, recorded (name, mean tensor, count) triples are collected instead of merged."""
        self._deferred = []

    def end_deferred(self) -> list[tuple[str, torch.Tensor, int]]:
        out, self._deferred = self._deferred or [], None
        return out

    def apply(self, entries) -> None:
        """Merge triples collected in deferred mode (their tensors hold the values of the latest graph replay)."""
        for name, mean, count in entries:
            self._data.setdefault(name, Metric()).update(mean, count)

    def _merge(self, name: str, mean: torch.Tensor, count: int) -> None:
        if self._deferred is not None:
            self._deferred.append((name, mean, count))
        else:
            self._data.setdefault(name, Metric()).update(mean, count)

    def clear(self) -> None:
        self._data.clear()

    def __getitem__(self, name: str) -> Metric:
        return self._data[name]

    def __iter__(self):
        return iter(self._data)

    def __len__(self) -> int:
        return len(self._data)

    def items(self):
        return self._data.items()

    def keys(self):
        return self._data.keys()

    def get(self, name: str, default=None):
        return self._data.get(name, default)

    @torch.no_grad()
    def record(self, metrics: Mapping[str, Any] | None = None, /, **kwargs: Any) -> None:
        for name, value in {**(metrics or {}), **kwargs}.items():
            if value is None:
                continue
            try:
                value = torch.as_tensor(value, dtype=torch.float32)
            except Exception as error:
                raise ValueError(f"Failed to update metric '{name}'") from error
            if value.numel() == 0:
                continue
            self._merge(name, value.mean(), value.numel())

    def record_mean(self, name: str, mean: torch.Tensor, count: int) -> None:
        """Record a mean that a kernel already reduced over `count` elements (0-dim device tensor)."""
        self._merge(name, mean.detach().reshape(()), count)

    def summary(self, prefix: str = "") -> dict[str, float]:
        if prefix and not prefix.endswith("/"):
            prefix += "/"
        if not self._data:
            return {}
        names = list(self._data)
        device = next((m.mean.device for m in self._data.values() if m.mean.is_cuda), torch.device("cpu"))
        packed = torch.stack([self._data[n].mean.reshape(()).to(device) for n in names])
        values = packed.tolist()  # the single host sync of an update
        return {f"{prefix}{n}": v for n, v in zip(names, values)}
