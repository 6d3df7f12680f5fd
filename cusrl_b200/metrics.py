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
    """Mean of the recorded means, weighted by their element counts (reference utils/metrics.py:11-40).

    A metric recorded once per update (KL divergence, action std, ...) is kept exactly as recorded.  From the second record
    on the weighted SUM is accumulated with one fused ``add_(mean, alpha=count)`` per record and divided once when read:
    the reference's running-mean update costs three small kernels per record, 20 records per metric and update."""

    __slots__ = ("_mean", "_sum", "count")

    def __init__(self):
        self._mean: torch.Tensor | None = None
        self._sum: torch.Tensor | None = None
        self.count: int = 0

    @property
    def mean(self) -> torch.Tensor:
        if self.count == 0:
            return torch.tensor([])
        return self._mean if self._sum is None else self._sum / float(self.count)

    @torch.no_grad()
    def update(self, mean: torch.Tensor, count: int) -> None:
        if count == 0:
            return
        if self.count == 0:
            self._mean, self._sum, self.count = mean.clone(), None, count
            return
        if self._sum is None:
            self._sum = self._mean * float(self.count)
            self._mean = None
        self._sum.add_(mean.to(self._sum.device), alpha=float(count))
        self.count += count


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
        """Merge triples collected in deferred mode (their tensors hold the values of the latest graph replay).  Once
        every metric of the list is in weighted-sum mode the merge is two multi-tensor launches (scale the means by their
        counts, add them to the sums) instead of one ``add_`` per metric."""
        metrics = [self._data.get(name) for name, _, _ in entries]
        if entries and all(m is not None and m._sum is not None for m in metrics) and len({id(m) for m in metrics}) == len(metrics):
            scaled = torch._foreach_mul([mean for _, mean, _ in entries], [float(count) for _, _, count in entries])
            torch._foreach_add_([m._sum for m in metrics], scaled)
            for m, (_, _, count) in zip(metrics, entries):
                m.count += count
            return
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

    def values(self):
        return self._data.values()

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
        means = [self._data[n].mean for n in names]
        device = next((m.device for m in means if m.is_cuda), torch.device("cpu"))
        packed = torch.stack([m.reshape(()).to(device) for m in means])
        values = packed.tolist()  # the single host sync of an update
        return {f"{prefix}{n}": v for n, v in zip(names, values)}
