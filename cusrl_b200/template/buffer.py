"""Rollout storage: time-major ``[capacity(T), parallelism(N), ...]`` leaves resident in HBM.

Interface-compatible with the reference's ``Buffer`` / ``Sampler`` (cusrl/template/buffer.py:16-207:
``push``, ``sample``, mapping access by top-level field name, lazily allocated flat ``storage`` keyed by
dotted leaf names, ``cursor`` / ``full``).  B200-specific layout decision: wide float leaves whose row
size is not a multiple of 16 bytes (``observation``: 235 floats = 940 B) are allocated with the last
dimension padded to a 16-byte multiple (236) and exposed as a narrow view, so the rows are legal TMA /
128-bit-vector sources for the gather and first-layer GEMM.  The public shape is unchanged.
"""

from __future__ import annotations

from collections.abc import Callable, Iterator, Mapping, MutableMapping
from typing import Any

import torch

from ..runtime import device as resolve_device

__all__ = ["Buffer", "Sampler", "flatten_nested", "rebuild_nested"]

Nested = Any  # tensor | None | dict[str, Nested] | tuple[Nested, ...]


def flatten_nested(value: Nested, prefix: str) -> Iterator[tuple[str, Any]]:
    """Yield (dotted leaf name, leaf) pairs; dict keys and tuple positions become name components."""
    if isinstance(value, Mapping):
        for k, v in value.items():
            yield from flatten_nested(v, f"{prefix}.{k}" if prefix else str(k))
    elif isinstance(value, (tuple, list)):
        for i, v in enumerate(value):
            yield from flatten_nested(v, f"{prefix}.{i}" if prefix else str(i))
    elif value is not None:
        yield prefix, value


def schema_of(value: Nested, prefix: str) -> Nested:
    """Same nesting as `value` with every leaf replaced by its dotted name."""
    if isinstance(value, Mapping):
        return {k: schema_of(v, f"{prefix}.{k}" if prefix else str(k)) for k, v in value.items()}
    if isinstance(value, (tuple, list)):
        return tuple(schema_of(v, f"{prefix}.{i}" if prefix else str(i)) for i, v in enumerate(value))
    return None if value is None else prefix


def rebuild_nested(leaves: Mapping[str, Any], schema: Nested) -> Nested:
    if isinstance(schema, Mapping):
        return {k: rebuild_nested(leaves, v) for k, v in schema.items()}
    if isinstance(schema, tuple):
        return tuple(rebuild_nested(leaves, v) for v in schema)
    return None if schema is None else leaves[schema]


def padded_width(width: int, dtype: torch.dtype) -> int:
    """Row width (elements) after padding rows of >= 64 bytes to a multiple of 16 bytes."""
    item = torch.empty((), dtype=dtype).element_size()
    if width * item < 64 or (width * item) % 16 == 0:
        return width
    per16 = 16 // item
    return (width + per16 - 1) // per16 * per16


class Buffer(MutableMapping):
    """Circular storage for nested tensors keyed by top-level field name."""

    def __init__(self, capacity: int, parallelism: int, device: str | torch.device | None = None):
        self.capacity = int(capacity)
        self.parallelism = int(parallelism)
        self.device = resolve_device(device)
        self.cursor = 0
        self.full = False
        self.schema: dict[str, Nested] = {}
        self.storage: dict[str, torch.Tensor] = {}   # public (possibly narrow) views
        self._backing: dict[str, torch.Tensor] = {}  # padded allocations behind `storage`
        # hand-shakes between ADJACENT B200 hooks that fuse their kernels across the hook boundary (hook/on_policy.py:
        # next_value -> GAE -> advantage statistics in one launch); never part of the data contract
        self.private: dict[str, Any] = {}
        # bumped whenever a leaf is allocated or dropped: holders of resolved slot addresses (template/rollout.py) re-resolve
        self.layout_version = 0

    # ---- bookkeeping ---------------------------------------------------------------------------
    def get_parallelism(self) -> int:
        return self.parallelism

    def clear(self) -> None:
        self.cursor, self.full = 0, False
        self.layout_version += 1
        self.private.clear()
        self.storage.clear()
        self._backing.clear()
        self.schema.clear()

    def reset_cursor(self) -> None:
        self.cursor = 0

    def resize(self, capacity: int) -> None:
        if capacity != self.capacity:
            self.clear()
            self.capacity = int(capacity)

    def backing(self, key: str) -> torch.Tensor:
        """The dense padded allocation behind leaf `key` (== storage[key] when no padding was needed)."""
        return self._backing[key]

    # ---- mapping interface over top-level fields -------------------------------------------------
    def __iter__(self):
        yield from self.schema

    def __contains__(self, key) -> bool:
        return key in self.schema

    def __len__(self) -> int:
        return len(self.schema)

    def __getitem__(self, key: str):
        return rebuild_nested(self.storage, self.schema[key])

    def get(self, key: str, default=None):
        schema = self.schema.get(key)
        return default if schema is None else rebuild_nested(self.storage, schema)

    def __setitem__(self, name: str, data: Nested) -> None:
        """Register or overwrite a whole field; every leaf must be ``[capacity, parallelism, ...]``."""
        if data is None:
            return
        self._check_schema(name, data)
        for key, value in flatten_nested(data, name):
            value = torch.as_tensor(value, device=self.device)
            if value.dim() < 3:
                raise ValueError(f"Field '{key}' must have shape [capacity, parallelism, ...]")
            if value.shape[0] != self.capacity:
                raise ValueError(f"Capacity mismatch for field '{key}': expected {self.capacity}, got {value.shape[0]}")
            if value.shape[1] != self.parallelism:
                raise ValueError(f"Parallelism mismatch for field '{key}': expected {self.parallelism}, got {value.shape[1]}")
            store = self.storage.get(key)
            if store is None:
                store = self._allocate(key, value.shape[1:], value.dtype)
            if store.data_ptr() != value.data_ptr() or store.stride() != value.stride():
                store.copy_(value)

    def __delitem__(self, name: str) -> None:
        if name not in self.schema:
            raise KeyError(f"Field '{name}' was not found")
        for _, leaf in flatten_nested(self.schema[name], ""):  # schema leaves are the dotted leaf names
            self.storage.pop(leaf, None)
            self._backing.pop(leaf, None)
        self.layout_version += 1
        del self.schema[name]

    # ---- rollout writes --------------------------------------------------------------------------
    def push(self, data: Mapping[str, Nested]) -> None:
        """Append one step; each leaf is ``[parallelism, ...]``.  First write fixes schema + allocation."""
        for name, nested in data.items():
            if nested is None:
                continue
            self._check_schema(name, nested)
            for key, value in flatten_nested(nested, name):
                value = torch.as_tensor(value, device=self.device)
                store = self.storage.get(key)
                if store is None:
                    if value.dim() < 2:
                        raise ValueError(f"A step of field '{key}' must have shape [parallelism, ...]")
                    if value.shape[0] != self.parallelism:
                        raise ValueError(
                            f"Parallelism mismatch for field '{key}': expected {self.parallelism}, got {value.shape[0]}")
                    store = self._allocate(key, value.shape, value.dtype)
                store[self.cursor].copy_(value)
        self.cursor += 1
        if self.cursor == self.capacity:
            self.full, self.cursor = True, 0

    def advance(self) -> None:
        """Move the cursor past a step whose leaves were written in place (template/rollout.py); same wrap-around as
        :meth:`push`."""
        self.cursor += 1
        if self.cursor == self.capacity:
            self.full, self.cursor = True, 0

    def sample(self, sampler: Callable[[str, torch.Tensor], torch.Tensor]) -> dict[str, Nested]:
        """Apply ``sampler(leaf_name, leaf_storage)`` to every leaf and rebuild the nesting."""
        batch = {key: sampler(key, leaf) for key, leaf in self.storage.items()}
        return {name: rebuild_nested(batch, schema) for name, schema in self.schema.items()}

    # ---- internals -------------------------------------------------------------------------------
    def _allocate(self, key: str, step_shape, dtype: torch.dtype) -> torch.Tensor:
        step_shape = tuple(step_shape)
        width = step_shape[-1] if len(step_shape) >= 2 else 1
        padded = padded_width(width, dtype) if len(step_shape) >= 2 else width
        if padded != width:
            back = torch.zeros(self.capacity, *step_shape[:-1], padded, dtype=dtype, device=self.device)
            view = back[..., :width]
        else:
            back = torch.zeros(self.capacity, *step_shape, dtype=dtype, device=self.device)
            view = back
        self._backing[key] = back
        self.storage[key] = view
        self.layout_version += 1
        return view

    def _check_schema(self, name: str, data: Nested) -> None:
        current = schema_of(data, name)
        known = self.schema.get(name)
        if known is None:
            self.schema[name] = current
        elif known != current:
            raise ValueError(f"Schema mismatch for field '{name}': expected '{known}', got '{current}'")


class Sampler:
    """Base sampler: one batch = every stored leaf untouched (reference template/buffer.py:193-207)."""

    def __call__(self, buffer: Buffer) -> Iterator[tuple[dict[str, Any], dict[str, Nested]]]:
        yield {}, buffer.sample(lambda _name, tensor: tensor)
