"""Minibatch samplers: bit-exact index generation + ONE gather launch per minibatch (K8).

Mirrors the reference's ``MiniBatchSampler`` / ``TemporalMiniBatchSampler`` / ``AutoMiniBatchSampler``
(cusrl/sampler/mini_batch_sampler.py:12-140): same constructor arguments, metadata dicts and errors, and
the exact ``torch.randperm`` call pattern (one call before epoch 0, ``randperm(out=)`` at the start of
every later epoch, :56,67-68) so the permutations are bit-identical to the reference's on the same device
and seed.  The gather itself replaces 14 fancy-index kernels and 14 fresh allocations per minibatch by a
single launch into reusable destination buffers, optionally restricted to the leaves the objective
consumes (``fields``), with wide rows padded to 16-byte multiples for the first-layer GEMM.
"""

from __future__ import annotations

from collections.abc import Iterator, Sequence
from typing import Any

import torch

from . import ops
from .template.buffer import Buffer, Sampler, padded_width, rebuild_nested

__all__ = ["AutoMiniBatchSampler", "MiniBatchSampler", "TemporalMiniBatchSampler"]


class MiniBatchSampler(Sampler):
    """Shuffled minibatches of individual transitions from a full buffer."""

    temporal = False
    memory_first_step_only = False  # temporal sampler: gather only step 0 of `*memory` leaves (a [1, Nmb, .] batch leaf)

    def __init__(self, num_epochs: int = 1, num_mini_batches: int | Sequence[int] = 1, shuffle: bool = True,
                 fields: Sequence[str] | None = None):
        if num_epochs <= 0:
            raise ValueError("'num_epochs' must be positive")
        self.num_epochs = num_epochs
        if isinstance(num_mini_batches, int):
            if num_mini_batches <= 0:
                raise ValueError("'num_mini_batches' must be positive")
            self.num_mini_batches: int | tuple[int, ...] = num_mini_batches
        else:
            self.num_mini_batches = tuple(num_mini_batches)
            if len(self.num_mini_batches) != num_epochs:
                raise ValueError(
                    "'num_mini_batches' must be an integer or a sequence of integers with length "
                    f"equal to 'num_epochs' ({num_epochs}); got {len(self.num_mini_batches)} values")
            if any(v <= 0 for v in self.num_mini_batches):
                raise ValueError("'num_mini_batches' values must be positive")
        self.shuffle = shuffle
        self.fields = None if fields is None else tuple(fields)
        # leaves whose fp32 minibatch copy may be skipped when their fp16 pair is emitted (set by the agent when it can prove
        # that nothing but the f16x3 first layers consumes them: ActorCritic._configure_sampler)
        self.pair_only: frozenset[str] = frozenset()
        self._dst: dict[tuple, torch.Tensor] = {}
        self._pair_dst: dict[tuple, object] = {}
        self._pair_bounds: dict[str, torch.Tensor] = {}
        self._pair_bound_storage: dict[str, torch.Tensor] = {}

    # ---- index generation (pure torch, device agnostic: covered by the CPU tests) ----------------
    def _get_num_samples(self, buffer: Buffer) -> int:
        return buffer.capacity * buffer.get_parallelism()

    def _get_metadata(self) -> dict[str, Any]:
        return {"temporal": self.temporal}

    def indices(self, buffer: Buffer) -> Iterator[tuple[dict[str, Any], torch.Tensor]]:
        """Yield (metadata, index slice) exactly as the reference's sampler loop does (:52-78)."""
        if not (buffer.full and buffer.cursor == 0):
            raise RuntimeError("MiniBatchSampler requires a full buffer with cursor reset to 0")
        num_samples = self._get_num_samples(buffer)
        epoch_indices = torch.randperm(num_samples, device=buffer.device)
        for epoch in range(self.num_epochs):
            k = self.num_mini_batches if isinstance(self.num_mini_batches, int) else self.num_mini_batches[epoch]
            if k > num_samples:
                raise ValueError(f"'num_mini_batches' ({k}) cannot exceed the number of samples ({num_samples})")
            size = num_samples // k
            if self.shuffle and epoch > 0:
                torch.randperm(num_samples, device=buffer.device, out=epoch_indices)
            for mb in range(k):
                metadata = {"epoch_index": epoch, "mini_batch_index": mb, "total_epochs": self.num_epochs,
                            "total_mini_batches": k} | self._get_metadata()
                yield metadata, epoch_indices[mb * size : (mb + 1) * size]

    # ---- gather ------------------------------------------------------------------------------------
    def _selected(self, buffer: Buffer) -> list[str]:
        if self.fields is None:
            return list(buffer.storage)
        keep = []
        for key in buffer.storage:
            top = key.split(".")[0]
            if key in self.fields or top in self.fields:
                keep.append(key)
        return keep

    def _dst_for(self, key: str, leaf: torch.Tensor, lead: tuple[int, ...]) -> tuple[torch.Tensor, torch.Tensor]:
        """(dense padded destination, public view) for a leaf; reused across minibatches."""
        width = leaf.shape[-1] if leaf.dim() >= 3 else 1
        inner = tuple(leaf.shape[2:-1])
        padded = padded_width(width, leaf.dtype)
        cache_key = (key, lead, inner, padded, leaf.dtype)
        dst = self._dst.get(cache_key)
        if dst is None:
            dst = torch.zeros(*lead, *inner, padded, dtype=leaf.dtype, device=leaf.device)
            self._dst[cache_key] = dst
        return dst, (dst if padded == width else dst[..., :width])

    def _gather(self, buffer: Buffer, keys: list[str], idx: torch.Tensor) -> dict[str, torch.Tensor]:
        T, N = buffer.capacity, buffer.get_parallelism()
        out: dict[str, torch.Tensor] = {}
        groups: dict[int, tuple[torch.Tensor, list]] = {}
        emit_pairs = not self.temporal and ops.GEMM_PRECISION == 2 and idx.numel() >= ops.F16X3_MIN_ROWS
        if self.temporal:
            n_mb = idx.numel()
            all_rows = (torch.arange(T, device=idx.device).unsqueeze(1) * N + idx.unsqueeze(0)).reshape(-1)
        for key in keys:
            leaf = buffer.storage[key]
            back = buffer.backing(key)
            if not self.temporal:
                rows, lead = idx, (idx.numel(),)
            elif self.memory_first_step_only and key.split(".")[0].endswith("memory"):
                # only step 0 of a stored recurrent memory is ever consumed (reference recurrent.py:202-212)
                rows, lead = idx, (1, n_mb)
            else:
                rows, lead = all_rows, (T, n_mb)
            dst, view = self._dst_for(key, leaf, lead)
            # backing rows are dense [T*N, padded]: pass the padded payload so the row copies are 16-byte vectors
            src2 = back.reshape(T * N, -1)
            dst2 = dst.reshape(rows.numel(), -1)
            pair = (src2 if src2.shape[1] == dst2.shape[1] else src2[:, : min(src2.shape[1], dst2.shape[1])], dst2)
            out[key] = view
            if emit_pairs and key in self.pair_only and leaf.dim() == 3 and leaf.dtype == torch.float32:
                ops.mark_pair_only(view)   # its fp16 pair is emitted below; the fp32 rows are not gathered
                continue
            if self.pair_only:
                ops.unmark_pair_only(view)
            groups.setdefault(id(rows), (rows, []))[1].append(pair)
        for rows, pairs in groups.values():
            for i in range(0, len(pairs), 24):  # at most CUSRL_B200_MAX_GATHER_FIELDS (24) fields per launch
                ops.gather_rows(pairs[i : i + 24], rows)
        if emit_pairs:
            # f16x3 dense layers: the network inputs are ALSO emitted as fp16 hi / lo pairs by the gather itself (scale from
            # the amax of the whole leaf, computed once per update), so the first layer does not re-read the minibatch to
            # split it; the fp32 leaf stays in the batch for every other consumer
            for key in ("observation", "state"):
                if key not in out or buffer.storage[key].dim() != 3 or buffer.storage[key].dtype != torch.float32:
                    continue
                width = buffer.storage[key].shape[-1]
                src = buffer.backing(key).reshape(T * N, -1)[:, :width]
                bound = self._pair_bounds.get(key)
                if bound is None:
                    # a PERSISTENT device scalar: a captured train step has its address baked into the kernel arguments
                    bound = self._pair_bounds[key] = ops.amax(src, out=self._pair_bound_storage.setdefault(
                        key, torch.zeros(1, dtype=torch.float32, device=src.device)))
                dst = self._pair_dst.get((key, idx.numel()))
                pair = ops.gather_split_f16(src, idx, bound, out=dst)
                self._pair_dst[(key, idx.numel())] = pair
                ops.attach_pair(out[key], pair)
        return out

    def __call__(self, buffer: Buffer):
        keys = None
        self._pair_bounds = {}   # per update: the leaves change every rollout
        for metadata, idx in self.indices(buffer):
            if keys is None:
                keys = self._selected(buffer)
            leaves = self._gather(buffer, keys, idx)
            batch = {}
            for name, schema in buffer.schema.items():
                try:
                    batch[name] = rebuild_nested(leaves, schema)
                except KeyError:
                    continue  # field not selected
            yield metadata, batch


class TemporalMiniBatchSampler(MiniBatchSampler):
    """Shuffled minibatches of whole env columns (``leaf[:, idx]``), for recurrent policies (:92-114)."""

    temporal = True

    def _get_num_samples(self, buffer: Buffer) -> int:
        return buffer.get_parallelism()


class AutoMiniBatchSampler(Sampler):
    """Temporal iff any top-level field name ends with ``memory`` (:117-140)."""

    def __init__(self, num_epochs: int = 1, num_mini_batches: int | Sequence[int] = 1, shuffle: bool = True,
                 fields: Sequence[str] | None = None, memory_first_step_only: bool = False):
        self.num_epochs, self.num_mini_batches, self.shuffle, self.fields = num_epochs, num_mini_batches, shuffle, fields
        self.memory_first_step_only = memory_first_step_only
        self.pair_only: frozenset[str] = frozenset()
        self._impl: MiniBatchSampler | None = None

    def _resolve(self, buffer: Buffer) -> MiniBatchSampler:
        is_temporal = any(key.split(".")[0].endswith("memory") for key in buffer)
        cls = TemporalMiniBatchSampler if is_temporal else MiniBatchSampler
        if not isinstance(self._impl, cls) or type(self._impl) is not cls:
            self._impl = cls(self.num_epochs, self.num_mini_batches, self.shuffle, self.fields)
            self._impl.memory_first_step_only = self.memory_first_step_only
        self._impl.pair_only = self.pair_only
        return self._impl

    def indices(self, buffer: Buffer):
        return self._resolve(buffer).indices(buffer)

    def __call__(self, buffer: Buffer):
        return self._resolve(buffer)(buffer)
