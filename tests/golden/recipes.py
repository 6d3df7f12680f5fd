"""Seeded recipes shared by ``make_golden.py`` (build container, live reference) and the ``-m gpu`` parity tests.

Fixtures at BASELINE.json's shapes would be hundreds of MB if inputs were stored.  Instead both sides REGENERATE the
inputs from CPU generators with fixed seeds (bit-identical wherever this torch build runs) and the fixture stores only
what the reference computed from them: scalar series, checksums and strided samples (:func:`fingerprint`).
Test infrastructure only.
"""

from __future__ import annotations

import math
import zlib

import torch

OBS, ACT = 235, 12
_REAL_RANDPERM = torch.randperm  # bound at import: SeededRandperm is installed AS torch.randperm by its users


def seeded_parameter(name: str, shape, seed: int = 0) -> torch.Tensor:
    """Deterministic value for parameter `name`: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's Linear default, from a
    generator seeded by the NAME (independent of construction order)."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2**31))
    shape = tuple(shape)
    fan_in = shape[-1] if len(shape) > 1 else max(shape[0], 1)
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound


def set_seeded_parameters(named_parameters, seed: int = 0, skip=("std.param",)) -> None:
    with torch.no_grad():
        for name, p in named_parameters:
            if any(name.endswith(s) for s in skip):
                continue
            p.copy_(seeded_parameter(name, p.shape, seed).to(p.device))


def anymal_stream(T: int, N: int, seed: int, obs_dim: int = OBS, p_term: float = 0.01, p_trunc: float = 0.001) -> dict:
    """Synthetic env stream of SURVEY.md section 8(d) from a CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return {
        "obs": torch.randn(T + 1, N, obs_dim, generator=g),
        "reward": torch.randn(T, N, 1, generator=g),
        "terminated": torch.rand(T, N, 1, generator=g) < p_term,
        "truncated": torch.rand(T, N, 1, generator=g) < p_trunc,
    }


def noise_stream(T: int, N: int, A: int, seed: int) -> torch.Tensor:
    """Standard-normal exploration noise [T, N, A] (replaces the global-RNG draw of Normal.rsample on both sides)."""
    return torch.randn(T, N, A, generator=torch.Generator().manual_seed(seed))


class SeededRandperm:
    """Drop-in for ``torch.randperm`` drawing from a private CPU generator (identical on every machine)."""

    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)
        self.calls = 0

    def __call__(self, n, *, device=None, out=None, **kw):
        self.calls += 1
        p = _REAL_RANDPERM(n, generator=self.g)
        if out is not None:
            out.copy_(p)
            return out
        return p if device is None else p.to(device)


def fingerprint(t: torch.Tensor, samples: int = 256) -> dict:
    """{sum, abs_sum, sample}: fp64 checksums + `samples` strided entries of the flattened tensor."""
    flat = t.detach().reshape(-1).to("cpu")
    stride = max(1, flat.numel() // samples)
    f64 = flat.double()
    return {"sum": f64.sum().reshape(1), "abs_sum": f64.abs().sum().reshape(1), "sample": flat[::stride][:samples].clone()}


def flatten_fingerprints(prefix: str, fp: dict) -> dict:
    return {f"{prefix}/{k}": v for k, v in fp.items()}


def mirror_tables(dim: int, seed: int) -> tuple[list[int], list[int]]:
    """A self-inverse index-permute + sign-flip transform of width `dim` (pairs swap places and share a sign, like the
    reference's cusrl_test/_helpers.py:17-35): (destination_indices, flipped_indices) for ``MirrorDef``."""
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(dim, generator=g).tolist()
    signs = (torch.rand(dim, generator=g) < 0.5).tolist()
    dest, flipped = list(range(dim)), []
    for i in range(dim // 2):
        a, b = order[2 * i], order[2 * i + 1]
        dest[a], dest[b] = b, a
        if signs[i]:
            flipped += [a, b]
    if dim % 2 == 1 and signs[dim // 2]:
        flipped.append(order[-1])
    return dest, sorted(flipped)


SYMMETRY_VARIANTS = ("augmentation", "mirror_loss", "transition_mirroring", "architecture")
SYMMETRY_SHAPE = dict(N=32, T=6, obs=19, act=5, hidden=(64, 32, 128))
