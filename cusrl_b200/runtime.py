"""Process-level runtime: rank / device discovery from the torchrun environment and the lazily
initialised process group.  Mirrors the reference's ``CONFIG`` singleton and ``configure_distributed``
(cusrl/utils/config.py:31-38,160-187): one process per GPU, ``cuda:{LOCAL_RANK}``, NCCL on CUDA and
Gloo on CPU (used only by the host-logic tests)."""

from __future__ import annotations

import atexit
import os

import torch

__all__ = ["CONFIG", "configure_distributed", "device"]


class _Config:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.distributed = self.world_size > 1
        self.seed: int | None = None
        self._device: torch.device | None = None
        self._pg_ready = False

    @property
    def device(self) -> torch.device:
        if self._device is None:
            if torch.cuda.is_available():
                self._device = torch.device("cuda", self.local_rank % max(torch.cuda.device_count(), 1))
                torch.cuda.set_device(self._device)
            else:
                self._device = torch.device("cpu")
        return self._device

    @device.setter
    def device(self, value):
        self._device = torch.device(value)
        if self._device.type == "cuda":
            torch.cuda.set_device(self._device)


CONFIG = _Config()


def device(value: str | torch.device | None = None) -> torch.device:
    """Resolve an optional device argument against the process default (cusrl.device)."""
    return CONFIG.device if value is None else torch.device(value)


def configure_distributed() -> bool:
    """Initialise the default process group on first use; returns whether distributed mode is on.
    Rendezvous comes from the torchrun environment (MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE)."""
    if not CONFIG.distributed:
        return False
    if not CONFIG._pg_ready:
        if not torch.distributed.is_initialized():
            backend = "nccl" if CONFIG.device.type == "cuda" else "gloo"
            kwargs = {"device_id": CONFIG.device} if backend == "nccl" else {}
            torch.distributed.init_process_group(backend=backend, **kwargs)
            atexit.register(_shutdown)
        CONFIG._pg_ready = True
    return True


def _shutdown():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        try:
            torch.distributed.destroy_process_group()
        except Exception:
            pass
