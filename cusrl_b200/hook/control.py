"""Control hooks the PPO preset can switch on (reference cusrl/hook/control/).  The by-name schedules and the other control
hooks of the reference are used as they are (``HookComposite`` accepts the reference's own hook objects); this one is here
because ``ppo_hook_suite(empty_cuda_cache=True)`` (preset/ppo.py:34,63) must work without the reference on the path."""

from __future__ import annotations

import torch

from ..template.hook import Hook

__all__ = ["EmptyCudaCache"]


class EmptyCudaCache(Hook):
    """Returns the caching allocator's unused blocks to the driver after each update (control/empty_cuda_cache.py:8-13).
    Memory a captured train-step graph owns lives in the graph's private pool and is not affected."""

    def post_update(self):
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
