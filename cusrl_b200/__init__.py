"""cusrl_b200 -- B200-native (sm_100a) implementation of CusRL's vectorised on-policy PPO hot path.

Python/PyTorch host code mirroring the reference's Agent / Hook / Sampler plugin surface over a thin C-ABI
library of hand-written CUDA kernels (``include/cusrl_b200.h``).  See DESIGN.md and INTEGRATION.md.
"""

from . import distributed, hook, nn, ops, preset
from .environment import EnvironmentSpec, SyntheticEnvironment
from .hook import *  # noqa: F401,F403
from .nn import Actor, Mlp, NormalDist, Rnn, Value
from .preset import PpoAgentFactory, RecurrentPpoAgentFactory, anymal_c_rough_ppo, ppo_hook_suite
from .runtime import CONFIG, device
from .sampler import AutoMiniBatchSampler, MiniBatchSampler, TemporalMiniBatchSampler
from .template import ActorCritic, ActorCriticFactory, AdamFactory, Buffer, FlatAdam, Hook, HookComposite, HookList, Sampler
from .trainer import Trainer

__version__ = "0.1.0"
