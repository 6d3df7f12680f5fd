from .actor_critic import ActorCritic, ActorCriticFactory, HookList
from .buffer import Buffer, Sampler
from .hook import Hook, HookComposite
from .optimizer import AdamFactory, FlatAdam, ParamArena

__all__ = ["ActorCritic", "ActorCriticFactory", "AdamFactory", "Buffer", "FlatAdam", "Hook", "HookComposite",
           "HookList", "ParamArena", "Sampler"]
