from .modules import Actor, ActorFactory, Mlp, MlpFactory, Module, NormalDist, NormalDistFactory, Value, ValueFactory

__all__ = ["Actor", "ActorFactory", "Mlp", "MlpFactory", "Module", "NormalDist", "NormalDistFactory", "Value", "ValueFactory"]
