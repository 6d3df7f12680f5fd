from .modules import Actor, ActorFactory, Mlp, MlpFactory, Module, NormalDist, NormalDistFactory, Value, ValueFactory
from .recurrent import Rnn, RnnFactory

__all__ = ["Actor", "ActorFactory", "Mlp", "MlpFactory", "Module", "NormalDist", "NormalDistFactory", "Rnn", "RnnFactory",
           "Value", "ValueFactory"]
