from .modules import Actor, ActorFactory, Mlp, MlpFactory, Module, NormalDist, NormalDistFactory, Value, ValueFactory
from .recurrent import Rnn, RnnFactory
from .rms import RunningMeanStd

__all__ = ["Actor", "ActorFactory", "Mlp", "MlpFactory", "Module", "NormalDist", "NormalDistFactory", "Rnn", "RnnFactory", "RunningMeanStd",
           "Value", "ValueFactory"]
