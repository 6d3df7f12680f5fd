from .auxiliary import RandomNetworkDistillation
from .control import EmptyCudaCache
from .mdp import ObservationNanToNum, ObservationNormalization
from .symmetry import (
    MirrorDef,
    MirrorSymmetryLoss,
    SymmetricActor,
    SymmetricArchitecture,
    SymmetricDataAugmentation,
    TransitionMirroring,
)
from .on_policy import (
    AdaptiveLRSchedule,
    AdvantageNormalization,
    AdvantageReduction,
    EntropyLoss,
    GeneralizedAdvantageEstimation,
    GradientClipping,
    ModuleInitialization,
    OnPolicyPreparation,
    OnPolicyStatistics,
    PpoSurrogateLoss,
    ValueComputation,
    ValueLoss,
)

__all__ = [
    "AdaptiveLRSchedule",
    "AdvantageNormalization",
    "AdvantageReduction",
    "EmptyCudaCache",
    "EntropyLoss",
    "GeneralizedAdvantageEstimation",
    "GradientClipping",
    "MirrorDef",
    "MirrorSymmetryLoss",
    "ModuleInitialization",
    "ObservationNanToNum",
    "ObservationNormalization",
    "OnPolicyPreparation",
    "OnPolicyStatistics",
    "PpoSurrogateLoss",
    "RandomNetworkDistillation",
    "SymmetricActor",
    "SymmetricArchitecture",
    "SymmetricDataAugmentation",
    "TransitionMirroring",
    "ValueComputation",
    "ValueLoss",
]
