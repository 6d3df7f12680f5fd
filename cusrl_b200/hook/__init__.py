from .auxiliary import RandomNetworkDistillation
from .mdp import ObservationNanToNum, ObservationNormalization
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
    "EntropyLoss",
    "GeneralizedAdvantageEstimation",
    "GradientClipping",
    "ModuleInitialization",
    "ObservationNanToNum",
    "ObservationNormalization",
    "OnPolicyPreparation",
    "OnPolicyStatistics",
    "PpoSurrogateLoss",
    "RandomNetworkDistillation",
    "ValueComputation",
    "ValueLoss",
]
