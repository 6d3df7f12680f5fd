from .auxiliary import RandomNetworkDistillation
from .on_policy import (
    AdaptiveLRSchedule,
    AdvantageNormalization,
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
    "EntropyLoss",
    "GeneralizedAdvantageEstimation",
    "GradientClipping",
    "ModuleInitialization",
    "OnPolicyPreparation",
    "OnPolicyStatistics",
    "PpoSurrogateLoss",
    "RandomNetworkDistillation",
    "ValueComputation",
    "ValueLoss",
]
