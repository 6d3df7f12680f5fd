"""ctypes binding of ``libcusrl_b200.so`` (the C ABI declared in ``include/cusrl_b200.h``).

There is deliberately NO fallback: if the library is missing the import of any product module that
needs a kernel raises, and every wrapper raises ``RuntimeError`` on a non-zero return code.
"""

from __future__ import annotations

import ctypes
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libcusrl_b200.so"
HEADER_PATH = PKG_DIR.parent / "include" / "cusrl_b200.h"


class GatherField(ctypes.Structure):
    """Mirror of ``cusrl_b200_gather_field``."""

    _fields_ = [
        ("src", c_void_p),
        ("dst", c_void_p),
        ("row_bytes", c_int64),
        ("src_stride", c_int64),
        ("dst_stride", c_int64),
    ]


P = c_void_p  # every device pointer crosses the ABI as void*

_SIGNATURES: dict[str, tuple[object, list[object]]] = {
    "cusrl_b200_abi_version": (c_int, []),
    "cusrl_b200_last_error": (c_char_p, []),
    "cusrl_b200_sm_count": (c_int, []),
    "cusrl_b200_next_value_f32": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int64, c_float, P]),
    "cusrl_b200_gae_f32": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int64, c_double, c_double, c_double, P]),
    "cusrl_b200_gae_fused_f32": (
        c_int,
        [P, P, P, P, P, c_float, P, P, P, c_int64, c_int64, c_int64, c_double, c_double, c_double, P],
    ),
    "cusrl_b200_gae_chain_supported": (c_int, [c_int64, c_int64]),
    "cusrl_b200_gae_chain_scratch_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_gae_set_chain_threads": (c_int, [c_int]),
    "cusrl_b200_gae_chain_f32": (
        c_int, [P, P, P, P, P, c_float, P, P, P, c_int64, c_int64, c_double, c_double, c_double, P, P, c_size_t, P]),
    "cusrl_b200_gae_set_config": (c_int, [c_int, c_int]),
    "cusrl_b200_gae_set_variant": (c_int, [c_int, c_int, c_int, c_int]),
    "cusrl_b200_gae_set_schedule": (c_int, [c_int]),
    "cusrl_b200_advantage_stats_scratch_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_advantage_stats_f32": (c_int, [P, c_int64, c_int64, P, P, c_size_t, P]),
    "cusrl_b200_advantage_normalize_f32": (c_int, [P, c_int64, c_int64, P, c_float, P]),
    "cusrl_b200_merge_mean_var_f32": (c_int, [P, c_int64, c_int64, P, P]),
    "cusrl_b200_gather_rows": (c_int, [POINTER(GatherField), c_int, P, c_int64, c_int64, P]),
    "cusrl_b200_ppo_loss_scratch_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_ppo_loss_f32": (
        c_int,
        [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int]
        + [c_float] * 5
        + [P] * 10
        + [P, c_size_t, P],
    ),
    "cusrl_b200_scale_f32": (c_int, [P, c_int64, P, P]),
    "cusrl_b200_act_grad_mul_f32": (c_int, [P, P, P, c_int64, c_int, P]),
    "cusrl_b200_policy_stats_scratch_bytes": (c_size_t, []),
    "cusrl_b200_policy_stats_f32": (c_int, [P] * 7 + [c_int64, c_int64, P, P, c_size_t, P]),
    "cusrl_b200_rnd_scratch_bytes": (c_size_t, []),
    "cusrl_b200_rnd_reward_f32": (c_int, [P, P, c_int64, c_int64, c_float, P, c_int64, P, P, P, c_size_t, P]),
    "cusrl_b200_mse_f32": (c_int, [P, P, c_int64, c_int64, P, P, P, c_size_t, P]),
    "cusrl_b200_lstm_cell_fwd_f32": (c_int, [P, c_int64, P, P, P, P, P, P, P, P, c_int64, c_int64, P]),
    "cusrl_b200_lstm_cell_bwd_f32": (c_int, [P, c_int64, P, P, P, P, P, P, P, P, c_int64, c_int64, P]),
    "cusrl_b200_grad_sumsq_f32": (c_int, [P, c_int64, P, P]),
    "cusrl_b200_clip_coef_f32": (c_int, [P, c_float, P, P, P]),
    "cusrl_b200_adam_step_f32": (c_int, [P, P, P, P, c_int64, P] + [c_float] * 5 + [c_int64, P]),
    "cusrl_b200_adam_step_dev_f32": (c_int, [P, P, P, P, c_int64, P, P, P] + [c_float] * 4 + [P]),
    "cusrl_b200_weight_prep_f32": (c_int, [P, c_int64, c_int64, P, P, c_int64, P, P, c_int64, P]),
    "cusrl_b200_linear_fwd_tf32": (c_int, [P, c_int64, P, P, c_int64, P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "cusrl_b200_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "cusrl_b200_linear_wgrad_tf32": (
        c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int64, c_int, c_int, P, c_size_t, P]),
    "cusrl_b200_head_fwd_f32": (c_int, [P, c_int64, P, P, P, c_int64, c_int64, c_int64, P]),
    "cusrl_b200_head_bwd_scratch_bytes": (c_size_t, [c_int64, c_int64]),
    "cusrl_b200_head_bwd_f32": (
        c_int, [P, P, c_int64, P, c_int, P, c_int64, P, P, c_int64, c_int64, c_int64, c_int, P, c_int, P, c_size_t, P]),
    "cusrl_b200_linear_dgrad_tf32": (
        c_int, [P, c_int64, P, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, P, c_int, P,
                c_size_t, P]),
    "cusrl_b200_dgrad_workspace_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_colsum_workspace_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_colsum_f32": (c_int, [P, c_int64, c_int64, c_int64, P, c_int, P, c_size_t, P]),
    "cusrl_b200_amax_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P]),
    "cusrl_b200_split_f16": (c_int, [P, c_int64, c_int64, c_int64, P, P, P, c_int64, P]),
    "cusrl_b200_weight_prep_f16": (c_int, [P, c_int64, c_int64, P, P, P, c_int64, P, P, c_int64, P, P]),
    "cusrl_b200_gather_split_f16": (c_int, [P, c_int64, P, c_int64, c_int64, c_int64, P, P, P, c_int64, P]),
    "cusrl_b200_head_bwd_f16pair": (
        c_int, [P, P, P, c_int64, P, c_int, P, P, c_int64, P, P, P, c_int64, c_int64, c_int64, c_int, P, c_int, P, c_size_t, P]),
    "cusrl_b200_f16x3_set_prefetch": (c_int, [c_int]),
    "cusrl_b200_f16x3_set_tile": (c_int, [c_int]),
    "cusrl_b200_linear_fwd_f16x3": (
        c_int, [P, P, c_int64, P, P, P, c_int64, P, P, P, c_int64, P, P, c_int64, P, c_int64, c_int64, c_int64, c_int, P]),
    "cusrl_b200_linear_dgrad_f16x3": (
        c_int, [P, P, c_int64, P, P, P, c_int64, P, P, P, c_int64, P, P, c_int64, P, P, c_int64, P, c_int64, c_int64, c_int64,
                c_int, P, c_int, P, c_size_t, P]),
    "cusrl_b200_wgrad_f16x3_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "cusrl_b200_linear_wgrad_f16x3": (
        c_int, [P, P, c_int64, P, P, P, c_int64, P, P, c_int64, c_int64, c_int64, c_int64, c_int, P, c_size_t, P]),
    "cusrl_b200_column_stats_scratch_bytes": (c_size_t, [c_int64]),
    "cusrl_b200_column_stats_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_size_t, P]),
    "cusrl_b200_rms_merge_f32": (c_int, [P, P, P, P, P, c_int64, c_double, c_double, c_float, P]),
    "cusrl_b200_rms_normalize_f32": (c_int, [P, c_int64, P, c_int64, c_int64, c_int64, P, P, c_float, c_int, P]),
    "cusrl_b200_copy_rows_padded_f32": (c_int, [P, c_int64, P, c_int64, c_int64, c_int64, P]),
    "cusrl_b200_rollout_store_step_f32": (
        c_int, [P, c_int64, P, c_int64, c_int64, P, c_int64, P, c_int64, c_int64, P, P, c_int64, P, P, P, P, P, c_int64, P]),
    "cusrl_b200_sample_logp_f32": (c_int, [P, P, P, c_int64, c_int64, c_int, P, P, P, P]),
    "cusrl_b200_lstm_seq_supported": (c_int, [c_int64]),
    "cusrl_b200_lstm_seq_set_debug": (c_int, [c_int]),
    "cusrl_b200_lstm_seq_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "cusrl_b200_lstm_seq_fwd_f32": (
        c_int, [P, c_int64, P, P, c_int64, P, P, P, P, P, c_int64, P, P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int64, P, c_size_t, P]),
    "cusrl_b200_lstm_seq_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "cusrl_b200_lstm_seq_bwd_f32": (
        c_int, [P, c_int64, P, P, P, P, P, P, c_int64, P, P, c_int64, c_int64, c_int64, P, c_size_t, P]),
    "cusrl_b200_memory_reset_store_f32": (c_int, [P, P, P, c_int64, P, c_int64, c_int64, P]),
    "cusrl_b200_weight_prep_f16_multi": (c_int, [c_int64, P, P, P, P, P, P, P, P, P, P, P, P]),
    "cusrl_b200_mirror_rows_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int64, P, c_int64, c_int64, c_int64, P]),
}


def declared_symbols() -> list[str]:
    """Every function name declared in ``include/cusrl_b200.h`` (used by the CPU symbol test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cusrl_b200_[a-z0-9_]+)\s*\(", text)))


_lib: ctypes.CDLL | None = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach argument/return types."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m cusrl_b200.build` "
            "(cusrl_b200 has no CPU or PyTorch fallback for its kernels)"
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cusrl_b200_abi_version() != 1:
        raise RuntimeError("libcusrl_b200.so ABI version mismatch; rebuild with `python -m cusrl_b200.build -f`")
    variant = _gae_variant_from_env()
    if lib.cusrl_b200_gae_set_variant(*variant) != 0:
        raise RuntimeError(f"CUSRL_B200_GAE_VARIANT={variant} is not a valid (variant, warps, stages, ctas_per_sm)")
    if lib.cusrl_b200_gae_set_schedule(GAE_DEFAULT_SCHEDULE) != 0:
        raise RuntimeError(f"GAE_DEFAULT_SCHEDULE={GAE_DEFAULT_SCHEDULE} is not a valid schedule")
    _lib = lib
    return lib


# Kernel variant of the GAE scan installed at load time: (variant, warps, stages, ctas_per_sm) of
# cusrl_b200_gae_set_variant.  Chosen from the round-1 sweep on a B200 (profiles/); CUSRL_B200_GAE_VARIANT="v,w,s,c"
# overrides it for experiments.  Both variants are bit-identical.
GAE_DEFAULT_VARIANT = (0, 0, 2, 2)
# Instruction schedule of the register-resident GAE kernel (cusrl_b200_gae_set_schedule): 0 = chunked kernel, 1 =
# exact-length kernel.  Measured on a B200 (profiles/r01_kbench_gae_variants.jsonl): 9.2-9.3 us vs 9.5 us -> 0.
GAE_DEFAULT_SCHEDULE = int(__import__("os").environ.get("CUSRL_B200_GAE_SCHEDULE", "0"))


def _gae_variant_from_env() -> tuple[int, int, int, int]:
    import os

    raw = os.environ.get("CUSRL_B200_GAE_VARIANT")
    if not raw:
        return GAE_DEFAULT_VARIANT
    parts = tuple(int(x) for x in raw.split(","))
    if len(parts) != 4:
        raise RuntimeError("CUSRL_B200_GAE_VARIANT must be 'variant,warps,stages,ctas_per_sm'")
    return parts  # type: ignore[return-value]


KERNEL_LAUNCHES = 0  # launches of cusrl_b200 kernels issued through this binding (bench.py's gpu_launches)


def check(code: int, what: str, launches: int = 1) -> None:
    """Raise like the reference's hooks do (ValueError for bad arguments, RuntimeError otherwise)."""
    global KERNEL_LAUNCHES
    if code == 0:
        KERNEL_LAUNCHES += launches
        return
    msg = load().cusrl_b200_last_error().decode(errors="replace")
    if code == -1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed with code {code}: {msg}")
