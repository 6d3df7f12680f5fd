"""Thin Python wrappers: torch CUDA tensors -> raw pointers -> C ABI (``include/cusrl_b200.h``).

PyTorch is plumbing here (device memory, streams).  Every function launches on the current CUDA
stream, never synchronises, and refuses non-CUDA tensors: there is no CPU path.
"""

from __future__ import annotations

import ctypes

import weakref

import torch

from . import _lib
from ._lib import GatherField

__all__ = [
    "next_value",
    "gae",
    "gae_fused",
    "gae_chain",
    "advantage_stats",
    "advantage_normalize_",
    "merge_mean_var",
    "gather_rows",
    "ppo_loss",
    "scale_",
    "policy_stats",
    "grad_sumsq_",
    "clip_coef",
    "adam_step_",
    "adam_step_dev_",
]

_scratch: dict[tuple[int, str], torch.Tensor] = {}
# Workspaces that were outgrown.  They are never handed back to the allocator: a captured CUDA graph (template/graphs.py)
# has the pointer of the workspace it was captured with baked into its kernel arguments, and replaying it after the
# allocator reused that memory would corrupt whatever lives there now.  Workspaces are small (<= a few MB) and grow a
# handful of times per process.
_retired_scratch: list[torch.Tensor] = []
CHECK_INDICES = __import__("os").environ.get("CUSRL_B200_CHECK_INDICES", "0") not in ("", "0")


GAE_DEFAULT_VARIANT = _lib.GAE_DEFAULT_VARIANT
GAE_DEFAULT_SCHEDULE = _lib.GAE_DEFAULT_SCHEDULE


def launch_count() -> int:
    """Number of cusrl_b200 kernel launches issued so far by this process."""
    return _lib.KERNEL_LAUNCHES


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter is ~30x cheaper than building a
    torch.cuda.Stream object, and this is called once per kernel launch)."""
    if _raw_stream is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    """Every tensor that crosses the C ABI must live on the GPU: there is no CPU or PyTorch fallback for any kernel."""
    if not t.is_cuda:
        raise RuntimeError(f"cusrl_b200: '{name}' must be a CUDA tensor (no CPU fallback exists)")


def _ptr(t: torch.Tensor | None, dtype: torch.dtype | None = None, name: str = "tensor") -> int | None:
    if t is None:
        return None
    _require_cuda(t, name)
    if not t.is_contiguous():
        raise ValueError(f"cusrl_b200: '{name}' must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"cusrl_b200: '{name}' must have dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def _flag_ptr(t: torch.Tensor, name: str) -> int:
    if t.dtype not in (torch.bool, torch.uint8):
        raise TypeError(f"'{name}' must have dtype bool")
    return _ptr(t, None, name)


def _get_scratch(device: torch.device, key: str, nbytes: int) -> torch.Tensor:
    k = (device.index or 0, key)
    buf = _scratch.get(k)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired_scratch.append(buf)
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _scratch[k] = buf
    return buf


def _tnd(t: torch.Tensor) -> tuple[int, int, int]:
    if t.dim() != 3:
        raise ValueError(f"expected a [T, N, Dv] tensor, got shape {tuple(t.shape)}")
    return t.shape[0], t.shape[1], t.shape[2]


# ---------------------------------------------------------------------------------------------- K3
def next_value(
    value: torch.Tensor,
    terminated: torch.Tensor,
    truncated: torch.Tensor,
    boot_value: torch.Tensor,
    termination_value: float = 0.0,
    trunc_value: torch.Tensor | None = None,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """ValueComputation.pre_update arithmetic (reference hook/on_policy/value.py:68-82)."""
    T, N, Dv = _tnd(value)
    out = torch.empty_like(value) if out is None else out
    code = _lib.load().cusrl_b200_next_value_f32(
        _ptr(value, torch.float32, "value"),
        _flag_ptr(terminated, "terminated"),
        _flag_ptr(truncated, "truncated"),
        _ptr(boot_value, torch.float32, "boot_value"),
        _ptr(trunc_value, torch.float32, "trunc_value"),
        _ptr(out, torch.float32, "next_value"),
        T, N, Dv, float(termination_value), _stream(),
    )  # fmt: skip
    _lib.check(code, "next_value")
    return out


# ---------------------------------------------------------------------------------------------- K1
def gae(
    reward: torch.Tensor,
    done: torch.Tensor,
    value: torch.Tensor,
    next_value: torch.Tensor,
    gamma: float,
    lamda: float,
    lamda_value: float | None = None,
    advantage: torch.Tensor | None = None,
    ret: torch.Tensor | None = None,
    compute_return: bool = True,
) -> tuple[torch.Tensor, torch.Tensor | None]:
    """GAE advantage (+return) scan (reference hook/on_policy/gae.py:8-20,85-110). Bit-exact."""
    T, N, Dv = _tnd(value)
    advantage = torch.empty_like(value) if advantage is None else advantage
    if compute_return and ret is None:
        ret = torch.empty_like(value)
    code = _lib.load().cusrl_b200_gae_f32(
        _ptr(reward, torch.float32, "reward"),
        _flag_ptr(done, "done"),
        _ptr(value, torch.float32, "value"),
        _ptr(next_value, torch.float32, "next_value"),
        _ptr(advantage, torch.float32, "advantage"),
        _ptr(ret, torch.float32, "return"),
        T, N, Dv, float(gamma), float(lamda), -1.0 if lamda_value is None else float(lamda_value), _stream(),
    )  # fmt: skip
    _lib.check(code, "gae")
    return advantage, ret


def gae_fused(
    reward: torch.Tensor,
    terminated: torch.Tensor,
    truncated: torch.Tensor,
    value: torch.Tensor,
    boot_value: torch.Tensor,
    gamma: float,
    lamda: float,
    lamda_value: float | None = None,
    termination_value: float = 0.0,
    next_value_out: torch.Tensor | None = None,
    advantage: torch.Tensor | None = None,
    ret: torch.Tensor | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """K1 with K3 folded in: next_value is formed on the fly (value.py:68-82 + gae.py:8-20)."""
    T, N, Dv = _tnd(value)
    advantage = torch.empty_like(value) if advantage is None else advantage
    ret = torch.empty_like(value) if ret is None else ret
    code = _lib.load().cusrl_b200_gae_fused_f32(
        _ptr(reward, torch.float32, "reward"),
        _flag_ptr(terminated, "terminated"),
        _flag_ptr(truncated, "truncated"),
        _ptr(value, torch.float32, "value"),
        _ptr(boot_value, torch.float32, "boot_value"),
        float(termination_value),
        _ptr(next_value_out, torch.float32, "next_value"),
        _ptr(advantage, torch.float32, "advantage"),
        _ptr(ret, torch.float32, "return"),
        T, N, Dv, float(gamma), float(lamda), -1.0 if lamda_value is None else float(lamda_value), _stream(),
    )  # fmt: skip
    _lib.check(code, "gae_fused")
    return advantage, ret


def gae_chain_supported(T: int, Dv: int) -> bool:
    return bool(_lib.load().cusrl_b200_gae_chain_supported(T, Dv))


def gae_chain(
    reward: torch.Tensor,
    terminated: torch.Tensor,
    truncated: torch.Tensor,
    value: torch.Tensor,
    boot_value: torch.Tensor,
    gamma: float,
    lamda: float,
    lamda_value: float | None,
    termination_value: float,
    next_value_out: torch.Tensor | None,
    advantage: torch.Tensor,
    ret: torch.Tensor | None,
    mean_var: torch.Tensor | None = None,
) -> torch.Tensor:
    """K3 + K1 + the K2 statistics in one launch (value.py:68-82, gae.py:8-20,85-110, advantage.py:110-111), Dv == 1.
    Returns mean_var = [mean | unbiased var] of the advantages."""
    T, N, Dv = _tnd(value)
    if Dv != 1:
        raise ValueError("gae_chain: value_dim must be 1")
    lib = _lib.load()
    scratch = _get_scratch(value.device, "gaechain", lib.cusrl_b200_gae_chain_scratch_bytes(N))
    mean_var = torch.empty(2, dtype=torch.float32, device=value.device) if mean_var is None else mean_var
    f32 = torch.float32
    code = lib.cusrl_b200_gae_chain_f32(
        _ptr(reward, f32, "reward"), _flag_ptr(terminated, "terminated"), _flag_ptr(truncated, "truncated"),
        _ptr(value, f32, "value"), _ptr(boot_value, f32, "boot_value"), float(termination_value),
        _ptr(next_value_out, f32, "next_value"), _ptr(advantage, f32, "advantage"), _ptr(ret, f32, "return"),
        T, N, float(gamma), float(lamda), -1.0 if lamda_value is None else float(lamda_value),
        _ptr(mean_var, f32, "mean_var"), scratch.data_ptr(), scratch.numel(), _stream())
    _lib.check(code, "gae_chain", launches=2)
    return mean_var


# ---------------------------------------------------------------------------------------------- K2
def advantage_stats(advantage: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[mean(Dv) | unbiased var(Dv)] over all but the last dim (advantage.py:110-111)."""
    Dv = advantage.shape[-1]
    E = advantage.numel() // Dv
    lib = _lib.load()
    nbytes = lib.cusrl_b200_advantage_stats_scratch_bytes(Dv)
    scratch = _get_scratch(advantage.device, "advstats", nbytes)
    out = torch.empty(2 * Dv, dtype=torch.float32, device=advantage.device) if out is None else out
    code = lib.cusrl_b200_advantage_stats_f32(
        _ptr(advantage, torch.float32, "advantage"), E, Dv, _ptr(out, torch.float32, "mean_var"),
        scratch.data_ptr(), scratch.numel(), _stream(),
    )  # fmt: skip
    _lib.check(code, "advantage_stats", launches=2)
    return out


def advantage_normalize_(advantage: torch.Tensor, mean_var: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """In-place (adv - mean) / sqrt(var + eps) (advantage.py:114-115)."""
    Dv = advantage.shape[-1]
    E = advantage.numel() // Dv
    code = _lib.load().cusrl_b200_advantage_normalize_f32(
        _ptr(advantage, torch.float32, "advantage"), E, Dv, _ptr(mean_var, torch.float32, "mean_var"),
        float(eps), _stream(),
    )  # fmt: skip
    _lib.check(code, "advantage_normalize")
    return advantage


def merge_mean_var(gathered: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Equal-weight cross-rank merge of stacked [mean | var] rows (utils/distributed.py:175-183)."""
    W, two_dv = gathered.shape
    out = torch.empty(two_dv, dtype=torch.float32, device=gathered.device) if out is None else out
    code = _lib.load().cusrl_b200_merge_mean_var_f32(
        _ptr(gathered, torch.float32, "gathered"), W, two_dv // 2, _ptr(out, torch.float32, "mean_var"), _stream()
    )
    _lib.check(code, "merge_mean_var")
    return out


# ---------------------------------------------------------------------------------------------- K8
def gather_rows(
    fields: list[tuple[torch.Tensor, torch.Tensor]],
    index: torch.Tensor,
) -> None:
    """For every (src [E, ...], dst [n, ...]) pair: dst[i] = src[index[i]] in ONE launch.

    Rows may be padded: only the last dim may be non-dense (stride(0) >= row payload); the padding
    bytes of every destination row are written as zeros. Reference: mini_batch_sampler.py:76,89.
    """
    n_src = None
    if CHECK_INDICES and index.numel():
        # debug aid (one host sync): the kernel clamps out-of-range indices instead of faulting
        lo, hi = int(index.min()), int(index.max())
        limit = fields[0][0].shape[0] if fields else 0
        if lo < 0 or hi >= limit:
            raise IndexError(f"gather_rows: indices span [{lo}, {hi}] but the sources have {limit} rows")
    arr = (GatherField * len(fields))()
    for k, (src, dst) in enumerate(fields):
        _require_cuda(src, "gather_rows source")
        _require_cuda(dst, "gather_rows destination")
        if src.dtype != dst.dtype:
            raise TypeError("gather_rows: src/dst dtype mismatch")
        s2 = src.reshape(src.shape[0], -1) if src.is_contiguous() else src
        d2 = dst.reshape(dst.shape[0], -1) if dst.is_contiguous() else dst
        if s2.dim() != 2 or d2.dim() != 2 or s2.stride(1) != 1 or d2.stride(1) != 1:
            raise ValueError("gather_rows: fields must be [rows, width] with a dense last dim")
        if d2.shape[0] != index.numel():
            raise ValueError("gather_rows: destination row count must equal the number of indices")
        item = src.element_size()
        row_bytes = min(s2.shape[1], d2.shape[1]) * item
        n_src = s2.shape[0] if n_src is None else n_src
        if s2.shape[0] != n_src:
            raise ValueError("gather_rows: all sources must have the same number of rows")
        arr[k] = GatherField(s2.data_ptr(), d2.data_ptr(), row_bytes, s2.stride(0) * item, d2.stride(0) * item)
    code = _lib.load().cusrl_b200_gather_rows(
        arr, len(fields), _ptr(index, torch.int64, "index"), index.numel(), n_src, _stream()
    )
    _lib.check(code, "gather_rows")


# ---------------------------------------------------------------------------------------------- K4
def ppo_loss(
    mean: torch.Tensor,
    std: torch.Tensor,
    action: torch.Tensor,
    logp_old: torch.Tensor,
    advantage: torch.Tensor,
    ret: torch.Tensor | None,
    value_old: torch.Tensor | None,
    curr_value: torch.Tensor | None,
    clip_ratio: float,
    w_surrogate: float,
    w_entropy: float,
    w_value: float,
    value_clip: float | None = None,
    want_per_sample: bool = True,
    want_grads: bool = True,
) -> dict[str, torch.Tensor]:
    """Fused PPO objective + unit gradients (see ``cusrl_b200_ppo_loss_f32`` in the header)."""
    B, A = mean.shape
    dev = mean.device
    has_value = curr_value is not None
    Dv = curr_value.shape[-1] if has_value else 1
    f32 = torch.float32
    out: dict[str, torch.Tensor] = {
        "losses": torch.empty(3, dtype=f32, device=dev),
        "metrics": torch.empty(3, dtype=f32, device=dev),
    }
    if want_per_sample:
        for k in ("logp", "entropy", "logp_ratio", "prob_ratio"):
            out[k] = torch.empty(B, 1, dtype=f32, device=dev)
    if want_grads:
        out["d_mean"] = torch.empty(B, A, dtype=f32, device=dev)
        out["d_std_surr"] = torch.empty(A, dtype=f32, device=dev)
        out["d_std_ent"] = torch.empty(A, dtype=f32, device=dev)
        if has_value:
            out["d_value"] = torch.empty(B, Dv, dtype=f32, device=dev)
    lib = _lib.load()
    scratch = _get_scratch(dev, "ppoloss", lib.cusrl_b200_ppo_loss_scratch_bytes(A))
    g = out.get
    code = lib.cusrl_b200_ppo_loss_f32(
        _ptr(mean, f32, "mean"), _ptr(std, f32, "std"), _ptr(action, f32, "action"),
        _ptr(logp_old, f32, "action_logp"), _ptr(advantage, f32, "advantage"), _ptr(ret, f32, "return"),
        _ptr(value_old, f32, "value"), _ptr(curr_value, f32, "curr_value"),
        B, A, Dv, int(has_value),
        float(clip_ratio), float(w_surrogate), float(w_entropy), float(w_value),
        -1.0 if value_clip is None else float(value_clip),
        _ptr(g("logp")), _ptr(g("entropy")), _ptr(g("logp_ratio")), _ptr(g("prob_ratio")),
        _ptr(out["losses"]), _ptr(out["metrics"]),
        _ptr(g("d_mean")), _ptr(g("d_std_surr")), _ptr(g("d_std_ent")), _ptr(g("d_value")),
        scratch.data_ptr(), scratch.numel(), _stream(),
    )  # fmt: skip
    _lib.check(code, "ppo_loss", launches=2)
    return out


def scale_(x: torch.Tensor, scale_dev: torch.Tensor) -> torch.Tensor:
    """x *= scale (a 1-element device tensor) without reading the scalar on the host."""
    code = _lib.load().cusrl_b200_scale_f32(
        _ptr(x, torch.float32, "x"), x.numel(), _ptr(scale_dev, torch.float32, "scale"), _stream()
    )
    _lib.check(code, "scale")
    return x


def policy_stats(
    mean_old: torch.Tensor,
    std_old: torch.Tensor,
    mean_new: torch.Tensor,
    std_new: torch.Tensor,
    action: torch.Tensor,
    logp_old: torch.Tensor,
    advantage: torch.Tensor,
) -> torch.Tensor:
    """[mean KL(old||new), mean importance-weighted advantage, mean std] (stats.py:29-40)."""
    A = mean_old.shape[-1]
    E = mean_old.numel() // A
    lib = _lib.load()
    scratch = _get_scratch(mean_old.device, "polstats", lib.cusrl_b200_policy_stats_scratch_bytes())
    out = torch.empty(3, dtype=torch.float32, device=mean_old.device)
    f32 = torch.float32
    code = lib.cusrl_b200_policy_stats_f32(
        _ptr(mean_old, f32, "mean_old"), _ptr(std_old, f32, "std_old"), _ptr(mean_new, f32, "mean_new"),
        _ptr(std_new, f32, "std_new"), _ptr(action, f32, "action"), _ptr(logp_old, f32, "action_logp"),
        _ptr(advantage, f32, "advantage"), E, A, _ptr(out), scratch.data_ptr(), scratch.numel(), _stream(),
    )  # fmt: skip
    _lib.check(code, "policy_stats", launches=2)
    return out


# ---------------------------------------------------------------------------------------------- K9
def grad_sumsq_(grad: torch.Tensor, sumsq: torch.Tensor) -> torch.Tensor:
    """sumsq (1 double on device) += sum(grad**2)."""
    code = _lib.load().cusrl_b200_grad_sumsq_f32(
        _ptr(grad, torch.float32, "grad"), grad.numel(), _ptr(sumsq, torch.float64, "sumsq"), _stream()
    )
    _lib.check(code, "grad_sumsq")
    return sumsq


def clip_coef(sumsq: torch.Tensor, max_norm: float, norm: torch.Tensor, coef: torch.Tensor) -> None:
    """norm = sqrt(sumsq); coef = min(1, max_norm / (norm + 1e-6)) (torch clip_grad_norm_)."""
    code = _lib.load().cusrl_b200_clip_coef_f32(
        _ptr(sumsq, torch.float64, "sumsq"), float(max_norm), _ptr(norm, torch.float32, "norm"),
        _ptr(coef, torch.float32, "coef"), _stream(),
    )  # fmt: skip
    _lib.check(code, "clip_coef")


def adam_step_(
    param: torch.Tensor,
    grad: torch.Tensor,
    exp_avg: torch.Tensor,
    exp_avg_sq: torch.Tensor,
    step: int,
    lr: float,
    betas: tuple[float, float] = (0.9, 0.999),
    eps: float = 1e-8,
    weight_decay: float = 0.0,
    coef: torch.Tensor | None = None,
) -> None:
    """Fused (clip-scale +) Adam step on flat arenas (torch.optim.Adam semantics)."""
    f32 = torch.float32
    code = _lib.load().cusrl_b200_adam_step_f32(
        _ptr(param, f32, "param"), _ptr(grad, f32, "grad"), _ptr(exp_avg, f32, "exp_avg"),
        _ptr(exp_avg_sq, f32, "exp_avg_sq"), param.numel(), _ptr(coef, f32, "coef"),
        float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), _stream(),
    )  # fmt: skip
    _lib.check(code, "adam_step")


def adam_step_dev_(
    param: torch.Tensor,
    grad: torch.Tensor,
    exp_avg: torch.Tensor,
    exp_avg_sq: torch.Tensor,
    step_dev: torch.Tensor,
    lr_dev: torch.Tensor,
    betas: tuple[float, float] = (0.9, 0.999),
    eps: float = 1e-8,
    weight_decay: float = 0.0,
    coef: torch.Tensor | None = None,
) -> None:
    """:func:`adam_step_` with the step count (int64 [1]) and learning rate (float32 [1]) read from device memory, the
    form a captured CUDA graph can replay while both keep changing."""
    f32 = torch.float32
    code = _lib.load().cusrl_b200_adam_step_dev_f32(
        _ptr(param, f32, "param"), _ptr(grad, f32, "grad"), _ptr(exp_avg, f32, "exp_avg"),
        _ptr(exp_avg_sq, f32, "exp_avg_sq"), param.numel(), _ptr(coef, f32, "coef"), _ptr(lr_dev, f32, "lr_dev"),
        _ptr(step_dev, torch.int64, "step_dev"), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), _stream(),
    )  # fmt: skip
    _lib.check(code, "adam_step_dev")


# ---------------------------------------------------------------------------------------------- K6
# Dense layers on tcgen05 (csrc/gemm_tf32.cu, csrc/gemm_wgrad_tf32.cu) and SIMT heads (csrc/head_kernels.cu).
# 3 = 3xTF32 (fp32-equivalent), 2 = f16x3 (fp16 hi / lo split, fp32-equivalent, half the tensor time), 1 = single-pass TF32
GEMM_PRECISION = int(__import__("os").environ.get("CUSRL_B200_GEMM_PRECISION", "2"))


# f16x3 pays a few extra small launches per layer (range statistics, splits of the trunk inputs): below this many rows the
# 3xTF32 kernels (equally fp32-equivalent) are used instead.  Measured on a B200: 24 576-row minibatches (4096 envs) 27.1 ms
# per iteration with f16x3 against 22.9 ms with 3xTF32; 98 304 rows (16 384 envs) 59.2 against 61.2; 393 216 rows 135 against 180.
F16X3_MIN_ROWS = int(__import__("os").environ.get("CUSRL_B200_F16X3_MIN_ROWS", "49152"))


def tf32_passes() -> int:
    """`precision` argument for the tf32 kernels: layers the f16x3 path does not cover run as 3xTF32 when it is selected."""
    return 3 if GEMM_PRECISION == 2 else GEMM_PRECISION

_weights_epoch = 0
_weight_cache: dict[int, tuple[tuple, dict[str, torch.Tensor], "weakref.ref"]] = {}


def invalidate_weight_cache() -> None:
    """Parameters were rewritten through raw pointers (optimizer step / checkpoint load): operand copies are stale."""
    global _weights_epoch
    _weights_epoch += 1


def prepared_weight(w: torch.Tensor) -> dict[str, torch.Tensor]:
    """hi/lo (+ transposed) tensor-core operand copies of a weight matrix, rebuilt once per optimizer step."""
    key = w.data_ptr()
    stamp = (_weights_epoch, w._version, tuple(w.shape))
    hit = _weight_cache.get(key)
    # An entry is only trusted while the tensor object it was built from is alive: once that tensor is freed the
    # allocator may hand its address (with an equal version counter) to a different matrix.
    alive = hit is not None and hit[2]() is not None
    if alive and hit[0] == stamp:
        return hit[1]
    reuse = alive and hit[0][2] == stamp[2]
    wp = weight_prep(w, out=hit[1]) if reuse else weight_prep(w)
    owner = hit[2] if reuse else weakref.ref(w)
    _weight_cache[key] = (stamp, wp, owner)
    if len(_weight_cache) > 256:  # drop entries of freed tensors (tests create many short-lived matrices)
        for k in [k for k, v in _weight_cache.items() if v[2]() is None]:
            del _weight_cache[k]
    return wp


def act_grad_mul(dy: torch.Tensor, y: torch.Tensor, act: int) -> torch.Tensor:
    """dZ = dY * act'(Z) from the stored post-activation Y."""
    dy = dy.contiguous()
    if act == 0:
        return dy
    y = y.contiguous()
    dz = torch.empty_like(dy)
    code = _lib.load().cusrl_b200_act_grad_mul_f32(_ptr(dy, torch.float32, "dy"), _ptr(y, torch.float32, "y"),
                                                   dz.data_ptr(), dy.numel(), act, _stream())
    _lib.check(code, "act_grad_mul")
    return dz


# ---- tcgen05 dense-layer entry points (raw; the autograd-facing wrappers above will move onto these) ----
def weight_prep(w: torch.Tensor, transposed: bool = True, out: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor]:
    """hi/lo (and transposed hi/lo) operand copies of a weight matrix [N, K]; leading dims padded to 4 floats."""
    N, K = w.shape
    ld = (K + 3) // 4 * 4
    ldt = (N + 3) // 4 * 4
    if out is None:
        out = {"hi": torch.zeros(N, ld, device=w.device), "lo": torch.zeros(N, ld, device=w.device)}
        if transposed:
            out["hi_t"] = torch.zeros(K, ldt, device=w.device)
            out["lo_t"] = torch.zeros(K, ldt, device=w.device)
    transposed = "hi_t" in out
    code = _lib.load().cusrl_b200_weight_prep_f32(
        _ptr(w.detach(), torch.float32, "weight"), N, K, out["hi"].data_ptr(), out["lo"].data_ptr(), ld,
        out["hi_t"].data_ptr() if transposed else None, out["lo_t"].data_ptr() if transposed else None, ldt, _stream())
    _lib.check(code, "weight_prep")
    return out


def _rows(t: torch.Tensor, name: str) -> tuple[int, int]:
    """(data_ptr, leading dimension) of a 2-D tensor with unit inner stride."""
    _require_cuda(t, name)
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float32:
        raise ValueError(f"cusrl_b200: '{name}' must be a 2-D float32 tensor with a dense last dim")
    return t.data_ptr(), t.stride(0)


def tc_linear_fwd(x: torch.Tensor, wp: dict[str, torch.Tensor], bias: torch.Tensor | None, n_out: int, act: int,
                  precision: int = 3, out: torch.Tensor | None = None) -> torch.Tensor:
    """Y = act(X W^T + b) on tcgen05 (cusrl_b200_linear_fwd_tf32)."""
    M, K = x.shape
    xp, ldx = _rows(x, "x")
    y = torch.empty(M, n_out, device=x.device) if out is None else out
    yp, ldy = _rows(y, "y")
    code = _lib.load().cusrl_b200_linear_fwd_tf32(
        xp, ldx, wp["hi"].data_ptr(), wp["lo"].data_ptr(), wp["hi"].stride(0), _ptr(bias, torch.float32, "bias"),
        yp, ldy, M, n_out, K, act, precision, _stream())
    _lib.check(code, "linear_fwd")
    return y


def tc_linear_dgrad(dy: torch.Tensor, wp: dict[str, torch.Tensor], x_act: torch.Tensor | None, k_in: int, act: int,
                    precision: int = 3, out: torch.Tensor | None = None, db_below: torch.Tensor | None = None,
                    accumulate: bool = False) -> torch.Tensor:
    """dX = (dY W) * act'(x_act) on tcgen05 (cusrl_b200_linear_dgrad_tf32); `db_below` (+)= column sums of dX, the bias
    gradient of the layer below, produced by the same kernel's epilogue."""
    M, N = dy.shape
    dyp, lddy = _rows(dy, "dy")
    dx = torch.empty(M, k_in, device=dy.device) if out is None else out
    dxp, lddx = _rows(dx, "dx")
    xa, ldxa = (None, 0) if x_act is None else _rows(x_act, "x_act")
    lib = _lib.load()
    if db_below is not None:
        ws = _get_scratch(dy.device, "dgrad", lib.cusrl_b200_dgrad_workspace_bytes(k_in))
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    else:
        ws_ptr, ws_bytes = None, 0
    code = lib.cusrl_b200_linear_dgrad_tf32(
        dyp, lddy, wp["hi_t"].data_ptr(), wp["lo_t"].data_ptr(), wp["hi_t"].stride(0), xa, ldxa, dxp, lddx,
        M, N, k_in, act, precision, _ptr(db_below, torch.float32, "db_below"), int(accumulate), ws_ptr, ws_bytes, _stream())
    _lib.check(code, "linear_dgrad", launches=1 if db_below is None else 2)
    return dx


def tc_linear_wgrad(dz: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, db: torch.Tensor | None, precision: int = 3,
                    accumulate: bool = False) -> None:
    """dW (+)= dZ^T X and db (+)= colsum(dZ) on tcgen05 (cusrl_b200_linear_wgrad_tf32)."""
    M, N = dz.shape
    K = x.shape[1]
    dzp, lddz = _rows(dz, "dz")
    xp, ldx = _rows(x, "x")
    if not dw.is_contiguous() or dw.shape != (N, K):
        raise ValueError("tc_linear_wgrad: dw must be a contiguous [N, K] tensor")
    lib = _lib.load()
    ws = _get_scratch(dz.device, "wgrad", lib.cusrl_b200_wgrad_workspace_bytes(M, N, K))
    code = lib.cusrl_b200_linear_wgrad_tf32(dzp, lddz, xp, ldx, dw.data_ptr(), K, _ptr(db, torch.float32, "db"), M, N, K,
                                            precision, int(accumulate), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(code, "linear_wgrad", launches=4 if db is not None else 2)


def head_fwd(h: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Y = H W^T + b for a small-N fp32 head (cusrl_b200_head_fwd_f32); `out`: a dense [M, No] destination (e.g. the
    rollout buffer slot of the current step)."""
    M, K = h.shape
    No = w.shape[0]
    hp, ldh = _rows(h, "h")
    if out is None:
        y = torch.empty(M, No, device=h.device)
    else:
        y = out
        if y.shape != (M, No) or not y.is_contiguous() or y.dtype != torch.float32:
            raise ValueError("head_fwd: `out` must be a contiguous float32 [M, No] tensor")
        _require_cuda(y, "out")
    code = _lib.load().cusrl_b200_head_fwd_f32(hp, ldh, _ptr(w.detach(), torch.float32, "weight"),
                                               _ptr(None if b is None else b.detach(), torch.float32, "bias"),
                                               y.data_ptr(), M, K, No, _stream())
    _lib.check(code, "head_fwd")
    return y


def head_bwd(dy: torch.Tensor, h: torch.Tensor, w: torch.Tensor, act: int, dw: torch.Tensor, db: torch.Tensor | None,
             need_dh: bool = True, accumulate: bool = False, db_trunk: torch.Tensor | None = None,
             accumulate_trunk: bool = False) -> torch.Tensor | None:
    """dH = (dY W) * act'(H); dW (+)= dY^T H; db (+)= colsum(dY); db_trunk (+)= colsum(dH) (cusrl_b200_head_bwd_f32)."""
    M, K = h.shape
    No = w.shape[0]
    hp, ldh = _rows(h, "h")
    dh = torch.empty(M, K, device=h.device) if need_dh else None
    lib = _lib.load()
    scratch = _get_scratch(h.device, "headbwd", lib.cusrl_b200_head_bwd_scratch_bytes(K, No))
    code = lib.cusrl_b200_head_bwd_f32(
        _ptr(dy, torch.float32, "dy"), hp, ldh, _ptr(w.detach(), torch.float32, "weight"), act,
        None if dh is None else dh.data_ptr(), K, _ptr(dw, torch.float32, "dw"), _ptr(db, torch.float32, "db"),
        M, K, No, int(accumulate), _ptr(db_trunk, torch.float32, "db_trunk"), int(accumulate_trunk),
        scratch.data_ptr(), scratch.numel(), _stream())
    _lib.check(code, "head_bwd", launches=2)
    return dh


# ---------------------------------------------------------------------------------------------- K6, precision 2 (f16x3)
# Dense layers on tcgen05 kind::f16 with fp16 hi / lo split operands (csrc/f16x3_common.cuh, gemm_f16x3.cu,
# gemm_wgrad_f16x3.cu): fp32-equivalent products at half the tensor time of 3xTF32.
class Pair:
    """fp16 hi / lo split of an fp32 [rows, width] matrix: ``data[0]`` = hi, ``data[1]`` = lo (both [rows, ld], ld a multiple
    of 8 halves, columns >= width zero) and ``bound`` = a device scalar that is >= max|x| (it fixes the power-of-two scale)."""

    __slots__ = ("data", "bound", "width")

    def __init__(self, data: torch.Tensor, bound: torch.Tensor, width: int):
        self.data, self.bound, self.width = data, bound, width

    @property
    def rows(self) -> int:
        return self.data.shape[1]

    @property
    def ld(self) -> int:
        return self.data.shape[2]

    @property
    def hi(self) -> torch.Tensor:
        return self.data[0]

    @property
    def lo(self) -> torch.Tensor:
        return self.data[1]

    def float(self) -> torch.Tensor:
        """The fp32 matrix the pair represents (tests / debugging): (hi + lo) / scale."""
        import math

        b = float(self.bound)
        scale = 1.0 if not (b > 0) else 2.0 ** max(-60, min(60, 15 - math.frexp(b)[1]))
        return ((self.data[0].float() + self.data[1].float()) / scale)[:, : self.width]


def pair_empty(rows: int, width: int, device) -> Pair:
    ld = (width + 7) // 8 * 8
    return Pair(torch.empty(2, rows, ld, dtype=torch.float16, device=device), torch.empty(1, dtype=torch.float32, device=device), width)


def amax(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """max|x| of a 2-D fp32 tensor (unit inner stride) as a 1-element device tensor: the exact bound of an input."""
    xp, ld = _rows(x, "x")
    out = torch.empty(1, dtype=torch.float32, device=x.device) if out is None else out
    code = _lib.load().cusrl_b200_amax_f32(xp, ld, x.shape[0], x.shape[1], _ptr(out, torch.float32, "bound"), _stream())
    _lib.check(code, "amax")
    return out


def split_f16(x: torch.Tensor, bound: torch.Tensor | None = None, out: Pair | None = None) -> Pair:
    """The fp16 hi / lo pair of a 2-D fp32 tensor; `bound` defaults to its exact amax (one extra pass)."""
    xp, ld = _rows(x, "x")
    rows, width = x.shape
    out = pair_empty(rows, width, x.device) if out is None else out
    if bound is None:
        amax(x, out=out.bound)
    else:
        out.bound = bound
    code = _lib.load().cusrl_b200_split_f16(xp, ld, rows, width, _ptr(out.bound, torch.float32, "bound"), out.data[0].data_ptr(),
                                            out.data[1].data_ptr(), out.ld, _stream())
    _lib.check(code, "split_f16")
    return out


def weight_prep_f16(w: torch.Tensor, bias: torch.Tensor | None, out: dict | None = None) -> dict:
    """Pair, transposed pair and norm statistics of a weight matrix [N, K] (cusrl_b200_weight_prep_f16)."""
    N, K = w.shape
    ld, ldt = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    if out is None:
        f16 = torch.float16
        out = {"pair": torch.zeros(2, N, ld, dtype=f16, device=w.device), "pair_t": torch.zeros(2, K, ldt, dtype=f16, device=w.device),
               "stats": torch.zeros(4, dtype=torch.float32, device=w.device), "shape": (N, K)}
    code = _lib.load().cusrl_b200_weight_prep_f16(
        _ptr(w.detach(), torch.float32, "weight"), N, K, None if bias is None else _ptr(bias.detach(), torch.float32, "bias"),
        out["pair"][0].data_ptr(), out["pair"][1].data_ptr(), ld, out["pair_t"][0].data_ptr(), out["pair_t"][1].data_ptr(), ldt,
        out["stats"].data_ptr(), _stream())
    _lib.check(code, "weight_prep_f16", launches=2)
    return out


_weight_cache_f16: dict[int, tuple[tuple, dict, "weakref.ref"]] = {}


def _f16_stamp(w: torch.Tensor, bias: torch.Tensor | None) -> tuple:
    return (_weights_epoch, w._version, tuple(w.shape), None if bias is None else (bias.data_ptr(), bias._version))


def _f16_alloc(w: torch.Tensor) -> dict:
    N, K = w.shape
    ld, ldt = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    f16 = torch.float16
    return {"pair": torch.zeros(2, N, ld, dtype=f16, device=w.device), "pair_t": torch.zeros(2, K, ldt, dtype=f16, device=w.device),
            "stats": torch.zeros(4, dtype=torch.float32, device=w.device), "shape": (N, K)}


def prepared_weight_f16(w: torch.Tensor, bias: torch.Tensor | None) -> dict:
    """:func:`weight_prep_f16` of a parameter, rebuilt once per optimizer step (same protocol as :func:`prepared_weight`)."""
    key = w.data_ptr()
    stamp = _f16_stamp(w, bias)
    hit = _weight_cache_f16.get(key)
    alive = hit is not None and hit[2]() is not None
    if alive and hit[0] == stamp:
        return hit[1]
    reuse = alive and hit[0][2] == stamp[2]
    wp = weight_prep_f16(w, bias, out=hit[1] if reuse else None)
    _weight_cache_f16[key] = (stamp, wp, hit[2] if reuse else weakref.ref(w))
    if len(_weight_cache_f16) > 256:
        for k in [k for k, v in _weight_cache_f16.items() if v[2]() is None]:
            del _weight_cache_f16[k]
    return wp


def prepare_weights_f16(layers) -> None:
    """Refresh the f16x3 operand copies of several layers ``[(weight, bias), ...]`` at once: the stale ones are re-split by
    ONE multi-matrix call (cusrl_b200_weight_prep_f16_multi: three launches) instead of three launches per layer; the
    per-layer :func:`prepared_weight_f16` lookups that follow are cache hits."""
    stale = []
    for w, bias in layers:
        key = w.data_ptr()
        stamp = _f16_stamp(w, bias)
        hit = _weight_cache_f16.get(key)
        alive = hit is not None and hit[2]() is not None
        if alive and hit[0] == stamp:
            continue
        reuse = alive and hit[0][2] == stamp[2]
        wp = hit[1] if reuse else _f16_alloc(w)
        stale.append((w, bias, wp))
        _weight_cache_f16[key] = (stamp, wp, hit[2] if reuse else weakref.ref(w))
    for i in range(0, len(stale), 8):
        group = stale[i : i + 8]
        n = len(group)
        ptrs, ints = ctypes.c_void_p * n, ctypes.c_int64 * n
        W = ptrs(*[_ptr(w.detach(), torch.float32, "weight") for w, _, _ in group])
        B = ptrs(*[None if b is None else _ptr(b.detach(), torch.float32, "bias") for _, b, _ in group])
        Ns, Ks = ints(*[w.shape[0] for w, _, _ in group]), ints(*[w.shape[1] for w, _, _ in group])
        hi, lo = ptrs(*[wp["pair"][0].data_ptr() for *_, wp in group]), ptrs(*[wp["pair"][1].data_ptr() for *_, wp in group])
        hit_, lot = ptrs(*[wp["pair_t"][0].data_ptr() for *_, wp in group]), ptrs(*[wp["pair_t"][1].data_ptr() for *_, wp in group])
        ld, ldt = ints(*[wp["pair"].shape[2] for *_, wp in group]), ints(*[wp["pair_t"].shape[2] for *_, wp in group])
        stats = ptrs(*[wp["stats"].data_ptr() for *_, wp in group])
        code = _lib.load().cusrl_b200_weight_prep_f16_multi(n, W, Ns, Ks, B, hi, lo, ld, hit_, lot, ldt, stats, _stream())
        _lib.check(code, "weight_prep_f16_multi", launches=3)


def f16_linear_fwd(x: Pair, wp: dict, bias: torch.Tensor | None, act: int, out_pair: bool = True,
                   out: "Pair | torch.Tensor | None" = None):
    """Y = act(X W^T + b) from pairs; returns a Pair (its bound is the analytic one, published by the kernel) or fp32."""
    N, K = wp["shape"]
    M = x.rows
    if x.width != K:
        raise ValueError(f"f16_linear_fwd: input width {x.width} != weight K {K}")
    lib = _lib.load()
    w = wp["pair"]
    if out_pair:
        y = pair_empty(M, N, x.data.device) if out is None else out
        args = (None, 0, y.data[0].data_ptr(), y.data[1].data_ptr(), y.ld, y.bound.data_ptr())
    else:
        y = torch.empty(M, N, device=x.data.device) if out is None else out
        yp, ldy = _rows(y, "y")
        args = (yp, ldy, None, None, 0, None)
    code = lib.cusrl_b200_linear_fwd_f16x3(
        x.data[0].data_ptr(), x.data[1].data_ptr(), x.ld, x.bound.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), w.shape[2],
        wp["stats"].data_ptr(), None if bias is None else _ptr(bias.detach(), torch.float32, "bias"), *args, M, N, K, act, _stream())
    _lib.check(code, "linear_fwd_f16x3")
    return y


def f16_linear_dgrad(dy: Pair, wp: dict, x_act: Pair | None, act: int, out_pair: bool = True, db_below: torch.Tensor | None = None,
                     accumulate: bool = False):
    """dX = (dY W) * act'(x_act) from pairs (W as its transposed pair); `db_below` (+)= column sums of dX."""
    N, K = wp["shape"]
    M = dy.rows
    if dy.width != N:
        raise ValueError(f"f16_linear_dgrad: gradient width {dy.width} != weight N {N}")
    lib = _lib.load()
    wt = wp["pair_t"]
    dev = dy.data.device
    if out_pair:
        dx = pair_empty(M, K, dev)
        args = (None, 0, dx.data[0].data_ptr(), dx.data[1].data_ptr(), dx.ld, dx.bound.data_ptr())
    else:
        dx = torch.empty(M, (K + 3) // 4 * 4, device=dev)[:, :K]
        args = (dx.data_ptr(), dx.stride(0), None, None, 0, None)
    if db_below is not None:
        ws = _get_scratch(dev, "dgrad", lib.cusrl_b200_dgrad_workspace_bytes(K))
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    else:
        ws_ptr, ws_bytes = None, 0
    aux = (None, None, 0, None) if x_act is None else (x_act.data[0].data_ptr(), x_act.data[1].data_ptr(), x_act.ld, x_act.bound.data_ptr())
    code = lib.cusrl_b200_linear_dgrad_f16x3(
        dy.data[0].data_ptr(), dy.data[1].data_ptr(), dy.ld, dy.bound.data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), wt.shape[2],
        wp["stats"].data_ptr(), *aux, *args, M, N, K, act, _ptr(db_below, torch.float32, "db_below"), int(accumulate), ws_ptr,
        ws_bytes, _stream())
    _lib.check(code, "linear_dgrad_f16x3", launches=1 if db_below is None else 2)
    return dx


def gather_split_f16(src: torch.Tensor, index: torch.Tensor, bound: torch.Tensor, out: Pair | None = None) -> Pair:
    """pair[i] = split(src[index[i]]) for a 2-D fp32 source (unit inner stride), scale of `bound` (cusrl_b200_gather_split_f16)."""
    sp, lds = _rows(src, "src")
    n, width = index.numel(), src.shape[1]
    out = pair_empty(n, width, src.device) if out is None else out
    out.bound = bound
    code = _lib.load().cusrl_b200_gather_split_f16(sp, lds, _ptr(index, torch.int64, "index"), n, src.shape[0], width,
                                                   _ptr(bound, torch.float32, "bound"), out.data[0].data_ptr(),
                                                   out.data[1].data_ptr(), out.ld, _stream())
    _lib.check(code, "gather_split_f16")
    return out


# pairs that a producer (the sampler) prepared for a tensor a dense layer is about to consume: keyed by storage identity
_attached_pairs: dict[int, tuple["weakref.ref", tuple, Pair]] = {}


def attach_pair(tensor: torch.Tensor, pair: Pair) -> None:
    """Remember that `pair` is the f16x3 split of `tensor` as it is NOW (same memory, shape and version counter)."""
    if len(_attached_pairs) > 64:
        for k in [k for k, v in _attached_pairs.items() if v[0]() is None]:
            del _attached_pairs[k]
    _attached_pairs[tensor.data_ptr()] = (weakref.ref(tensor), (tuple(tensor.shape), tuple(tensor.stride()), tensor._version), pair)


_pair_only: dict[int, "weakref.ref"] = {}


def mark_pair_only(tensor: torch.Tensor) -> None:
    """`tensor` is a minibatch leaf whose fp32 rows were NOT gathered (only its attached fp16 pair is current)."""
    _pair_only[tensor.data_ptr()] = weakref.ref(tensor)


def unmark_pair_only(tensor: torch.Tensor) -> None:
    _pair_only.pop(tensor.data_ptr(), None)


def is_pair_only(tensor: torch.Tensor) -> bool:
    ref = _pair_only.get(tensor.data_ptr())
    return ref is not None and ref() is not None


def attached_pair(tensor: torch.Tensor) -> Pair | None:
    hit = _attached_pairs.get(tensor.data_ptr())
    if hit is None or hit[0]() is None:
        return None
    if hit[1] != (tuple(tensor.shape), tuple(tensor.stride()), tensor._version):
        return None
    return hit[2]


def head_bwd_pair(dy: torch.Tensor, h: torch.Tensor, w: torch.Tensor, act: int, dw: torch.Tensor, db: torch.Tensor | None,
                  accumulate: bool = False, db_trunk: torch.Tensor | None = None, accumulate_trunk: bool = False) -> Pair:
    """:func:`head_bwd` with dH returned as an f16x3 :class:`Pair` (cusrl_b200_head_bwd_f16pair): no fp32 dH, no split pass."""
    M, K = h.shape
    No = w.shape[0]
    hp, ldh = _rows(h, "h")
    out = pair_empty(M, K, h.device)
    dy_amax = amax(dy.reshape(M, No))
    lib = _lib.load()
    scratch = _get_scratch(h.device, "headbwd", lib.cusrl_b200_head_bwd_scratch_bytes(K, No))
    code = lib.cusrl_b200_head_bwd_f16pair(
        _ptr(dy, torch.float32, "dy"), dy_amax.data_ptr(), hp, ldh, _ptr(w.detach(), torch.float32, "weight"), act,
        out.data[0].data_ptr(), out.data[1].data_ptr(), out.ld, out.bound.data_ptr(), _ptr(dw, torch.float32, "dw"),
        _ptr(db, torch.float32, "db"), M, K, No, int(accumulate), _ptr(db_trunk, torch.float32, "db_trunk"),
        int(accumulate_trunk), scratch.data_ptr(), scratch.numel(), _stream())
    _lib.check(code, "head_bwd_f16pair", launches=2)
    return out


def colsum_(dz: torch.Tensor, db: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    """db (+)= column sums of a 2-D fp32 tensor (the bias gradient of a dense layer)."""
    M, N = dz.shape
    dzp, ld = _rows(dz, "dz")
    lib = _lib.load()
    ws = _get_scratch(dz.device, "colsum", lib.cusrl_b200_colsum_workspace_bytes(N))
    code = lib.cusrl_b200_colsum_f32(dzp, ld, M, N, _ptr(db, torch.float32, "db"), int(accumulate), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(code, "colsum", launches=2)
    return db


def f16_linear_wgrad(dz: Pair, x: Pair, dw: torch.Tensor, accumulate: bool = False) -> None:
    """dW (+)= dZ^T X from the two pairs (cusrl_b200_linear_wgrad_f16x3)."""
    M, N, K = dz.rows, dz.width, x.width
    if x.rows != M or not dw.is_contiguous() or dw.shape != (N, K):
        raise ValueError("f16_linear_wgrad: shape mismatch (dw must be a contiguous [N, K] tensor)")
    lib = _lib.load()
    ws = _get_scratch(dz.data.device, "wgrad16", lib.cusrl_b200_wgrad_f16x3_workspace_bytes(M, N, K))
    code = lib.cusrl_b200_linear_wgrad_f16x3(
        dz.data[0].data_ptr(), dz.data[1].data_ptr(), dz.ld, dz.bound.data_ptr(), x.data[0].data_ptr(), x.data[1].data_ptr(), x.ld,
        x.bound.data_ptr(), dw.data_ptr(), K, M, N, K, int(accumulate), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(code, "linear_wgrad_f16x3", launches=2)


# ---------------------------------------------------------------------------------------------- rollout (f1)
def copy_rows_padded(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[:, :width] = src, padding columns of dst's backing rows zeroed; `dst` is a [rows, width] view of a padded slot."""
    sp, lds = _rows(src, "src")
    dp, ldd = _rows(dst, "dst")
    code = _lib.load().cusrl_b200_copy_rows_padded_f32(sp, lds, dp, ldd, src.shape[0], src.shape[1], _stream())
    _lib.check(code, "copy_rows_padded")
    return dst


def rollout_store_step(next_obs, next_obs_slot, next_state, next_state_slot, reward, reward_slot, terminated, truncated,
                       terminated_slot, truncated_slot, done_slot) -> None:
    """ActorCritic.step's writes into the buffer slots of the current step in one launch (actor_critic.py:255-291,
    buffer.py:124-151): next_observation / next_state rows, reward, terminated, truncated, done = terminated | truncated."""
    def wide(src, dst, name):
        if src is None:
            return None, 0, None, 0, 0
        sp, lds = _rows(src, name)
        dp, ldd = _rows(dst, name + " slot")
        return sp, lds, dp, ldd, src.shape[1]

    N = reward.shape[0]
    code = _lib.load().cusrl_b200_rollout_store_step_f32(
        *wide(next_obs, next_obs_slot, "next_observation"), *wide(next_state, next_state_slot, "next_state"),
        _ptr(reward, torch.float32, "reward"), _ptr(reward_slot, torch.float32, "reward slot"), reward.shape[-1],
        _flag_ptr(terminated, "terminated"), _flag_ptr(truncated, "truncated"), _flag_ptr(terminated_slot, "terminated slot"),
        _flag_ptr(truncated_slot, "truncated slot"), _flag_ptr(done_slot, "done slot"), N, _stream())
    _lib.check(code, "rollout_store_step")


def sample_logp(mean: torch.Tensor, sigma: torch.Tensor, eps: torch.Tensor | None, std_out: torch.Tensor,
                action_out: torch.Tensor, logp_out: torch.Tensor, deterministic: bool = False) -> None:
    """Normal.rsample + log_prob into the buffer slots (distribution.py:195-213); see the header."""
    N, A = mean.shape
    f32 = torch.float32
    code = _lib.load().cusrl_b200_sample_logp_f32(
        _ptr(mean, f32, "mean"), _ptr(sigma, f32, "sigma"), _ptr(eps, f32, "eps"), N, A, int(deterministic),
        _ptr(std_out, f32, "std"), _ptr(action_out, f32, "action"), _ptr(logp_out, f32, "action_logp"), _stream())
    _lib.check(code, "sample_logp")


# ---------------------------------------------------------------------------------------------- observation normalisation (f2)
def column_stats(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[mean(C) | population var(C)] over the rows of a 2-D fp32 tensor (normalization.py:15-49, correction=0)."""
    xp, ld = _rows(x, "x")
    rows, C = x.shape
    lib = _lib.load()
    scratch = _get_scratch(x.device, "colstats", lib.cusrl_b200_column_stats_scratch_bytes(C))
    out = torch.empty(2 * C, dtype=torch.float32, device=x.device) if out is None else out
    code = lib.cusrl_b200_column_stats_f32(xp, ld, rows, C, _ptr(out, torch.float32, "mean_var"), scratch.data_ptr(), scratch.numel(), _stream())
    _lib.check(code, "column_stats", launches=2)
    return out


def rms_merge_(mean: torch.Tensor, var: torch.Tensor, std: torch.Tensor, batch_mean: torch.Tensor, batch_var: torch.Tensor,
               w_old: float, w_new: float, eps: float) -> None:
    """merge_mean_var_ + std refresh on the running statistics (normalization.py:78-93, rms.py:161-163)."""
    f32 = torch.float32
    code = _lib.load().cusrl_b200_rms_merge_f32(
        _ptr(mean, f32, "mean"), _ptr(var, f32, "var"), _ptr(std, f32, "std"), _ptr(batch_mean, f32, "batch_mean"),
        _ptr(batch_var, f32, "batch_var"), mean.numel(), float(w_old), float(w_new), float(eps), _stream())
    _lib.check(code, "rms_merge")


def rms_normalize(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, clamp: float | None, out: torch.Tensor | None = None,
                  zero_padding: bool = False) -> torch.Tensor:
    """(x - mean) / std, clamped (rms.py:202-214); x / out: 2-D fp32 with unit inner stride (out may alias x)."""
    xp, ldx = _rows(x, "x")
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device) if out is None else out
    op, ldo = _rows(out, "out")
    f32 = torch.float32
    code = _lib.load().cusrl_b200_rms_normalize_f32(xp, ldx, op, ldo, x.shape[0], x.shape[1], _ptr(mean, f32, "mean"), _ptr(std, f32, "std"),
                                                    -1.0 if clamp is None else float(clamp), int(zero_padding), _stream())
    _lib.check(code, "rms_normalize")
    return out


# ---------------------------------------------------------------------------------------------- K5
def rnd_reward_(target: torch.Tensor, pred: torch.Tensor, reward: torch.Tensor, reward_scale: float
                ) -> tuple[torch.Tensor, torch.Tensor]:
    """reward += scale * mean_d (target - pred)^2 (in place); returns (rnd_reward [.., 1], its mean) (rnd.py:68-75)."""
    D = target.shape[-1]
    M = target.numel() // D
    Dr = reward.shape[-1]
    lib = _lib.load()
    scratch = _get_scratch(target.device, "rnd", lib.cusrl_b200_rnd_scratch_bytes())
    rnd = torch.empty(*target.shape[:-1], 1, device=target.device)
    mean = torch.empty(1, device=target.device)
    code = lib.cusrl_b200_rnd_reward_f32(_ptr(target, torch.float32, "target"), _ptr(pred, torch.float32, "prediction"), M, D,
                                         float(reward_scale), _ptr(reward, torch.float32, "reward"), Dr, rnd.data_ptr(),
                                         mean.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream())
    _lib.check(code, "rnd_reward", launches=2)
    return rnd, mean


def mse_loss(pred: torch.Tensor, target: torch.Tensor, want_grad: bool = True) -> tuple[torch.Tensor, torch.Tensor | None]:
    """(mean squared error [1], d loss / d pred) in one pass (nn.MSELoss forward + backward)."""
    D = pred.shape[-1]
    M = pred.numel() // D
    lib = _lib.load()
    scratch = _get_scratch(pred.device, "rnd", lib.cusrl_b200_rnd_scratch_bytes())
    loss = torch.empty(1, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    code = lib.cusrl_b200_mse_f32(_ptr(pred, torch.float32, "prediction"), _ptr(target, torch.float32, "target"), M, D,
                                  loss.data_ptr(), None if grad is None else grad.data_ptr(), scratch.data_ptr(),
                                  scratch.numel(), _stream())
    _lib.check(code, "mse", launches=2)
    return loss, grad


def memory_reset_store(memories: list[torch.Tensor], done: torch.Tensor, dst_a: list | None = None, dst_b: list | None = None) -> None:
    """``memory[done] = 0`` in place for up to four dense [N, W] tensors, each result also copied to ``dst_a[k]`` / ``dst_b[k]``
    (entries may be None) in the same launch (cusrl_b200_memory_reset_store_f32)."""
    count = len(memories)
    N, W = memories[0].shape
    arr = ctypes.c_void_p * count

    def pointers(items):
        if items is None:
            return None
        out = []
        for m in items:
            if m is None:
                out.append(None)
                continue
            if m.shape != (N, W):
                raise ValueError("memory_reset_store: all tensors must share one [N, W] shape")
            out.append(_ptr(m, torch.float32, "memory"))
        return arr(*out)

    code = _lib.load().cusrl_b200_memory_reset_store_f32(pointers(memories), pointers(dst_a), pointers(dst_b), count,
                                                         _flag_ptr(done, "done"), N, W, _stream())
    _lib.check(code, "memory_reset_store")


# ---------------------------------------------------------------------------------------------- K7
def lstm_cell_fwd(xp: torch.Tensor, hp: torch.Tensor, c_in: torch.Tensor, done: torch.Tensor | None, gates: torch.Tensor,
                  c_out: torch.Tensor, h_out: torch.Tensor, c_next: torch.Tensor | None, h_next: torch.Tensor | None) -> None:
    """One LSTM step for all rows (cusrl_b200_lstm_cell_fwd_f32); xp may be a row-strided slice."""
    Nb, H = c_in.shape
    xpp, ldxp = _rows(xp, "xp")
    code = _lib.load().cusrl_b200_lstm_cell_fwd_f32(
        xpp, ldxp, _ptr(hp, torch.float32, "hp"), _ptr(c_in, torch.float32, "c_in"), None if done is None else _flag_ptr(done, "done"),
        _ptr(gates, torch.float32, "gates"), _ptr(c_out, torch.float32, "c_out"), _ptr(h_out, torch.float32, "h_out"),
        _ptr(c_next, torch.float32, "c_next"), _ptr(h_next, torch.float32, "h_next"), Nb, H, _stream())
    _lib.check(code, "lstm_cell_fwd")


def lstm_cell_bwd(dh_above: torch.Tensor, dh_rec: torch.Tensor | None, dc_rec: torch.Tensor | None, done: torch.Tensor | None,
                  gates: torch.Tensor, c: torch.Tensor, c_in: torch.Tensor, dgates: torch.Tensor, dc_prev: torch.Tensor) -> None:
    """BPTT of one LSTM step (cusrl_b200_lstm_cell_bwd_f32)."""
    Nb, H = c.shape
    dhp, lddh = _rows(dh_above, "dh_above")
    code = _lib.load().cusrl_b200_lstm_cell_bwd_f32(
        dhp, lddh, _ptr(dh_rec, torch.float32, "dh_rec"), _ptr(dc_rec, torch.float32, "dc_rec"),
        None if done is None else _flag_ptr(done, "done"), _ptr(gates, torch.float32, "gates"), _ptr(c, torch.float32, "c"),
        _ptr(c_in, torch.float32, "c_in"), _ptr(dgates, torch.float32, "dgates"), _ptr(dc_prev, torch.float32, "dc_prev"),
        Nb, H, _stream())
    _lib.check(code, "lstm_cell_bwd")


LSTM_SEQ = __import__("os").environ.get("CUSRL_B200_LSTM_SEQ", "1") not in ("", "0")   # 0: per-step recurrence everywhere


def lstm_seq_supported(H: int) -> bool:
    """The sequence-resident LSTM kernels (csrc/lstm_seq.cu) cover hidden sizes that are multiples of 64 up to 256."""
    return LSTM_SEQ and bool(_lib.load().cusrl_b200_lstm_seq_supported(H))


def lstm_seq_private(T: int, Nb: int, H: int, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(gates, cseq, cin) buffers of the sequence-resident LSTM kernels: opaque, in the kernels' private tiled layout
    [T][ceil(Nb/128)][H/4][(4 gates)][128 rows][4] (include/cusrl_b200.h) -- written by lstm_seq_fwd, read by lstm_seq_bwd."""
    nbp = (Nb + 127) // 128 * 128
    return (torch.empty(T, nbp, 4 * H, device=device), torch.empty(T, nbp, H, device=device), torch.empty(T, nbp, H, device=device))


def lstm_seq_fwd(xp: torch.Tensor, wp: dict, b_hh: torch.Tensor | None, h0: torch.Tensor | None, c0: torch.Tensor | None,
                 done: torch.Tensor | None, private: tuple[torch.Tensor, torch.Tensor, torch.Tensor] | None, out: torch.Tensor,
                 hin: torch.Tensor | None, c_last: torch.Tensor | None, h_last: torch.Tensor | None = None) -> None:
    """All T steps of one LSTM layer in one launch (cusrl_b200_lstm_seq_fwd_f32).  xp [T*Nb, 4H] = input projection incl.
    b_ih; `wp` = prepared_weight_f16(W_hh); writes out = h_t and hin = the hidden state entering each step ([T, Nb, H],
    row-major), h_last / c_last = h_{T-1} / c_{T-1}, and the `private` buffers of lstm_seq_private for the backward kernel
    (None: inference, nothing is saved); done [T, Nb] resets the carried state after the steps where it is set.  h0 / c0 and
    h_last / c_last may be column slices of a flat [N, layers * H] memory (unit inner stride, any row pitch)."""
    T, Nb, H = out.shape
    gates, cseq, cin = private if private is not None else (None, None, None)
    xpp, ldxp = _rows(xp, "xp")
    lib = _lib.load()
    need = lib.cusrl_b200_lstm_seq_workspace_bytes(T, Nb, H)
    ws = _get_scratch(out.device, "lstm_seq", need)
    w = wp["pair"]
    if tuple(wp["shape"]) != (4 * H, H):
        raise ValueError(f"lstm_seq_fwd: W_hh must be [{4 * H}, {H}], got {wp['shape']}")
    if done is not None and (done.numel() != T * Nb or not done.is_contiguous()):
        raise ValueError("lstm_seq_fwd: 'done' must be a contiguous [T, Nb] tensor")
    nbp = (Nb + 127) // 128 * 128
    if private is not None and (gates.numel() != T * nbp * 4 * H or cseq.numel() != T * nbp * H or cin.numel() != T * nbp * H):
        raise ValueError("lstm_seq_fwd: private buffers must come from lstm_seq_private(T, Nb, H)")

    def pitched(a, b, name):
        """(ptr a, ptr b, common row pitch) of two optional [Nb, H] views with unit inner stride."""
        ld = None
        ptrs = []
        for t in (a, b):
            if t is None:
                ptrs.append(None)
                continue
            tp, tl = _rows(t, name)
            if t.shape != (Nb, H) or (ld is not None and tl != ld):
                raise ValueError(f"lstm_seq_fwd: '{name}' tensors must be [{Nb}, {H}] views with one common row pitch")
            ld = tl
            ptrs.append(tp)
        return ptrs[0], ptrs[1], (ld or H)

    h0p, c0p, ld0 = pitched(h0, c0, "h0/c0")
    hlp, clp, ldl = pitched(h_last, c_last, "h_last/c_last")
    code = lib.cusrl_b200_lstm_seq_fwd_f32(
        xpp, ldxp, w[0].data_ptr(), w[1].data_ptr(), w.shape[2], wp["stats"].data_ptr(),
        None if b_hh is None else _ptr(b_hh.detach(), torch.float32, "b_hh"), h0p, c0p,
        None if done is None else _flag_ptr(done, "done"), ld0, _ptr(gates, torch.float32, "gates"),
        _ptr(cseq, torch.float32, "cseq"), _ptr(out, torch.float32, "out"), _ptr(hin, torch.float32, "hin"),
        _ptr(cin, torch.float32, "cin"), hlp, clp, ldl, T, Nb, H, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(code, "lstm_seq_fwd", launches=1)


def lstm_seq_bwd(dout: torch.Tensor, private: tuple[torch.Tensor, torch.Tensor, torch.Tensor], done: torch.Tensor | None,
                 wp: dict, dgates: torch.Tensor) -> None:
    """Backward through time of one LSTM layer in one launch (cusrl_b200_lstm_seq_bwd_f32): dgates [T, Nb, 4H] (row-major)
    from dout [T, Nb, H] and the forward's `private` buffers; `wp` = prepared_weight_f16(W_hh) (its transposed pair is the
    operand)."""
    T, Nb, H4 = dgates.shape
    H = H4 // 4
    gates, cseq, cin = private
    d2 = dout.reshape(T * Nb, H)
    dp, lddo = _rows(d2, "dout")
    lib = _lib.load()
    need = lib.cusrl_b200_lstm_seq_bwd_workspace_bytes(T, Nb, H)
    ws = _get_scratch(dgates.device, "lstm_seq_bwd", need)
    wt = wp["pair_t"]
    code = lib.cusrl_b200_lstm_seq_bwd_f32(
        dp, lddo, _ptr(gates, torch.float32, "gates"), _ptr(cseq, torch.float32, "cseq"), _ptr(cin, torch.float32, "cin"),
        None if done is None else _flag_ptr(done, "done"), wt[0].data_ptr(), wt[1].data_ptr(), wt.shape[2], wp["stats"].data_ptr(),
        _ptr(dgates, torch.float32, "dgates"), T, Nb, H, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(code, "lstm_seq_bwd", launches=2)


# ---------------------------------------------------------------------------------------------- symmetry (f3)
def mirror_rows(x: torch.Tensor, dest: torch.Tensor, mult: torch.Tensor, layout: str = "same",
                out: torch.Tensor | None = None) -> torch.Tensor:
    """Index-permute + sign-flip transforms of the rows of `x` ([..., C] fp32) in one launch (symmetry.py:58-61,84-95,334-339).

    `dest` int32 [V, C] / `mult` fp32 [V, C] are V transforms ``x[..., dest[v]] * mult[v]``.  Layouts of the result:
      "same"      V must be 1: the shape of `x`                                    (MirrorDef.__call__)
      "stacked"   [V, *x.shape]                                                    (_build_mirrored)
      "augmented" [*x.shape[:-1], V, C] -- with an identity first table row this is torch.cat([x.unsqueeze(-2), mirrored], -2),
                  the tensor SymmetricDataAugmentation stores                      (_build_augmented_tensor)
    `out` (augmented layout only) may be a narrow view of 16-byte-padded rows: the padding columns are zeroed."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32 or dest.dtype != torch.int32 or mult.dtype != torch.float32:
        raise TypeError("mirror_rows: x / mult must be float32 and dest int32")
    C = x.shape[-1]
    V = dest.shape[0]
    if dest.shape != (V, C) or mult.shape != (V, C):
        raise ValueError(f"mirror_rows: transform tables must be [V, {C}], got {tuple(dest.shape)} / {tuple(mult.shape)}")
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < C):   # e.g. an expanded std vector (row stride 0)
        x2 = x2.contiguous()
    rows = x2.shape[0]
    pad_to = 0
    if layout == "same":
        if V != 1:
            raise ValueError("mirror_rows: layout 'same' takes exactly one transform")
        res = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        target, stride_r, stride_v = res, C, 0
    elif layout == "stacked":
        res = torch.empty((V, *x.shape), dtype=x.dtype, device=x.device)
        target, stride_r, stride_v = res, C, rows * C
    elif layout == "augmented":
        if out is None:
            res = torch.empty((*x.shape[:-1], V, C), dtype=x.dtype, device=x.device)
            target, stride_r, stride_v = res, V * C, C
        else:
            if out.shape != (*x.shape[:-1], V, C) or out.stride(-1) != 1 or out.dtype != torch.float32:
                raise ValueError("mirror_rows: `out` must be a float32 [..., V, C] tensor with a dense last dim")
            ldo = out.stride(-2)
            o2 = out.reshape(-1, V, C) if out.dim() != 3 else out
            if o2.data_ptr() != out.data_ptr() or o2.stride(1) != ldo or (rows > 1 and o2.stride(0) != V * ldo):
                raise ValueError("mirror_rows: `out` rows must be uniformly pitched")
            res, target, stride_r, stride_v, pad_to = out, out, V * ldo, ldo, (ldo if ldo != C else 0)
    else:
        raise ValueError(f"mirror_rows: unknown layout '{layout}'")
    if rows == 0:
        return res
    code = _lib.load().cusrl_b200_mirror_rows_f32(
        x2.data_ptr(), x2.stride(0), rows, C, _ptr(dest, torch.int32, "dest"), _ptr(mult, torch.float32, "mult"), V,
        target.data_ptr(), stride_r, stride_v, pad_to, _stream())
    _lib.check(code, "mirror_rows")
    return res
