"""Flat parameter / gradient arenas and the fused clip+Adam optimizer (K9).

Replaces the reference's ``OptimizerFactory`` -> ``torch.optim.Adam`` (cusrl/template/optimizer.py:94-251,
cusrl/preset/optimizer.py:9-23) on the hot path.  All trainable parameters of the agent (actor, critic,
hook modules) are re-pointed at slices of ONE contiguous fp32 arena; gradients live in a second arena
of the same layout.  That makes

* the per-minibatch gradient allreduce a single in-place collective (no cat / copy-back), and
* grad-norm clipping + Adam two streaming kernels over 2.3 MB instead of ~10 foreach launches.

``param_groups`` keeps the keys the reference's hooks read (``params``, ``param_names``, ``lr``) so
``GradientClipping`` and the KL-adaptive LR schedule work unchanged.
"""

from __future__ import annotations

from collections.abc import Iterable
from typing import Any

import torch
from torch import nn

from .. import ops

__all__ = ["AdamFactory", "FlatAdam", "ParamArena"]


class ParamArena:
    """Re-homes parameters into one flat tensor (16-byte aligned slices) with a matching grad arena."""

    ALIGN = 4  # floats

    def __init__(self, named_parameters: Iterable[tuple[str, nn.Parameter]]):
        self.names: list[str] = []
        self.params: list[nn.Parameter] = []
        self.offsets: list[int] = []
        total = 0
        for name, p in named_parameters:
            if not p.requires_grad:
                continue
            if p.dtype != torch.float32:
                raise TypeError(f"parameter '{name}' must be float32")
            self.names.append(name)
            self.params.append(p)
            self.offsets.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if not self.params:
            raise ValueError("No trainable parameters matched the optimizer filter")
        device = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            self.flat[off : off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off : off + n].view_as(p)
            p.grad = self.flat_grad[off : off + n].view_as(p)
        self.numel = total

    def rebind_gradients(self) -> int:
        """Make every ``p.grad`` the arena view again.  ``module.zero_grad()`` (set_to_none=True by default) or a user
        hook's ``p.grad = None`` detaches a parameter from the arena: autograd then allocates a fresh gradient tensor, the
        weight-gradient kernels stop accumulating in place and the optimizer would step on a stale arena.  A detached
        gradient that holds values (autograd wrote it) is copied into its arena slice first, so nothing is lost.
        Returns the number of parameters that had to be re-bound."""
        fixed = 0
        base = self.flat_grad.data_ptr()
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            if g is not None and g.data_ptr() == base + 4 * off:
                continue
            view = self.flat_grad[off : off + p.numel()].view_as(p)
            if g is not None:
                view.copy_(g)
            p.grad = view
            fixed += 1
        return fixed

    def segment(self, name_prefix: str) -> list[tuple[int, int]]:
        """(offset, count) ranges of every parameter whose name equals or starts with `prefix.`."""
        out = []
        for name, p, off in zip(self.names, self.params, self.offsets):
            if name == name_prefix or name.startswith(name_prefix + "."):
                out.append((off, p.numel()))
        return out


class FlatAdam:
    """torch.optim.Adam semantics (amsgrad=False) on a :class:`ParamArena`, one kernel per step."""

    def __init__(self, named_parameters, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.arena = ParamArena(named_parameters)
        self.defaults = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay}
        self.param_groups: list[dict[str, Any]] = [
            {"params": list(self.arena.params), "param_names": list(self.arena.names), **self.defaults}
        ]
        dev = self.arena.flat.device
        self.exp_avg = torch.zeros_like(self.arena.flat)
        self.exp_avg_sq = torch.zeros_like(self.arena.flat)
        self.step_count = 0
        # clip state written by GradientClipping.pre_optim, consumed (then cleared) by step()
        self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.clip_coef = torch.ones(1, dtype=torch.float32, device=dev)
        self._clip_pending = False
        # device-resident copies of the step count and learning rate (see use_device_scalars)
        self.step_dev: torch.Tensor | None = None
        self.lr_dev: torch.Tensor | None = None
        self._lr_uploaded: float | None = None

    # ---- device-resident scalars: what a captured CUDA graph of the optimizer step needs ----------------------------
    def use_device_scalars(self) -> None:
        """From now on ``step()`` reads the step count and the learning rate from device memory
        (``cusrl_b200_adam_step_dev_f32``), so a CUDA graph containing the step stays valid across replays.  The step
        counter is advanced on the device by the step itself; the learning rate is uploaded by
        :meth:`sync_device_scalars`, which callers invoke OUTSIDE any capture."""
        if self.step_dev is None:
            dev = self.arena.flat.device
            self.step_dev = torch.full((1,), self.step_count, dtype=torch.int64, device=dev)
            self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            self._lr_uploaded = None
        self.sync_device_scalars()

    def sync_device_scalars(self) -> None:
        """Upload the learning rate if a schedule changed it (one fill, no host sync).  Never call while capturing."""
        if self.step_dev is None:
            return
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_uploaded:
            self.lr_dev.fill_(lr)
            self._lr_uploaded = lr

    # the two arenas, exposed for the collective (distributed.reduce_gradients) and for kernels
    @property
    def flat_param(self) -> torch.Tensor:
        return self.arena.flat

    @property
    def flat_grad(self) -> torch.Tensor:
        return self.arena.flat_grad

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Gradients are accumulated in place by the wgrad kernels, so zero the arena (one memset) -- after re-binding
        any ``p.grad`` that something detached from it since the last step."""
        self.arena.rebind_gradients()
        self.arena.flat_grad.zero_()

    def compute_grad_norm(self, max_norm: float) -> torch.Tensor:
        """L2 norm of the whole arena + clip coefficient (torch clip_grad_norm_ arithmetic), no host sync.
        The scaling itself is folded into the Adam kernel."""
        self.grad_sumsq.zero_()
        ops.grad_sumsq_(self.arena.flat_grad, self.grad_sumsq)
        ops.clip_coef(self.grad_sumsq, max_norm, self.grad_norm, self.clip_coef)
        self._clip_pending = True
        return self.grad_norm

    def step(self) -> None:
        g = self.param_groups[0]
        self.arena.rebind_gradients()  # gradients autograd produced outside the arena are folded in, never dropped
        self.step_count += 1
        coef = self.clip_coef if self._clip_pending else None
        if self.step_dev is not None:
            if float(g["lr"]) != self._lr_uploaded:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FlatAdam: the learning rate changed inside a CUDA-graph capture; call "
                                       "sync_device_scalars() before capturing")
                self.sync_device_scalars()
            self.step_dev.add_(1)
            ops.adam_step_dev_(self.arena.flat, self.arena.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_dev,
                               self.lr_dev, g["betas"], g["eps"], g["weight_decay"], coef=coef)
        else:
            ops.adam_step_(self.arena.flat, self.arena.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                           g["lr"], g["betas"], g["eps"], g["weight_decay"], coef=coef)
        self._clip_pending = False
        ops.invalidate_weight_cache()  # the Adam kernel rewrote the parameters through raw pointers

    # ---- checkpoints: the on-disk format is torch.optim.Adam's, i.e. the reference's (template/agent.py:283-330) ----
    _TORCH_ADAM_GROUP_DEFAULTS = {"amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                                  "differentiable": False, "fused": None, "decoupled_weight_decay": False}

    def state_dict(self) -> dict[str, Any]:
        """``{"state": {i: {step, exp_avg, exp_avg_sq}}, "param_groups": [...]}`` exactly as ``torch.optim.Adam`` (the
        reference's optimizer) writes it: per-parameter moment tensors in parameter order, ``step`` as an fp32 scalar
        tensor, the group carrying ``param_names`` -- so checkpoints move between the reference agent and this one."""
        state: dict[int, dict[str, torch.Tensor]] = {}
        if self.step_count > 0:
            for i, (p, off) in enumerate(zip(self.arena.params, self.arena.offsets)):
                n = p.numel()
                state[i] = {
                    "step": torch.tensor(float(self.step_count), dtype=torch.float32),
                    "exp_avg": self.exp_avg[off : off + n].view_as(p).clone(),
                    "exp_avg_sq": self.exp_avg_sq[off : off + n].view_as(p).clone(),
                }
        groups = []
        for g in self.param_groups:
            out = {k: v for k, v in g.items() if k != "params"}
            for k, v in self._TORCH_ADAM_GROUP_DEFAULTS.items():
                out.setdefault(k, v)
            out["params"] = list(range(len(g["params"])))
            groups.append(out)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state: dict[str, Any]) -> None:
        """Accepts a ``torch.optim.Adam`` state dict (from the reference or from :meth:`state_dict`); parameters are
        matched by ``param_names`` when the checkpoint has them, by position otherwise.  The flat form written by
        earlier versions of this class (``step`` / ``exp_avg`` / ``exp_avg_sq`` arenas) still loads."""
        if "state" not in state:  # legacy flat form
            self.step_count = int(state["step"])
            self.exp_avg.copy_(state["exp_avg"])
            self.exp_avg_sq.copy_(state["exp_avg_sq"])
        else:
            saved_groups = state.get("param_groups", [])
            saved_names = [n for g in saved_groups for n in g.get("param_names", [])]
            saved_ids = [i for g in saved_groups for i in g.get("params", [])]
            if len(saved_ids) != len(self.arena.params):
                raise ValueError(f"loaded state dict has {len(saved_ids)} parameters, the optimizer has {len(self.arena.params)}")
            if saved_names and len(saved_names) == len(saved_ids):
                if sorted(saved_names) != sorted(self.arena.names):
                    raise ValueError("loaded state dict names parameters this optimizer does not have: "
                                     f"{sorted(set(saved_names) ^ set(self.arena.names))[:4]}")
                by_name = dict(zip(saved_names, saved_ids))
                order = [by_name[n] for n in self.arena.names]
            else:
                order = saved_ids
            entries = state["state"]
            steps = set()
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for p, off, key in zip(self.arena.params, self.arena.offsets, order):
                entry = entries.get(key, entries.get(str(key)))
                if entry is None:
                    continue
                if tuple(entry["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {key} has shape {tuple(entry['exp_avg'].shape)}, parameter has {tuple(p.shape)}")
                n = p.numel()
                self.exp_avg[off : off + n].copy_(entry["exp_avg"].reshape(-1))
                self.exp_avg_sq[off : off + n].copy_(entry["exp_avg_sq"].reshape(-1))
                steps.add(int(float(entry["step"])))
            if len(steps) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): FlatAdam keeps one step count")
            self.step_count = steps.pop() if steps else 0
        ops.invalidate_weight_cache()
        if self.step_dev is not None:
            self.step_dev.fill_(self.step_count)
        for g, saved in zip(self.param_groups, state.get("param_groups", [])):
            g.update({k: v for k, v in saved.items() if k in ("lr", "betas", "eps", "weight_decay")})
            g["betas"] = tuple(g["betas"])


class AdamFactory:
    """Builds a :class:`FlatAdam` from named parameters (reference preset/optimizer.py:9-23)."""

    def __init__(self, defaults: dict[str, Any] | None = None):
        self.defaults = dict(defaults or {})
        unknown = set(self.defaults) - {"lr", "betas", "eps", "weight_decay"}
        if unknown:
            raise ValueError(f"unsupported Adam options: {sorted(unknown)}")

    def __call__(self, named_parameters) -> FlatAdam:
        return FlatAdam(named_parameters, **self.defaults)
