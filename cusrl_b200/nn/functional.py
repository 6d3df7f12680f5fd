"""Dense-layer autograd nodes of the actor / critic networks, built on the K6 kernels.

One autograd node owns forward AND backward of a whole network (trunk + optional small output head):

* forward  : per trunk layer one tcgen05 GEMM with fused bias + activation epilogue (``tc_linear_fwd``), then the
             fp32 SIMT head (``head_fwd``);
* backward : ``head_bwd`` produces the gradient w.r.t. the last trunk pre-activation directly (activation derivative
             fused), then per layer a split-K tcgen05 weight-gradient GEMM writing straight into the flat gradient
             arena and a data-gradient GEMM whose epilogue applies the next activation derivative.

Activations are saved post-activation only (ELU' and ReLU' are functions of the output), nothing is recomputed.
Reference arithmetic: nn.Linear + activation stacks of cusrl/nn/module/mlp.py:77-90, heads of
cusrl/nn/module/distribution.py:56,272 and cusrl/nn/module/critic.py:87-88, and their autograd.
"""

from __future__ import annotations

import torch

from .. import ops

__all__ = ["ACTIVATIONS", "f16_trunk_forward", "f16x3_supported", "linear_head", "mlp_forward", "mlp_head_forward"]

ACTIVATIONS = {"Identity": 0, "ELU": 1, "ReLU": 2}


def _act_code(name: str) -> int:
    try:
        return ACTIVATIONS[name]
    except KeyError:
        raise ValueError(f"cusrl_b200 supports activations {sorted(ACTIVATIONS)}; got '{name}'") from None


def _rows_ok(x: torch.Tensor) -> torch.Tensor:
    """2-D fp32 view with unit inner stride and a 16-byte-multiple row pitch (what TMA needs); copies only if not."""
    if x.stride(-1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0:
        return x
    width = x.shape[1]
    padded = torch.zeros(x.shape[0], (width + 3) // 4 * 4, dtype=x.dtype, device=x.device)
    padded[:, :width].copy_(x)
    return padded[:, :width]


def _in_arena(weight, bias) -> bool:
    """Both parameters of a layer accumulate straight into existing (flat-arena) gradient tensors."""
    return weight.grad is not None and weight.grad.is_contiguous() and (bias is None or bias.grad is not None)


def _wgrad(dz, inp, weight, bias, grads, slot, bias_slot=None, bias_done=False):
    """Weight / bias gradient of one layer: accumulate into the flat arena when the parameter has one, otherwise hand
    fresh tensors to autograd through ``grads[slot]`` / ``grads[bias_slot]`` (default: the adjacent slot).
    ``bias_done``: the bias gradient was already produced by the kernel that produced ``dz`` (its column sums)."""
    precision = ops.tf32_passes()
    bias_slot = slot + 1 if bias_slot is None else bias_slot
    if _in_arena(weight, bias):
        ops.tc_linear_wgrad(dz, inp, weight.grad, None if (bias is None or bias_done) else bias.grad, precision, accumulate=True)
    else:
        dw = torch.empty_like(weight)
        db = torch.empty_like(bias) if (bias is not None and not bias_done) else None
        ops.tc_linear_wgrad(dz, inp, dw, db, precision, accumulate=False)
        grads[slot] = dw
        if not bias_done:
            grads[bias_slot] = db


def _bias_target(weight, bias, grads, bias_slot):
    """(tensor, accumulate) the producer of a layer's dZ should add that layer's bias gradient to, or (None, False)."""
    if bias is None or not bias.requires_grad or not weight.requires_grad:
        return None, False
    if _in_arena(weight, bias):
        return bias.grad, True
    grads[bias_slot] = torch.empty_like(bias)
    return grads[bias_slot], False


# ---- precision 2: f16x3 (fp16 hi / lo split operands, csrc/f16x3_common.cuh) ---------------------------------------------
def f16x3_supported(weights, biases) -> bool:
    """The f16x3 kernels need output widths that are multiples of 4 and biases (for the analytic range bounds)."""
    return all(w.shape[0] % 4 == 0 for w in weights) and all(b is not None for b in biases)


def f16_trunk_forward(x: torch.Tensor, linears, act_code: int, last_act: bool):
    """Forward of a Linear(+activation) stack on the f16x3 kernels: the input is split once (exact amax + split), hidden
    activations stay fp16 hi / lo pairs written by the GEMM epilogues (never materialised in fp32), the last layer's output
    is fp32.  Returns (x pair, [hidden pairs..., fp32 output])."""
    ops.prepare_weights_f16(linears)   # every stale layer of the stack re-split by one multi-matrix call
    xp = ops.attached_pair(x)   # the sampler may already have produced the pair while gathering the minibatch
    if xp is None and ops.is_pair_only(x):
        raise RuntimeError("this minibatch leaf was gathered as an fp16 pair only (sampler.pair_only) and the pair is no longer "
                           "attached to it (the tensor was modified in place?)")
    if xp is None:
        xp = ops.split_f16(x if x.stride(-1) == 1 else x.contiguous())
        if x.stride(-1) == 1:
            ops.attach_pair(x, xp)   # actor and critic consume the same observations: split them once
    h, acts = xp, []
    n = len(linears)
    for i, (w, b) in enumerate(linears):
        code = act_code if (i < n - 1 or last_act) else 0
        h = ops.f16_linear_fwd(h, ops.prepared_weight_f16(w, b), b, code, out_pair=i < n - 1)
        acts.append(h)
    return xp, acts


def _f16_forward(ctx, x, act_code, last_act, has_head, n, params):
    weights, biases = params[0 : 2 * n : 2], params[1 : 2 * n : 2]
    xp, acts = f16_trunk_forward(x, list(zip(weights, biases)), act_code, last_act)
    latent = acts[-1]
    if has_head:
        out = ops.head_fwd(latent, params[2 * n], params[2 * n + 1])
    ctx.save_for_backward(latent, *params)
    ctx.pairs = (xp, acts[:-1])   # fp16 pairs are not autograd tensors: they ride on the context
    ctx.meta = (act_code, last_act, has_head, n)
    if has_head:
        ctx.mark_non_differentiable(latent)
        return out, latent
    return latent


def _f16_backward(ctx, grad_out):
    act_code, last_act, has_head, n = ctx.meta
    x_req = ctx.needs_input_grad[0]
    latent, params = ctx.saved_tensors[0], ctx.saved_tensors[1:]
    xp, hidden = ctx.pairs
    weights, biases = params[0 : 2 * n : 2], params[1 : 2 * n : 2]
    grads: list[torch.Tensor | None] = [None] * len(params)
    last_code = act_code if last_act else 0
    if has_head:
        head_w, head_b = params[2 * n], params[2 * n + 1]
        hw_grad, hb_grad = head_w.grad, (head_b.grad if head_b is not None else None)
        arena = hw_grad is not None and (head_b is None or hb_grad is not None)
        dw = hw_grad if arena else torch.empty_like(head_w)
        db = hb_grad if arena else (torch.empty_like(head_b) if head_b is not None else None)
        db_trunk, acc_trunk = _bias_target(weights[n - 1], biases[n - 1], grads, 2 * n - 1)
        # the head's backward emits the gradient entering the trunk directly as an fp16 pair (analytic bound)
        dzp = ops.head_bwd_pair(grad_out.contiguous(), latent, head_w, last_code, dw, db, accumulate=arena,
                                db_trunk=db_trunk, accumulate_trunk=acc_trunk)
        if not arena:
            grads[2 * n], grads[2 * n + 1] = dw, db
    else:
        dz = ops.act_grad_mul(grad_out, latent, last_code)
        db_last, acc_last = _bias_target(weights[n - 1], biases[n - 1], grads, 2 * n - 1)
        if db_last is not None:
            ops.colsum_(dz, db_last, accumulate=acc_last)
        dzp = ops.split_f16(dz)   # exact amax + split; below it stays in pairs
    for i in range(n - 1, -1, -1):
        inp = hidden[i - 1] if i > 0 else xp
        w = weights[i]
        if w.requires_grad:
            if _in_arena(w, biases[i]):
                ops.f16_linear_wgrad(dzp, inp, w.grad, accumulate=True)
            else:
                grads[2 * i] = torch.empty_like(w)
                ops.f16_linear_wgrad(dzp, inp, grads[2 * i], accumulate=False)
        if i > 0:
            db_below, acc_below = _bias_target(weights[i - 1], biases[i - 1], grads, 2 * i - 1)
            dzp = ops.f16_linear_dgrad(dzp, ops.prepared_weight_f16(w, biases[i]), hidden[i - 1], act_code, out_pair=True,
                                       db_below=db_below, accumulate=acc_below)
        elif x_req:
            dx = ops.f16_linear_dgrad(dzp, ops.prepared_weight_f16(w, biases[i]), None, 0, out_pair=False)
            return (dx, None, None, None, *grads)
    return (None, None, None, None, *grads)


class _MlpHeadFunction(torch.autograd.Function):
    """(latent, out) = trunk(x), head(trunk(x)); `has_head=False` returns the trunk output only."""

    @staticmethod
    def forward(ctx, x, act_code, last_act, has_head, *params):
        # the latent output is non-differentiable: without this autograd hands backward() a ZERO tensor of its shape for it
        # ([393 216, 128] fp32 = 201 MB filled per network per minibatch at the bench size, 1.1 ms per iteration)
        ctx.set_materialize_grads(False)
        ctx.n_params = len(params)
        precision = ops.tf32_passes()
        n = (len(params) - (2 if has_head else 0)) // 2
        weights, biases = params[0 : 2 * n : 2], params[1 : 2 * n : 2]
        ctx.f16 = ops.GEMM_PRECISION == 2 and x.shape[0] >= ops.F16X3_MIN_ROWS and f16x3_supported(weights, biases)
        if ctx.f16:
            return _f16_forward(ctx, x, act_code, last_act, has_head, n, params)
        if ops.is_pair_only(x):
            raise RuntimeError("this minibatch leaf was gathered as an fp16 pair only (sampler.pair_only) but the layer stack "
                               "takes the fp32 path: set sampler.pair_only = frozenset()")
        x = _rows_ok(x)
        acts = []
        h = x
        for i in range(n):
            code = act_code if (i < n - 1 or last_act) else 0
            h = ops.tc_linear_fwd(h, ops.prepared_weight(weights[i]), biases[i], weights[i].shape[0], code, precision)
            acts.append(h)
        if has_head:
            head_w, head_b = params[2 * n], params[2 * n + 1]
            out = ops.head_fwd(h, head_w, head_b)
        ctx.save_for_backward(x, *acts, *params)
        ctx.meta = (act_code, last_act, has_head, n)
        if has_head:
            ctx.mark_non_differentiable(h)
            return out, h
        return h

    @staticmethod
    def backward(ctx, grad_out, *unused):
        if grad_out is None:   # nothing flows into this network (set_materialize_grads(False))
            return (None,) * (4 + ctx.n_params)
        if ctx.f16:
            return _f16_backward(ctx, grad_out)
        act_code, last_act, has_head, n = ctx.meta
        x_req = ctx.needs_input_grad[0]
        precision = ops.tf32_passes()
        saved = ctx.saved_tensors
        x, acts, params = saved[0], saved[1 : 1 + n], saved[1 + n :]
        weights, biases = params[0 : 2 * n : 2], params[1 : 2 * n : 2]
        grads: list[torch.Tensor | None] = [None] * len(params)
        last_code = act_code if last_act else 0
        if has_head:
            head_w, head_b = params[2 * n], params[2 * n + 1]
            hw_grad, hb_grad = head_w.grad, (head_b.grad if head_b is not None else None)
            arena = hw_grad is not None and (head_b is None or hb_grad is not None)
            dw = hw_grad if arena else torch.empty_like(head_w)
            db = hb_grad if arena else (torch.empty_like(head_b) if head_b is not None else None)
            # dZ of the last trunk layer = (dOut W_head) * act'(latent), fused in the head kernel together with its
            # column sums (= that layer's bias gradient)
            db_trunk, acc_trunk = _bias_target(weights[n - 1], biases[n - 1], grads, 2 * n - 1)
            dz = ops.head_bwd(grad_out.contiguous(), acts[-1], head_w, last_code, dw, db, need_dh=True, accumulate=arena,
                              db_trunk=db_trunk, accumulate_trunk=acc_trunk)
            bias_done = db_trunk is not None
            if not arena:
                grads[2 * n], grads[2 * n + 1] = dw, db
        else:
            dz = ops.act_grad_mul(grad_out, acts[-1], last_code)
            bias_done = False
        for i in range(n - 1, -1, -1):
            inp = acts[i - 1] if i > 0 else x
            if weights[i].requires_grad:
                _wgrad(dz, inp, weights[i], biases[i], grads, 2 * i, bias_done=bias_done)
            if i > 0:
                # the data-gradient epilogue also emits the bias gradient of layer i-1 (column sums of its output)
                db_below, acc_below = _bias_target(weights[i - 1], biases[i - 1], grads, 2 * i - 1)
                dz = ops.tc_linear_dgrad(dz, ops.prepared_weight(weights[i]), acts[i - 1], weights[i].shape[1], act_code, precision,
                                         db_below=db_below, accumulate=acc_below)
                bias_done = db_below is not None
            elif x_req:
                dz = ops.tc_linear_dgrad(dz, ops.prepared_weight(weights[0]), None, weights[0].shape[1], 0, precision)
        return (dz if x_req else None, None, None, None, *grads)


def _flatten(x: torch.Tensor) -> tuple[torch.Tensor, tuple[int, ...]]:
    return x.reshape(-1, x.shape[-1]), tuple(x.shape[:-1])


def mlp_forward(x: torch.Tensor, weights, biases, activation: str, ends_with_activation: bool) -> torch.Tensor:
    """Trunk only (e.g. the RND networks): leading dims are flattened into the GEMM M dimension."""
    x2, lead = _flatten(x)
    params = [t for pair in zip(weights, biases) for t in pair]
    y = _MlpHeadFunction.apply(x2, _act_code(activation), bool(ends_with_activation), False, *params)
    return y.reshape(*lead, y.shape[-1])


def mlp_head_forward(x: torch.Tensor, weights, biases, activation: str, head_weight: torch.Tensor,
                     head_bias: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:
    """(head output, latent) of an activation-terminated trunk followed by a small fp32 linear head."""
    x2, lead = _flatten(x)
    params = [t for pair in zip(weights, biases) for t in pair] + [head_weight, head_bias]
    out, latent = _MlpHeadFunction.apply(x2, _act_code(activation), True, True, *params)
    return out.reshape(*lead, out.shape[-1]), latent.reshape(*lead, latent.shape[-1])


_SIMT_HEADS = {(no, 1) for no in range(1, 17)} | {(no, 2) for no in (1, 2, 4, 6, 8)}


def simt_head_supported(out_features: int, in_features: int) -> bool:
    """Shapes instantiated in csrc/head_kernels.cu (latent 128 with 1..16 outputs, latent 256 with 1/2/4/6/8)."""
    return in_features % 128 == 0 and (out_features, in_features // 128) in _SIMT_HEADS


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    out = torch.zeros(rows, *t.shape[1:], dtype=t.dtype, device=t.device)
    out[: t.shape[0]].copy_(t)
    return out


class _HeadFunction(torch.autograd.Function):
    """y = x W^T + b for an output head fed by a non-MLP backbone (e.g. the LSTM): the fp32 SIMT head when its shape is
    instantiated (csrc/head_kernels.cu), otherwise the tcgen05 dense-layer kernels (needs out_features % 4 == 0)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _rows_ok(x)
        No, K = weight.shape
        simt = simt_head_supported(No, K)
        ctx.padded = None
        if simt:
            y = ops.head_fwd(x, weight, bias)
        elif No % 4 == 0:
            y = ops.tc_linear_fwd(x, ops.prepared_weight(weight), bias, No, 0, ops.tf32_passes())
        else:
            # e.g. a 21-dimensional action head: the dense-layer kernels need 16-byte output rows, so the layer runs with
            # zero rows appended to W (and zero columns to dY in the backward); the public tensors keep their shapes
            Np = (No + 3) // 4 * 4
            ctx.padded = ops.weight_prep(_pad_rows(weight.detach(), Np))
            bias_p = None if bias is None else _pad_rows(bias.detach(), Np)
            y = ops.tc_linear_fwd(x, ctx.padded, bias_p, Np, 0, ops.tf32_passes())[:, :No]
        ctx.save_for_backward(x, weight, bias)
        ctx.simt = simt
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dy = dy.contiguous()
        grads: list[torch.Tensor | None] = [None, None]
        need_dx = ctx.needs_input_grad[0]
        if ctx.simt:
            w_grad, b_grad = weight.grad, (bias.grad if bias is not None else None)
            arena = w_grad is not None and (bias is None or b_grad is not None)
            dw = w_grad if arena else torch.empty_like(weight)
            db = b_grad if arena else (torch.empty_like(bias) if bias is not None else None)
            dx = ops.head_bwd(dy, x, weight, 0, dw, db, need_dh=need_dx, accumulate=arena)
            if not arena:
                grads = [dw, db]
        elif ctx.padded is None:
            _wgrad(dy, x, weight, bias, grads, 0)
            dx = ops.tc_linear_dgrad(dy, ops.prepared_weight(weight), None, weight.shape[1], 0, ops.tf32_passes()) if need_dx else None
        else:
            No, K = weight.shape
            Np = ctx.padded["hi"].shape[0]
            dy_p = torch.zeros(dy.shape[0], Np, device=dy.device)
            dy_p[:, :No].copy_(dy)
            dw_p = torch.empty(Np, K, device=dy.device)
            db_p = torch.empty(Np, device=dy.device) if bias is not None else None
            ops.tc_linear_wgrad(dy_p, x, dw_p, db_p, ops.tf32_passes(), accumulate=False)
            grads = [dw_p[:No], None if db_p is None else db_p[:No]]
            dx = ops.tc_linear_dgrad(dy_p, ctx.padded, None, K, 0, ops.tf32_passes()) if need_dx else None
        return dx, grads[0], grads[1]


def linear_head(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """Output head on arbitrary leading dims (mean_head / value_head behind a recurrent backbone)."""
    x2, lead = _flatten(x)
    y = _HeadFunction.apply(x2, weight, bias)
    return y.reshape(*lead, y.shape[-1])
