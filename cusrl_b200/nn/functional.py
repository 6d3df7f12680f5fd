"""Dense-layer primitives of the actor / critic networks.

``mlp_forward`` is the single entry point used by :class:`cusrl_b200.nn.Mlp`; it owns forward AND backward
of the whole trunk (one autograd node) so that activation derivatives are fused into the gradient GEMM
epilogues and weight gradients are written straight into the flat gradient arena.
"""

from __future__ import annotations

import torch

from .. import ops

__all__ = ["ACTIVATIONS", "mlp_forward", "head_linear"]

ACTIVATIONS = {"Identity": 0, "ELU": 1, "ReLU": 2}


def _act_code(name: str) -> int:
    try:
        return ACTIVATIONS[name]
    except KeyError:
        raise ValueError(f"cusrl_b200 supports activations {sorted(ACTIVATIONS)}; got '{name}'") from None


class _MlpFunction(torch.autograd.Function):
    """y = act(... act(x W0^T + b0) ... Wk^T + bk); activations after every layer except optionally the last."""

    @staticmethod
    def forward(ctx, x, act_code, last_act, *params):
        weights, biases = params[0::2], params[1::2]
        acts = []
        h = x
        n = len(weights)
        for i, (w, b) in enumerate(zip(weights, biases)):
            code = act_code if (i < n - 1 or last_act) else 0
            h = ops.linear_fwd(h, w, b, code)
            acts.append(h)
        ctx.save_for_backward(x, *acts, *weights, *biases)
        ctx.meta = (act_code, last_act, n, [w.requires_grad for w in weights], x.requires_grad)
        return h

    @staticmethod
    def backward(ctx, grad_out):
        act_code, last_act, n, w_req, x_req = ctx.meta
        saved = ctx.saved_tensors
        x, acts, weights, biases = saved[0], saved[1 : 1 + n], saved[1 + n : 1 + 2 * n], saved[1 + 2 * n :]
        grads: list[torch.Tensor | None] = [None] * (2 * n)
        # dZ of the last layer: grad_out * act'(y_last) (identity when the trunk does not end with an activation)
        dz = ops.act_backward(grad_out.contiguous(), acts[-1], act_code if last_act else 0)
        for i in range(n - 1, -1, -1):
            inp = acts[i - 1] if i > 0 else x
            if w_req[i]:
                w_grad, b_grad = weights[i].grad, biases[i].grad
                if w_grad is not None and b_grad is not None:
                    # flat gradient arena (template/optimizer.py): accumulate in place, nothing for autograd to add
                    ops.linear_wgrad(dz, inp, out_w=w_grad, out_b=b_grad)
                else:
                    grads[2 * i], grads[2 * i + 1] = ops.linear_wgrad(dz, inp)
            if i > 0:
                # dX = dZ W, multiplied in the epilogue by act'(previous layer's output) -> dZ of layer i-1
                dz = ops.linear_dgrad(dz, weights[i], acts[i - 1], act_code)
            elif x_req:
                dz = ops.linear_dgrad(dz, weights[0], None, 0)
        return (dz if x_req else None, None, None, *grads)


def mlp_forward(x: torch.Tensor, weights, biases, activation: str, ends_with_activation: bool) -> torch.Tensor:
    """Trunk forward over the trailing feature dim; leading dims are flattened into the GEMM M dimension."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    params = [t for pair in zip(weights, biases) for t in pair]
    y = _MlpFunction.apply(x2, _act_code(activation), bool(ends_with_activation), *params)
    return y.reshape(*lead, y.shape[-1])


def head_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """Small-N fp32 output head (mean_head 128->12, value_head 128->1; reference LinearFp32, layer/linear.py)."""
    lead = x.shape[:-1]
    y = _MlpFunction.apply(x.reshape(-1, x.shape[-1]), 0, False, weight, bias)
    return y.reshape(*lead, y.shape[-1])
