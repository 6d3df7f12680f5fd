"""LSTM backbone for recurrent PPO (K7).

Mirrors the reference's ``Rnn`` / ``Rnn.Factory`` with ``module_type="LSTM"`` (cusrl/nn/module/rnn.py:62-97,133-340):
same factory arguments, flat ``{"hidden", "cell"}`` memory of shape ``[N, layers*hidden]`` (layout ``n (k c)``), same
parameter names (``rnn.weight_ih_l{k}``, ``rnn.weight_hh_l{k}``, ``rnn.bias_ih_l{k}``, ``rnn.bias_hh_l{k}`` -- an
``nn.LSTM`` is kept as the parameter container).  The arithmetic runs on the B200 kernels: per layer one tcgen05 GEMM
projects the inputs of ALL time steps, then per step one small recurrent GEMM and one fused cell kernel; training-time
episode boundaries are handled by resetting the state in-line where ``done`` (mathematically identical to the
reference's split / pad / scatter, pinned by cusrl_test/nn/module/test_rnn.py:145-164) -- no host syncs, no re-packing.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from .. import ops
from .modules import Module

__all__ = ["Rnn", "RnnFactory", "lstm_forward"]


class _LstmFunction(torch.autograd.Function):
    """out [T,Nb,H], (h_n, c_n) [L,Nb,H] = LSTM(x [T,Nb,I], h0/c0 [L,Nb,H], done [T,Nb] or None, weights...)."""

    @staticmethod
    def forward(ctx, x, h0, c0, done, num_layers, *weights):
        ctx.set_materialize_grads(False)   # h_n / c_n are non-differentiable outputs: no zero tensors for them in backward
        precision = ops.tf32_passes()
        T, Nb, _ = x.shape
        H = h0.shape[-1]
        dev = x.device
        layer_in = x
        saved_layers = []
        h_n, c_n = [], []
        for layer in range(num_layers):
            w_ih, w_hh, b_ih, b_hh = weights[4 * layer : 4 * layer + 4]
            inp2 = ops_rows(layer_in.reshape(T * Nb, layer_in.shape[-1]))
            xp = ops.tc_linear_fwd(inp2, ops.prepared_weight(w_ih), b_ih, 4 * H, 0, precision)            # [T*Nb, 4H]
            out = torch.empty(T, Nb, H, device=dev)
            hin = torch.empty(T, Nb, H, device=dev)
            if ops.lstm_seq_supported(H):
                # sequence-resident kernel: all T steps of the layer in one launch (csrc/lstm_seq.cu); gates / c_t / c_in stay
                # in the kernels' private layout for the backward twin
                private = ops.lstm_seq_private(T, Nb, H, dev)
                c_last = torch.empty(Nb, H, device=dev)
                ops.lstm_seq_fwd(xp, ops.prepared_weight_f16(w_hh, b_hh), b_hh, h0[layer], c0[layer], done, private, out, hin, c_last)
                h_n.append(out[T - 1])
                c_n.append(c_last)
                saved_layers.append((inp2, private, None, hin, None))
                layer_in = out
                continue
            gates = torch.empty(T, Nb, 4 * H, device=dev)
            cseq = torch.empty(T, Nb, H, device=dev)
            cin = torch.empty(T, Nb, H, device=dev)
            hin[0].copy_(h0[layer])
            cin[0].copy_(c0[layer])
            wp_hh = ops.prepared_weight(w_hh)
            h_last = torch.empty(Nb, H, device=dev)
            c_last = torch.empty(Nb, H, device=dev)
            for t in range(T):
                hp = ops.tc_linear_fwd(hin[t], wp_hh, b_hh, 4 * H, 0, precision)                           # [Nb, 4H]
                last = t == T - 1
                ops.lstm_cell_fwd(xp[t * Nb : (t + 1) * Nb], hp, cin[t], None if done is None else done[t], gates[t], cseq[t],
                                  out[t], None if last else cin[t + 1], None if last else hin[t + 1])
            h_n.append(out[T - 1])
            c_n.append(cseq[T - 1])
            saved_layers.append((inp2, gates, cseq, hin, cin))
            layer_in = out
        ctx.layers = saved_layers
        ctx.meta = (T, Nb, H, num_layers)
        ctx.done = done
        ctx.save_for_backward(*weights)
        h_n, c_n = torch.stack(h_n), torch.stack(c_n)
        ctx.mark_non_differentiable(h_n, c_n)
        return layer_in, h_n, c_n

    @staticmethod
    def backward(ctx, d_out, _dh, _dc):
        from .functional import _wgrad

        if d_out is None:
            return (None,) * (5 + len(ctx.saved_tensors))

        precision = ops.tf32_passes()
        T, Nb, H, num_layers = ctx.meta
        weights = ctx.saved_tensors
        done = ctx.done
        dev = d_out.device
        grads: list[Tensor | None] = [None] * len(weights)
        d_layer_out = d_out.contiguous()
        for layer in range(num_layers - 1, -1, -1):
            w_ih, w_hh, b_ih, b_hh = weights[4 * layer : 4 * layer + 4]
            inp2, gates, cseq, hin, cin = ctx.layers[layer]
            dgates = torch.empty(T, Nb, 4 * H, device=dev)
            seq = cseq is None   # the forward pass ran the sequence-resident kernel: `gates` holds its private buffers
            if seq:
                ops.lstm_seq_bwd(d_layer_out, gates, done, ops.prepared_weight_f16(w_hh, b_hh), dgates)   # gates = private buffers
            dc_buf = [torch.empty(Nb, H, device=dev), torch.empty(Nb, H, device=dev)]
            dh_rec = None
            wp_hh = None if seq else ops.prepared_weight(w_hh)
            for t in range(T - 1, -1, -1) if not seq else ():
                dc_rec = None if t == T - 1 else dc_buf[(t + 1) & 1]
                ops.lstm_cell_bwd(d_layer_out[t], dh_rec, dc_rec, None if (done is None or t == T - 1) else done[t],
                                  gates[t], cseq[t], cin[t], dgates[t], dc_buf[t & 1])
                if t > 0:
                    dh_rec = ops.tc_linear_dgrad(dgates[t], wp_hh, None, H, 0, precision)                   # dgates_t @ W_hh
            dg2 = dgates.reshape(T * Nb, 4 * H)
            # weight gradients over all steps at once; both biases receive the column sums of dgates: ONE reduction pass
            # over dgates (it reads T * Nb * 4H floats), handed to both
            db = ops.colsum_(dg2, torch.empty(4 * H, device=dev))
            for bias, slot in ((b_ih, 4 * layer + 2), (b_hh, 4 * layer + 3)):
                if bias.grad is not None:
                    bias.grad.add_(db)
                else:
                    grads[slot] = db.clone()
            _wgrad(dg2, inp2, w_ih, b_ih, grads, 4 * layer, 4 * layer + 2, bias_done=True)
            _wgrad(dg2, hin.reshape(T * Nb, H), w_hh, b_hh, grads, 4 * layer + 1, 4 * layer + 3, bias_done=True)
            if layer > 0:
                d_layer_out = ops.tc_linear_dgrad(dg2, ops.prepared_weight(w_ih), None, w_ih.shape[1], 0, precision).reshape(T, Nb, -1)
        ctx.layers = None
        return (None, None, None, None, None, *grads)


def ops_rows(x: Tensor) -> Tensor:
    from .functional import _rows_ok

    return _rows_ok(x)


def lstm_forward(x: Tensor, h0: Tensor, c0: Tensor, done: Tensor | None, lstm: nn.LSTM):
    weights = []
    for layer in range(lstm.num_layers):
        weights += [getattr(lstm, f"weight_ih_l{layer}"), getattr(lstm, f"weight_hh_l{layer}"),
                    getattr(lstm, f"bias_ih_l{layer}"), getattr(lstm, f"bias_hh_l{layer}")]
    return _LstmFunction.apply(x, h0, c0, done, lstm.num_layers, *weights)


def _expand_done(done: Tensor, T: int, n_rows: int, batch_shape) -> Tensor:
    """``[T, n_rows]`` episode-end flags from ``[T, N, 1]`` (or already flat) flags."""
    if done.numel() == T * n_rows:
        return done.reshape(T, n_rows)
    # batch dims the flags do not carry (SymmetricDataAugmentation: input [T, N, 1 + V, C], done [T, N, 1]): every variant of
    # an environment shares its episode boundaries
    d = done.squeeze(-1) if done.dim() >= 3 and done.shape[-1] == 1 else done
    while d.dim() < 1 + len(batch_shape):
        d = d.unsqueeze(-1)
    return d.expand(T, *batch_shape).reshape(T, n_rows)


@dataclass
class RnnFactory:
    module_type: str
    hidden_size: int
    num_layers: int = 1
    bias: bool = True
    dropout: float = 0.0

    def __call__(self, input_dim: int | None = None, output_dim: int | None = None) -> "Rnn":
        assert input_dim is not None
        if self.module_type.lower() != "lstm":
            raise ValueError(f"cusrl_b200 implements the LSTM of the recurrent PPO preset; got '{self.module_type}'")
        if not self.bias or self.dropout != 0.0 or output_dim:
            raise ValueError("cusrl_b200.Rnn supports bias=True, dropout=0 and no output projection (the preset defaults)")
        return Rnn(nn.LSTM(input_dim, self.hidden_size, self.num_layers))


class Rnn(Module):
    """LSTM backbone with the reference's memory convention (``Rnn.forward``, rnn.py:200-250)."""

    Factory = RnnFactory
    REQUIRE_CUDA = True   # tools/host_overhead_cpu.py (kernels stubbed out) clears it to drive the inference path on the CPU

    def __init__(self, rnn: nn.LSTM):
        if rnn.hidden_size % 4:
            raise ValueError("hidden_size must be a multiple of 4")
        super().__init__(rnn.input_size, rnn.hidden_size, is_recurrent=True)
        self.rnn = rnn

    def _initial(self, memory, lead_shape, n_rows: int, device):
        L, H = self.rnn.num_layers, self.rnn.hidden_size
        if memory is None:
            z = torch.zeros(L, n_rows, H, device=device)
            return z, z.clone()
        hidden, cell = memory["hidden"], memory["cell"]
        if lead_shape is not None and tuple(hidden.shape[:-1]) == tuple(lead_shape) and hidden.dim() >= 3:  # sequence-aligned: take step 0 (recurrent.py:202-212)
            hidden, cell = hidden[0], cell[0]
        to_lnh = lambda m: m.reshape(n_rows, L, H).transpose(0, 1).contiguous()  # "n (k c) -> k n c"  # noqa: E731
        return to_lnh(hidden), to_lnh(cell)

    def forward(self, input: Tensor, memory=None, *, done: Tensor | None = None, sequential: bool = True, **kwargs):
        L, H = self.rnn.num_layers, self.rnn.hidden_size
        if done is not None and not sequential:
            raise ValueError("'done' can be provided only when 'sequential' is True")
        if sequential and input.dim() >= 3:
            T, batch_shape = input.shape[0], tuple(input.shape[1:-1])
        else:
            T, batch_shape = 1, tuple(input.shape[:-1])
        n_rows = 1
        for d in batch_shape:
            n_rows *= d
        x = input.reshape(T, n_rows, input.shape[-1])
        if not torch.is_grad_enabled() and (x.is_cuda or not self.REQUIRE_CUDA) and ops.lstm_seq_supported(H):
            return self._forward_inference(input, x, memory, done, T, n_rows, batch_shape, sequential)
        # only a SEQUENCE input can carry a per-step memory; a single step with extra batch dims ([N, V, C]) must not be cut
        h0, c0 = self._initial(memory, input.shape[:-1] if (sequential and input.dim() >= 3) else None, n_rows, input.device)
        d = None if done is None else _expand_done(done, T, n_rows, batch_shape)
        out, h_n, c_n = lstm_forward(x, h0, c0, d, self.rnn)
        out = out.reshape(*input.shape[:-1], H)
        if done is not None:
            return out, None  # like the reference without packing: no output memory for segmented sequences (rnn.py:288-296)
        to_flat = lambda m: m.transpose(0, 1).reshape(*batch_shape, L * H)  # "k n c -> n (k c)"  # noqa: E731
        return out, {"hidden": to_flat(h_n), "cell": to_flat(c_n)}

    def _forward_inference(self, input: Tensor, x: Tensor, memory, done, T: int, n_rows: int, batch_shape, sequential: bool):
        """Rollout / statistics / bootstrap-value calls (no autograd): per layer one projection GEMM and ONE sequence-kernel
        launch that saves nothing for a backward pass, reads the layer's slice of the flat ``[N, layers * H]`` memory in
        place and writes the next memory's slice in place -- no state transposes, no per-step tensors."""
        L, H = self.rnn.num_layers, self.rnn.hidden_size
        dev = x.device
        hidden_in = cell_in = None
        if memory is not None:
            hidden_in, cell_in = memory["hidden"], memory["cell"]
            if sequential and input.dim() >= 3 and tuple(hidden_in.shape[:-1]) == tuple(input.shape[:-1]):
                hidden_in, cell_in = hidden_in[0], cell_in[0]   # sequence-aligned memory: step 0 (recurrent.py:202-212)
            hidden_in, cell_in = hidden_in.reshape(n_rows, L * H), cell_in.reshape(n_rows, L * H)
        d = None
        if done is not None:
            d = _expand_done(done, T, n_rows, batch_shape)
        new_hidden = torch.empty(n_rows, L * H, device=dev)
        new_cell = torch.empty(n_rows, L * H, device=dev)
        precision = ops.tf32_passes()
        layer_in = x
        for layer in range(L):
            w_ih, w_hh = getattr(self.rnn, f"weight_ih_l{layer}"), getattr(self.rnn, f"weight_hh_l{layer}")
            b_ih, b_hh = getattr(self.rnn, f"bias_ih_l{layer}"), getattr(self.rnn, f"bias_hh_l{layer}")
            inp2 = ops_rows(layer_in.reshape(T * n_rows, layer_in.shape[-1]))
            xp = ops.tc_linear_fwd(inp2, ops.prepared_weight(w_ih), b_ih, 4 * H, 0, precision)
            out = torch.empty(T, n_rows, H, device=dev)
            sl = slice(layer * H, (layer + 1) * H)
            ops.lstm_seq_fwd(xp, ops.prepared_weight_f16(w_hh, b_hh), b_hh, None if hidden_in is None else hidden_in[:, sl],
                             None if cell_in is None else cell_in[:, sl], d, None, out, None, new_cell[:, sl], new_hidden[:, sl])
            layer_in = out
        out = layer_in.reshape(*input.shape[:-1], H)
        if done is not None:
            return out, None
        return out, {"hidden": new_hidden.reshape(*batch_shape, L * H), "cell": new_cell.reshape(*batch_shape, L * H)}

    def step_memory(self, input: Tensor, memory=None, sequential: bool = True, **kwargs):
        return self(input, memory, sequential=sequential)[1]
