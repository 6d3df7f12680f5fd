"""GPU: tcgen05 dense-layer kernels (K6) and the SIMT heads against fp64 references.

Tolerances: 3xTF32 must be fp32-equivalent (max error within a small factor of cuBLAS fp32 SGEMM's own error on the
same inputs); single-pass TF32 must be within TF32's 2^-10 input rounding."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from cusrl_b200 import build

    build.build()
    from cusrl_b200 import ops as _ops

    return _ops


def _act(z, act):
    return torch.nn.functional.elu(z) if act == 1 else (torch.relu(z) if act == 2 else z)


def _padded(M, K, g):
    return torch.randn(M, (K + 3) // 4 * 4, generator=g).to(DEV)[:, :K]


@pytest.mark.parametrize("M,K,N,act", [(128, 32, 128, 0), (300, 235, 512, 1), (1000, 512, 256, 1), (1000, 256, 128, 1),
                                       (777, 64, 16, 0), (4096, 19, 64, 2), (1, 128, 128, 1)])
@pytest.mark.parametrize("precision", [3, 1])
def test_linear_fwd(ops, M, K, N, act, precision):
    g = torch.Generator().manual_seed(M + K + N)
    x = _padded(M, K, g)
    w = (torch.randn(N, K, generator=g) / K**0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.tc_linear_fwd(x, ops.weight_prep(w), b, N, act, precision)
    ref = _act(torch.nn.functional.linear(x.double(), w.double(), b.double()), act)
    err = (y.double() - ref).abs().max().item()
    f32 = (_act(torch.nn.functional.linear(x, w, b), act).double() - ref).abs().max().item()
    bound = max(8 * f32, 2e-6 * ref.abs().max().item()) if precision == 3 else 4e-3 * max(ref.abs().max().item(), 1.0)
    assert err <= bound, (err, f32)


@pytest.mark.parametrize("M,N,K,act", [(1000, 128, 256, 1), (777, 256, 512, 1), (500, 16, 128, 2), (333, 64, 20, 0)])
def test_linear_dgrad(ops, M, N, K, act):
    g = torch.Generator().manual_seed(M + N)
    dy = torch.randn(M, N, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / N**0.5).to(DEV)
    xa = torch.randn(M, (K + 3) // 4 * 4, generator=g).to(DEV)[:, :K]
    dx = ops.tc_linear_dgrad(dy, ops.weight_prep(w), xa if act else None, K, act, 3,
                             out=torch.empty(M, (K + 3) // 4 * 4, device=DEV)[:, :K])
    ref = dy.double() @ w.double()
    if act == 1:
        ref = ref * torch.where(xa > 0, torch.ones_like(xa), xa + 1).double()
    elif act == 2:
        ref = ref * (xa > 0).double()
    f32 = ((dy @ w).double() - dy.double() @ w.double()).abs().max().item()
    assert (dx.double() - ref).abs().max().item() <= max(8 * f32, 2e-6 * ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,act", [(1000, 128, 256, 1), (40000, 256, 512, 1), (333, 64, 20, 0), (70000, 128, 132, 2)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_linear_dgrad_emits_bias_gradient_of_layer_below(ops, M, N, K, act, accumulate):
    """db_below (+)= column sums of dX, produced by the data-gradient epilogue (ragged M / K edges, multi-item CTAs)."""
    g = torch.Generator().manual_seed(M + N + 1)
    dy = torch.randn(M, N, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / N**0.5).to(DEV)
    xa = torch.randn(M, (K + 3) // 4 * 4, generator=g).to(DEV)[:, :K]
    base = torch.randn(K, generator=g).to(DEV)
    db = base.clone()
    dx = ops.tc_linear_dgrad(dy, ops.weight_prep(w), xa if act else None, K, act, 3, db_below=db, accumulate=accumulate)
    ref = dx.double().sum(0) + (base.double() if accumulate else 0.0)   # column sums of what the kernel itself wrote
    assert (db.double() - ref).abs().max().item() <= 2e-6 * max(dx.double().abs().sum(0).max().item(), 1.0)
    # and the run-to-run result is bitwise reproducible (fixed-order reduction)
    db2 = base.clone()
    ops.tc_linear_dgrad(dy, ops.weight_prep(w), xa if act else None, K, act, 3, db_below=db2, accumulate=accumulate)
    assert torch.equal(db, db2)


@pytest.mark.parametrize("M,N,K", [(1024, 128, 128), (4096, 512, 235), (5000, 256, 512), (5000, 128, 256), (3000, 16, 128),
                                   (100, 64, 19)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_linear_wgrad(ops, M, N, K, accumulate):
    g = torch.Generator().manual_seed(M + N + K)
    dz = torch.randn(M, N, generator=g).to(DEV)
    x = _padded(M, K, g)
    base_w, base_b = torch.randn(N, K, generator=g).to(DEV), torch.randn(N, generator=g).to(DEV)
    dw, db = base_w.clone(), base_b.clone()
    ops.tc_linear_wgrad(dz, x, dw, db, 3, accumulate=accumulate)
    ref_w, ref_b = dz.double().t() @ x.double(), dz.double().sum(0)
    if accumulate:
        ref_w, ref_b = ref_w + base_w.double(), ref_b + base_b.double()
    f32 = ((dz.t() @ x).double() - dz.double().t() @ x.double()).abs().max().item()
    assert (dw.double() - ref_w).abs().max().item() <= max(8 * f32, 3e-6 * ref_w.abs().max().item())
    assert (db.double() - ref_b).abs().max().item() <= 1e-5 * max(ref_b.abs().max().item(), 1.0)


# One minibatch of BASELINE.json's headline configuration: 65536 envs x 24 steps / 4 = 393 216 rows.  The weight gradient
# reduces over ALL of them (the longest accumulation on the path), so this is where fp32-equivalence is hardest.
BENCH_M = 393216


@pytest.mark.parametrize("N,K", [(512, 235), (256, 512), (128, 256)])
def test_linear_wgrad_at_bench_minibatch(ops, N, K):
    """dW = dZ^T X at M = 393 216 against fp64, bound = 8x the error cuBLAS fp32 SGEMM makes on the same inputs (the same
    bar as the small shapes above; dtype "f32" in bench.py rests on it)."""
    g = torch.Generator(device=DEV).manual_seed(N + K)
    dz = torch.randn(BENCH_M, N, device=DEV, generator=g)
    x = torch.randn(BENCH_M, (K + 3) // 4 * 4, device=DEV, generator=g)[:, :K]
    dw, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    ops.tc_linear_wgrad(dz, x, dw, db, 3, accumulate=False)
    ref_w = dz.double().t() @ x.double()
    f32 = ((dz.t() @ x).double() - ref_w).abs().max().item()
    err = (dw.double() - ref_w).abs().max().item()
    scale = ref_w.abs().max().item()
    print(f"wgrad M={BENCH_M} N={N} K={K}: max abs err {err:.3e} (scale {scale:.1f}, rel {err / scale:.2e}), cuBLAS fp32 {f32:.3e}, ratio {err / f32:.2f}")
    assert err <= max(8 * f32, 3e-6 * scale), (err, f32)
    ref_b = dz.double().sum(0)
    assert (db.double() - ref_b).abs().max().item() <= 1e-5 * max(ref_b.abs().max().item(), 1.0)


def test_network_node_at_bench_minibatch(ops):
    """Forward + backward of the whole 235-512-256-128-12 actor at M = 393 216 against fp64 torch autograd."""
    from cusrl_b200.nn import functional as F

    g = torch.Generator(device=DEV).manual_seed(1)
    B, dims, No = BENCH_M, (235, 512, 256, 128), 12
    x = torch.randn(B, 236, device=DEV, generator=g)[:, :235]
    cpu = torch.Generator().manual_seed(2)
    ws = [(torch.randn(o, i, generator=cpu) / i**0.5).to(DEV).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(o, generator=cpu).mul(0.1).to(DEV).requires_grad_(True) for o in dims[1:]]
    hw = (torch.randn(No, dims[-1], generator=cpu) / dims[-1]**0.5).to(DEV).requires_grad_(True)
    hb = torch.zeros(No, device=DEV, requires_grad=True)
    out, latent = F.mlp_head_forward(x, ws, bs, "ELU", hw, hb)
    gout = torch.randn(B, No, device=DEV, generator=g) / B   # mean-reduced loss: per-sample gradients ~ 1/B
    out.backward(gout)
    got = [p.grad.clone() for p in ws + bs + [hw, hb]]
    del out
    params64 = [p.detach().double().requires_grad_(True) for p in ws + bs + [hw, hb]]
    w64, b64, hw64, hb64 = params64[:3], params64[3:6], params64[6], params64[7]
    h = x.double()
    for w, b in zip(w64, b64):
        h = torch.nn.functional.elu(torch.nn.functional.linear(h, w, b))
    assert torch.allclose(latent.double(), h, rtol=1e-5, atol=2e-5)
    ref_out = torch.nn.functional.linear(h, hw64, hb64)
    ref_out.backward(gout.double())
    # the same network in plain torch fp32 (cuBLAS SGEMM, what the reference runs): the yardstick for "fp32-equivalent"
    params32 = [p.detach().clone().requires_grad_(True) for p in ws + bs + [hw, hb]]
    h32 = x
    for w, b in zip(params32[:3], params32[3:6]):
        h32 = torch.nn.functional.elu(torch.nn.functional.linear(h32, w, b))
    torch.nn.functional.linear(h32, params32[6], params32[7]).backward(gout)
    names = [f"w{i}" for i in range(3)] + [f"b{i}" for i in range(3)] + ["head_w", "head_b"]
    for name, a, r64, r32 in zip(names, got, params64, params32):
        scale = r64.grad.abs().max().item()
        err = (a.double() - r64.grad).abs().max().item()
        f32 = (r32.grad.double() - r64.grad).abs().max().item()
        print(f"{name}: rel err {err / scale:.2e}, torch fp32 {f32 / scale:.2e}")
        assert err <= max(8 * f32, 2e-5 * scale), (name, err, f32, scale)


@pytest.mark.parametrize("M,K,No", [(1000, 128, 12), (5000, 128, 1), (257, 256, 8), (64, 128, 5)])
def test_heads(ops, M, K, No):
    g = torch.Generator().manual_seed(M + No)
    h = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(No, K, generator=g) / K**0.5).to(DEV)
    b = torch.randn(No, generator=g).to(DEV)
    y = ops.head_fwd(h, w, b)
    assert torch.allclose(y.double(), torch.nn.functional.linear(h.double(), w.double(), b.double()), rtol=1e-5, atol=1e-5)
    dy = torch.randn(M, No, generator=g).to(DEV)
    dw, db = torch.zeros(No, K, device=DEV), torch.zeros(No, device=DEV)
    dbt = torch.ones(K, device=DEV)
    dh = ops.head_bwd(dy, h, w, 1, dw, db, db_trunk=dbt, accumulate_trunk=True)
    dh_ref = (dy.double() @ w.double()) * torch.where(h > 0, torch.ones_like(h), h + 1).double()
    assert torch.allclose(dh.double(), dh_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(dbt.double(), 1.0 + dh_ref.sum(0), rtol=1e-5, atol=1e-4)
    assert torch.allclose(dw.double(), dy.double().t() @ h.double(), rtol=1e-5, atol=1e-4)
    assert torch.allclose(db.double(), dy.double().sum(0), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("M,K,No", [(1000, 128, 21), (300, 64, 19), (513, 128, 24)])
def test_general_head_beyond_simt_shapes(ops, M, K, No):
    """Heads the SIMT kernels do not instantiate (more than 16 outputs, odd widths: e.g. a 21-dimensional action) run on
    the dense-layer kernels with the output dimension zero-padded to a multiple of 4."""
    from cusrl_b200.nn import functional as F

    g = torch.Generator().manual_seed(M + No)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(No, K, generator=g) / K**0.5).to(DEV).requires_grad_(True)
    b = torch.randn(No, generator=g).to(DEV).requires_grad_(True)
    y = F.linear_head(x, w, b)
    assert y.shape == (M, No)
    gout = torch.randn(M, No, generator=g).to(DEV)
    y.backward(gout)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.linear(x64, w64, b64)
    ref.backward(gout.double())
    assert torch.allclose(y.double(), ref, rtol=1e-5, atol=1e-5)
    for a, r in ((x, x64), (w, w64), (b, b64)):
        assert torch.allclose(a.grad.double(), r.grad, rtol=1e-5, atol=1e-4 * max(r.grad.abs().max().item(), 1.0))


def test_network_node_matches_torch_autograd(ops):
    """The fused trunk+head autograd node against plain torch autograd of the same network (fp64 reference)."""
    from cusrl_b200.nn import functional as F

    g = torch.Generator().manual_seed(0)
    B, dims, No = 2048, (235, 512, 256, 128), 12
    x = _padded(B, dims[0], g)
    ws = [(torch.randn(o, i, generator=g) / i**0.5).to(DEV).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(o, generator=g).mul(0.1).to(DEV).requires_grad_(True) for o in dims[1:]]
    hw = (torch.randn(No, dims[-1], generator=g) / dims[-1]**0.5).to(DEV).requires_grad_(True)
    hb = torch.zeros(No, device=DEV, requires_grad=True)
    out, latent = F.mlp_head_forward(x, ws, bs, "ELU", hw, hb)
    gout = torch.randn(B, No, generator=g).to(DEV)
    out.backward(gout)
    got = [p.grad.clone() for p in ws + bs + [hw, hb]]
    params64 = [p.detach().double().requires_grad_(True) for p in ws + bs + [hw, hb]]
    w64, b64, hw64, hb64 = params64[:3], params64[3:6], params64[6], params64[7]
    h = x.double()
    for w, b in zip(w64, b64):
        h = torch.nn.functional.elu(torch.nn.functional.linear(h, w, b))
    ref_out = torch.nn.functional.linear(h, hw64, hb64)
    ref_out.backward(gout.double())
    assert torch.allclose(out.double(), ref_out, rtol=1e-5, atol=2e-5)
    assert torch.allclose(latent.double(), h, rtol=1e-5, atol=2e-5)
    for a, r in zip(got, params64):
        scale = r.grad.abs().max().item()
        assert (a.double() - r.grad).abs().max().item() <= 2e-5 * scale + 1e-6
