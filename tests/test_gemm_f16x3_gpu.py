"""GPU: the f16x3 dense-layer kernels (precision 2: tcgen05 kind::f16 on fp16 hi / lo split operands) against fp64.

Bar: the same as 3xTF32 in tests/test_gemm_gpu.py -- fp32-equivalent, i.e. max error within 8x the error cuBLAS fp32 SGEMM
makes on the same inputs ("keep it only if it passes the same 8x cuBLAS-fp32 error bound at M = 393 216")."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BENCH_M = 393216


@pytest.fixture(scope="module")
def ops():
    from cusrl_b200 import build

    build.build()
    from cusrl_b200 import ops as _ops

    return _ops


def _act(z, act):
    return torch.nn.functional.elu(z) if act == 1 else (torch.relu(z) if act == 2 else z)


def _act_grad(y, act):
    if act == 1:
        return torch.where(y > 0, torch.ones_like(y), y + 1)
    if act == 2:
        return (y > 0).to(y.dtype)
    return torch.ones_like(y)


@pytest.mark.parametrize("shape", [(1000, 37), (4096, 235), (3, 8), (70000, 512)])
@pytest.mark.parametrize("scale", [1.0, 1e-6, 3e4])
def test_split_roundtrip(ops, shape, scale):
    """x -> pair -> x: relative to the tensor's bound the error is <= 2^-22 of the element (22 significant bits) plus the
    2^-40-of-the-bound floor of lo's gradual underflow; padding columns are zero; the bound is the exact amax."""
    g = torch.Generator().manual_seed(shape[0])
    x = (torch.randn(*shape, generator=g) * scale).to(DEV)
    x[0, 0] = 0.0
    p = ops.split_f16(x)
    assert float(p.bound) == float(x.abs().max())
    assert p.ld % 8 == 0 and p.ld >= shape[1]
    if p.ld > shape[1]:
        assert float(p.data[:, :, shape[1]:].abs().max()) == 0.0
    back = p.float().double()
    err = (back - x.double()).abs()
    bound = float(p.bound)
    assert bool((err <= x.double().abs() * 2.0**-21 + bound * 2.0**-39).all()), err.max().item()


@pytest.mark.parametrize("M,K,N,act", [(128, 64, 128, 0), (300, 235, 512, 1), (1000, 512, 256, 1), (1000, 256, 128, 1),
                                       (777, 64, 16, 0), (4096, 19, 64, 2), (1, 128, 128, 1), (5000, 235, 128, 1)])
@pytest.mark.parametrize("out_pair", [False, True])
def test_linear_fwd(ops, M, K, N, act, out_pair):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K**0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    wp = ops.weight_prep_f16(w, b)
    stats = wp["stats"].cpu()
    assert stats[0].item() == pytest.approx(w.abs().max().item(), rel=1e-6)
    # rounded sums: the kernels widen the bounds they build from the L1 norms by 1.001
    assert stats[1].item() * 1.001 >= w.abs().sum(1).max().item() and stats[2].item() * 1.001 >= w.abs().sum(0).max().item()
    assert stats[1].item() <= w.abs().sum(1).max().item() * 1.001 and stats[2].item() <= w.abs().sum(0).max().item() * 1.001
    assert stats[3].item() == pytest.approx(b.abs().max().item(), rel=1e-6)
    y = ops.f16_linear_fwd(ops.split_f16(x), wp, b, act, out_pair=out_pair)
    ref = _act(torch.nn.functional.linear(x.double(), w.double(), b.double()), act)
    f32 = (_act(torch.nn.functional.linear(x, w, b), act).double() - ref).abs().max().item()
    if out_pair:
        assert float(y.bound) >= ref.abs().max().item()   # the analytic bound is an upper bound
        assert y.rows == M and y.width == N
        if y.ld > N:
            assert float(y.data[:, :, N:].abs().max()) == 0.0
        got = y.float().double()
    else:
        got = y.double()
    err = (got - ref).abs().max().item()
    assert err <= max(8 * f32, 2e-6 * ref.abs().max().item()), (err, f32)


@pytest.mark.parametrize("M,N,K,act", [(1000, 128, 256, 1), (777, 256, 512, 1), (500, 16, 128, 2), (333, 64, 24, 0),
                                       (40000, 256, 512, 1)])
@pytest.mark.parametrize("out_pair", [False, True])
def test_linear_dgrad_with_bias_gradient(ops, M, N, K, act, out_pair):
    g = torch.Generator().manual_seed(M + N)
    dy = (torch.randn(M, N, generator=g) / M).to(DEV)
    w = (torch.randn(N, K, generator=g) / N**0.5).to(DEV)
    xa = _act(torch.randn(M, K, generator=g), act).to(DEV)
    wp = ops.weight_prep_f16(w, None)
    base = torch.randn(K, generator=g).to(DEV) * 1e-3
    db = base.clone()
    dx = ops.f16_linear_dgrad(ops.split_f16(dy), wp, ops.split_f16(xa) if act else None, act, out_pair=out_pair, db_below=db,
                              accumulate=True)
    ref = (dy.double() @ w.double()) * _act_grad(xa.double(), act)
    f32 = ((dy @ w).double() * _act_grad(xa.double(), act) - ref).abs().max().item()
    got = dx.float().double() if out_pair else dx.double()
    if out_pair:
        assert float(dx.bound) >= ref.abs().max().item()
    assert (got - ref).abs().max().item() <= max(8 * f32, 2e-6 * ref.abs().max().item())
    ref_b = base.double() + ref.sum(0)
    assert (db.double() - ref_b).abs().max().item() <= 1e-5 * max(ref.abs().sum(0).max().item(), 1e-12) + 1e-9


@pytest.mark.parametrize("M,N,K", [(1024, 128, 128), (4096, 512, 235), (5000, 256, 512), (5000, 128, 256), (3000, 16, 128),
                                   (100, 64, 19), (BENCH_M, 512, 235), (BENCH_M, 256, 512), (BENCH_M, 128, 256)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_linear_wgrad(ops, M, N, K, accumulate):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    dz = torch.randn(M, N, device=DEV, generator=g) / M
    x = torch.randn(M, K, device=DEV, generator=g)
    base = torch.randn(N, K, device=DEV, generator=g) * 1e-3
    dw = base.clone()
    ops.f16_linear_wgrad(ops.split_f16(dz), ops.split_f16(x), dw, accumulate=accumulate)
    ref = dz.double().t() @ x.double() + (base.double() if accumulate else 0.0)
    f32 = ((dz.t() @ x).double() + (base.double() if accumulate else 0.0) - ref).abs().max().item()
    err = (dw.double() - ref).abs().max().item()
    print(f"wgrad f16x3 M={M} N={N} K={K}: err {err:.3e}, cuBLAS fp32 {f32:.3e}, ratio {err / max(f32, 1e-30):.2f}")
    assert err <= max(8 * f32, 3e-6 * ref.abs().max().item()), (err, f32)


@pytest.mark.parametrize("K,N", [(235, 512), (512, 256), (256, 128)])
def test_linear_fwd_at_bench_minibatch(ops, K, N):
    g = torch.Generator(device=DEV).manual_seed(K)
    x = torch.randn(BENCH_M, K, device=DEV, generator=g)
    w = (torch.rand(N, K, device=DEV, generator=g) * 2 - 1) / K**0.5
    b = torch.randn(N, device=DEV, generator=g) * 0.1
    y = ops.f16_linear_fwd(ops.split_f16(x), ops.weight_prep_f16(w, b), b, 1, out_pair=True)
    ref = _act(torch.nn.functional.linear(x.double(), w.double(), b.double()), 1)
    f32 = (_act(torch.nn.functional.linear(x, w, b), 1).double() - ref).abs().max().item()
    err = (y.float().double() - ref).abs().max().item()
    print(f"fwd f16x3 M={BENCH_M} {K}->{N}: err {err:.3e}, cuBLAS fp32 {f32:.3e}")
    assert err <= max(8 * f32, 2e-6 * ref.abs().max().item())


@pytest.mark.parametrize("B", [2048, BENCH_M])
@pytest.mark.parametrize("has_head", [True, False])
def test_network_node_f16x3_matches_fp64(ops, monkeypatch, B, has_head):
    """The fused trunk(+head) autograd node with GEMM_PRECISION = 2 against fp64 torch autograd, next to the same network
    in plain torch fp32 (cuBLAS SGEMM, what the reference runs): the yardstick for "fp32-equivalent"."""
    from cusrl_b200.nn import functional as F

    monkeypatch.setattr(ops, "GEMM_PRECISION", 2)
    g = torch.Generator(device=DEV).manual_seed(1)
    dims, No = (235, 512, 256, 128), 12
    x = torch.randn(B, 236, device=DEV, generator=g)[:, :235]
    cpu = torch.Generator().manual_seed(2)
    ws = [(torch.randn(o, i, generator=cpu) / i**0.5).to(DEV).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(o, generator=cpu).mul(0.1).to(DEV).requires_grad_(True) for o in dims[1:]]
    hw = (torch.randn(No, dims[-1], generator=cpu) / dims[-1]**0.5).to(DEV).requires_grad_(True)
    hb = torch.zeros(No, device=DEV, requires_grad=True)
    if has_head:
        out, latent = F.mlp_head_forward(x, ws, bs, "ELU", hw, hb)
        leaves = ws + bs + [hw, hb]
    else:
        out = latent = F.mlp_forward(x, ws, bs, "ELU", True)
        leaves = ws + bs
    gout = torch.randn(*out.shape, device=DEV, generator=g) / B
    out.backward(gout)
    got = [p.grad.clone() for p in leaves]
    p64 = [p.detach().double().requires_grad_(True) for p in leaves]
    p32 = [p.detach().clone().requires_grad_(True) for p in leaves]

    def net(xx, params):
        h = xx
        for w, b in zip(params[:3], params[3:6]):
            h = torch.nn.functional.elu(torch.nn.functional.linear(h, w, b))
        return (torch.nn.functional.linear(h, params[6], params[7]) if has_head else h), h

    ref_out, ref_latent = net(x.double(), p64)
    assert torch.allclose(latent.double(), ref_latent, rtol=1e-5, atol=2e-5)
    assert torch.allclose(out.double(), ref_out, rtol=1e-5, atol=2e-5)
    ref_out.backward(gout.double())
    net(x, p32)[0].backward(gout)
    for i, (a, r64, r32) in enumerate(zip(got, p64, p32)):
        scale = r64.grad.abs().max().item()
        err = (a.double() - r64.grad).abs().max().item()
        f32 = (r32.grad.double() - r64.grad).abs().max().item()
        print(f"B={B} head={has_head} param {i}: rel err {err / scale:.2e}, torch fp32 {f32 / scale:.2e}")
        assert err <= max(8 * f32, 2e-5 * scale), (i, err, f32, scale)


def test_head_bwd_pair_matches_fp32_head_bwd(ops):
    g = torch.Generator().manual_seed(5)
    M, K, No = 5000, 128, 12
    h = torch.nn.functional.elu(torch.randn(M, K, generator=g)).to(DEV)
    w = (torch.randn(No, K, generator=g) / K**0.5).to(DEV)
    dy = (torch.randn(M, No, generator=g) / M).to(DEV)
    dw0, db0, dbt0 = torch.zeros(No, K, device=DEV), torch.zeros(No, device=DEV), torch.zeros(K, device=DEV)
    dh = ops.head_bwd(dy, h, w, 1, dw0, db0, db_trunk=dbt0)
    dw1, db1, dbt1 = torch.zeros(No, K, device=DEV), torch.zeros(No, device=DEV), torch.zeros(K, device=DEV)
    pair = ops.head_bwd_pair(dy, h, w, 1, dw1, db1, db_trunk=dbt1)
    assert float(pair.bound) >= float(dh.abs().max())
    scale = float(dh.abs().max())
    assert (pair.float().double() - dh.double()).abs().max().item() <= 2.0**-20 * scale
    assert torch.equal(dw0, dw1) and torch.equal(db0, db1) and torch.equal(dbt0, dbt1)


def test_gather_split_matches_gather_then_split(ops):
    g = torch.Generator().manual_seed(6)
    E, K, B = 5000, 235, 1200
    back = torch.zeros(E, 236)
    back[:, :K] = torch.randn(E, K, generator=g)
    back = back.to(DEV)
    src = back[:, :K]
    idx = torch.randperm(E, generator=g)[:B].to(DEV)
    bound = ops.amax(src)
    pair = ops.gather_split_f16(src, idx, bound)
    ref = ops.split_f16(src[idx].contiguous(), bound=bound)
    assert torch.equal(pair.data, ref.data) and pair.width == K and pair.ld == 240
    # attach / lookup protocol used by the sampler
    dst = src[idx].contiguous()
    ops.attach_pair(dst, pair)
    assert ops.attached_pair(dst) is pair
    dst.add_(1.0)
    assert ops.attached_pair(dst) is None   # a torch-level in-place change invalidates the attachment


@pytest.mark.parametrize("rows,width,pitch", [(393216, 1, 1), (393216, 12, 12), (1001, 3, 3), (5, 1, 1), (4096, 235, 236), (77, 7, 7)])
def test_amax_exact_for_dense_narrow_and_pitched_inputs(rows, width, pitch):
    """The exact range of a tensor (scale of its fp16 pair): dense inputs of any width take the flat vector path (a [M, 1]
    value gradient used to be reduced one element per warp), pitched rows the row path."""
    from cusrl_b200 import build, ops

    build.build()
    g = torch.Generator().manual_seed(rows + width)
    back = torch.randn(rows, pitch, generator=g).cuda()
    x = back[:, :width]
    back[rows // 2, width - 1] = -123.5
    assert ops.amax(x).item() == x.abs().max().item() == 123.5


def test_multi_matrix_weight_prep_equals_per_matrix():
    """cusrl_b200_weight_prep_f16_multi (all stale layers of a network in three launches) against the per-matrix call:
    pairs, transposed pairs and norm statistics bit-identical; the cache serves the following per-layer lookups."""
    from cusrl_b200 import build, ops

    build.build()
    torch.manual_seed(3)
    shapes = [(512, 235), (256, 512), (128, 256), (12, 128), (1024, 256)]
    ws = [torch.nn.Parameter(torch.randn(n, k, device="cuda") / k**0.5) for n, k in shapes]
    bs = [torch.nn.Parameter(torch.randn(n, device="cuda") * 0.1) for n, _ in shapes]
    bs[3] = None
    single = [ops.weight_prep_f16(w, b) for w, b in zip(ws, bs)]
    ops.invalidate_weight_cache()
    ops.prepare_weights_f16(list(zip(ws, bs)))
    before = ops.launch_count()
    for w, b, ref in zip(ws, bs, single):
        got = ops.prepared_weight_f16(w, b)
        assert torch.equal(got["pair"], ref["pair"]) and torch.equal(got["pair_t"], ref["pair_t"])
        assert torch.equal(got["stats"], ref["stats"])
    assert ops.launch_count() == before   # all cache hits
    # a parameter update makes exactly that layer stale
    with torch.no_grad():
        ws[1].mul_(1.5)
    ops.prepare_weights_f16(list(zip(ws, bs)))
    assert ops.launch_count() == before + 3
    assert torch.equal(ops.prepared_weight_f16(ws[1], bs[1])["stats"], ops.weight_prep_f16(ws[1], bs[1])["stats"])
