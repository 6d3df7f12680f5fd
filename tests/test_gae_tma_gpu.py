"""GPU parity of the TMA-staged GAE scan (csrc/gae_tma.cu, selected with cusrl_b200_gae_set_variant):
bit-exact against the CPU oracle and the register-resident kernel for every tile width / stage count /
residency the launcher can pick, ragged column counts (clipped boxes), multi-wave pipelines (stage refill),
the lamda_value scan, advantage-only calls, and the layouts that must fall back to the LDG kernel."""

from __future__ import annotations

import pytest
import torch

from oracle import ppo_path as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def lib_ops():
    from cusrl_b200 import build

    build.build()
    from cusrl_b200 import _lib, ops

    return _lib.load(), ops


@pytest.fixture()
def tma(lib_ops):
    lib, ops = lib_ops

    def select(warps=0, stages=2, ctas_per_sm=2):
        assert lib.cusrl_b200_gae_set_variant(1, warps, stages, ctas_per_sm) == 0

    yield select, ops
    default = ops.GAE_DEFAULT_VARIANT
    lib.cusrl_b200_gae_set_variant(*default)


def _inputs(T, N, seed, p_done=0.05):
    g = torch.Generator().manual_seed(seed)
    reward, value, nv = (torch.randn(T, N, 1, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, generator=g) < p_done
    return reward, done, value, nv


@pytest.mark.parametrize("warps,stages,ctas", [(0, 2, 2), (1, 1, 1), (1, 3, 1), (2, 2, 4), (4, 8, 1), (7, 2, 1), (7, 1, 2), (8, 2, 1)])
@pytest.mark.parametrize("T,N", [(24, 4096), (24, 4112), (7, 48), (1, 16), (100, 1040)])
def test_tma_gae_vs_oracle_bit_exact(tma, warps, stages, ctas, T, N):
    select, ops = tma
    reward, done, value, nv = _inputs(T, N, T * 7919 + N)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)
    select(warps, stages, ctas)
    adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95)
    assert torch.equal(adv.cpu(), ref_adv)
    assert torch.equal(ret.cpu(), ref_ret)


@pytest.mark.parametrize("warps,stages,ctas", [(1, 2, 1), (2, 3, 1), (1, 1, 2)])
def test_tma_gae_many_waves_refills_stages(tma, warps, stages, ctas):
    """N / (32 warps) tiles >> resident CTAs: every CTA walks through several tiles and re-arms its stages."""
    select, ops = tma
    T, N = 12, 32 * 148 * 9 + 16
    reward, done, value, nv = _inputs(T, N, 11)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)
    select(warps, stages, ctas)
    adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95)
    assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)


def test_tma_gae_two_lambda_and_advantage_only(tma):
    select, ops = tma
    reward, done, value, nv = _inputs(24, 2048, 3)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, 0.7)
    select()
    adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95, 0.7)
    assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)
    adv_only, none = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95, compute_return=False)
    assert none is None
    assert torch.equal(adv_only.cpu(), O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)[0])


def test_tma_gae_does_not_touch_inputs_or_neighbours(tma):
    """The reference returns fresh tensors and leaves its inputs alone (gae.py:17); the tiles are rewritten in shared
    memory only.  Outputs live inside a larger allocation whose guard rows must stay untouched (clipped stores)."""
    select, ops = tma
    T, N = 24, 1008  # ragged against every tile width
    reward, done, value, nv = _inputs(T, N, 21)
    dr, dd, dv, dn = reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV)
    keep = [t.clone() for t in (dr, dd, dv, dn)]
    guard = torch.full((T + 2, N, 1), 123.0, device=DEV)
    guard_ret = torch.full((T + 2, N, 1), 321.0, device=DEV)
    select(7, 2, 1)
    ops.gae(dr, dd, dv, dn, 0.99, 0.95, advantage=guard[1:-1], ret=guard_ret[1:-1])
    for a, b in zip((dr, dd, dv, dn), keep):
        assert torch.equal(a, b)
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)
    assert torch.equal(guard[1:-1].cpu(), ref_adv) and torch.equal(guard_ret[1:-1].cpu(), ref_ret)
    assert bool((guard[0] == 123.0).all()) and bool((guard[-1] == 123.0).all())
    assert bool((guard_ret[0] == 321.0).all()) and bool((guard_ret[-1] == 321.0).all())


@pytest.mark.parametrize("T,N,Dv", [(24, 4099, 1), (24, 1024, 3), (5, 40, 1)])
def test_tma_variant_falls_back_on_untileable_layouts(tma, T, N, Dv):
    """N % 16 != 0 or Dv > 1 cannot be described by the tensor maps: the LDG kernel must run, same results."""
    select, ops = tma
    g = torch.Generator().manual_seed(N)
    reward, value, nv = (torch.randn(T, N, Dv, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, generator=g) < 0.05
    ref_adv, ref_ret = O.advantage_and_return_ref(reward, done, value, nv, 0.99, 0.95, None)
    select()
    adv, ret = ops.gae(reward.to(DEV), done.to(DEV), value.to(DEV), nv.to(DEV), 0.99, 0.95)
    assert torch.equal(adv.cpu(), ref_adv) and torch.equal(ret.cpu(), ref_ret)


def test_tma_gae_full_size_equals_register_kernel(tma, lib_ops):
    """65536 x 24 (BASELINE.json): both kernel variants produce identical bits, and the size-independent properties hold."""
    select, ops = tma
    lib, _ = lib_ops
    T, N = 24, 65536
    g = torch.Generator(device=DEV).manual_seed(0)
    reward, value, nv = (torch.randn(T, N, 1, device=DEV, generator=g) for _ in range(3))
    done = torch.rand(T, N, 1, device=DEV, generator=g) < 0.011
    lib.cusrl_b200_gae_set_variant(0, 0, 2, 2)
    adv0, ret0 = ops.gae(reward, done, value, nv, 0.99, 0.95)
    for cfg in [(0, 2, 2), (7, 2, 1), (8, 1, 2), (4, 4, 1)]:
        select(*cfg)
        adv, ret = ops.gae(reward, done, value, nv, 0.99, 0.95)
        assert torch.equal(adv, adv0) and torch.equal(ret, ret0), cfg
    assert torch.equal(ret0, value + adv0)
    assert torch.equal(adv0[-1], reward[-1] + nv[-1] * 0.99 - value[-1])


def test_set_variant_rejects_bad_arguments(lib_ops):
    lib, _ = lib_ops
    for bad in [(2, 0, 2, 2), (1, 9, 2, 2), (1, 0, 0, 2), (1, 0, 9, 2), (1, 0, 2, 0), (1, -1, 2, 2)]:
        assert lib.cusrl_b200_gae_set_variant(*bad) != 0
