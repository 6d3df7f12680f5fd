"""Numerics + speed probe for an fp16 hi/lo split of the dense layers ("prototype an fp16 hi + scaled
fp16 lo split ... keep it only if it passes the same 8x cuBLAS-fp32 error bound at M = 393 216").

Emulates the scheme with cuBLAS fp16 GEMMs accumulating in fp32 (``torch.mm(..., out_dtype=torch.float32)``):

    x -> xs = x * s (s = 2^15 / bound(x), power of two)   hi = fp16(xs)   lo = fp16(xs - hi)      (lo unscaled, gradual underflow)
    A B^T ~= (Ah Bh^T + Ah Bl^T + Al Bh^T) / (sA sB)                                              (3 fp16 MMAs, fp32 accumulate)

and reports, per layer shape of the Anymal-C MLP at one bench minibatch, the max error against fp64 next to cuBLAS fp32
SGEMM's and (when the library is built) the shipped 3xTF32 kernels', plus kernel times.  `--loose` multiplies every bound by
2^8 to show that loose analytic range bounds cost nothing.  Run on a B200: python tools/fp16_split_probe.py
"""

from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def split(x: torch.Tensor, loose: float = 1.0):
    bound = float(x.abs().max()) * loose
    s = 2.0 ** (15 - math.ceil(math.log2(bound)))
    xs = x * s
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi, lo, s


def mm3(ah, al, bh, bl):
    """A [M,K] x B [N,K]^T with the three split products accumulated in fp32."""
    f32 = torch.float32
    return torch.mm(ah, bh.t(), out_dtype=f32) + torch.mm(ah, bl.t(), out_dtype=f32) + torch.mm(al, bh.t(), out_dtype=f32)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=393216)
    ap.add_argument("--loose", type=float, default=256.0)
    args = ap.parse_args()
    dev = "cuda"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(0)
    M = args.M
    try:
        from cusrl_b200 import ops
    except Exception:  # pragma: no cover
        ops = None
    rows = []
    for name, K, N in (("L1 235->512", 236, 512), ("L2 512->256", 512, 256), ("L3 256->128", 256, 128)):
        x = torch.randn(M, K, device=dev, generator=g)
        if name.startswith("L1"):
            x[:, 235:] = 0
        else:
            x = torch.nn.functional.elu(x)
        w = (torch.rand(N, K, device=dev, generator=g) * 2 - 1) / K**0.5
        dz = torch.randn(M, N, device=dev, generator=g) * (1.0 / M)      # gradients of a mean-reduced loss: ~1/M
        # ---- forward
        ref = x.double() @ w.double().t()
        f32 = ((x @ w.t()).double() - ref).abs().max().item()
        res = {"layer": name, "M": M, "fwd_scale": ref.abs().max().item(), "fwd_err_cublas_f32": f32}
        for tag, loose in (("exact_bound", 1.0), ("loose_bound", args.loose)):
            xh, xl, sx = split(x, loose)
            wh, wl, sw = split(w, loose)
            y = mm3(xh, xl, wh, wl) / (sx * sw)
            res[f"fwd_err_fp16x3_{tag}"] = (y.double() - ref).abs().max().item()
        res["fwd_us_cublas_f32"] = timeit(lambda: x @ w.t())
        res["fwd_us_cublas_fp16x3"] = timeit(lambda: mm3(xh, xl, wh, wl))
        res["fwd_us_cublas_fp16x1"] = timeit(lambda: torch.mm(xh, wh.t(), out_dtype=torch.float32))
        if ops is not None:
            wp = ops.weight_prep(w[:, : (235 if name.startswith("L1") else K)].contiguous())
            xin = x[:, :235] if name.startswith("L1") else x
            y3 = ops.tc_linear_fwd(xin, wp, None, N, 0, 3)
            res["fwd_err_3xtf32_kernel"] = (y3.double() - ref).abs().max().item()
            res["fwd_us_3xtf32_kernel"] = timeit(lambda: ops.tc_linear_fwd(xin, wp, None, N, 0, 3))
        del ref
        # ---- weight gradient dW = dZ^T X  (reduction over all M rows)
        refw = dz.double().t() @ x.double()
        res["wgrad_scale"] = refw.abs().max().item()
        res["wgrad_err_cublas_f32"] = ((dz.t() @ x).double() - refw).abs().max().item()
        for tag, loose in (("exact_bound", 1.0), ("loose_bound", args.loose)):
            zh, zl, sz = split(dz, loose)
            xh, xl, sx = split(x, loose)
            zt_h, zt_l, xt_h, xt_l = zh.t().contiguous(), zl.t().contiguous(), xh.t().contiguous(), xl.t().contiguous()
            dw = mm3(zt_h, zt_l, xt_h, xt_l) / (sz * sx)
            res[f"wgrad_err_fp16x3_{tag}"] = (dw.double() - refw).abs().max().item()
        res["wgrad_us_cublas_f32"] = timeit(lambda: dz.t() @ x)
        res["wgrad_us_cublas_fp16x3"] = timeit(lambda: mm3(zt_h, zt_l, xt_h, xt_l))
        if ops is not None:
            xin = x[:, :235] if name.startswith("L1") else x
            dwk = torch.zeros(N, xin.shape[1], device=dev)
            ops.tc_linear_wgrad(dz, xin, dwk, None, 3, accumulate=False)
            res["wgrad_err_3xtf32_kernel"] = (dwk.double() - refw[:, : xin.shape[1]]).abs().max().item()
            res["wgrad_us_3xtf32_kernel"] = timeit(lambda: ops.tc_linear_wgrad(dz, xin, dwk, None, 3, accumulate=False))
        del refw
        rows.append(res)
        print(json.dumps({k: (round(v, 10) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)


if __name__ == "__main__":
    main()
