"""Per-kernel times of the f16x3 dense-layer kernels next to the 3xTF32 ones at one bench minibatch (M = 393 216), CUDA
events over back-to-back launches on rotating buffers larger than L2.  python tools/f16x3_bench.py"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops  # noqa: E402

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 393216


def timeit(fns, reps=3):
    for f in fns:
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for f in fns:
            f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * len(fns))


g = torch.Generator(device=dev).manual_seed(0)
for name, K, N in (("L1 235->512", 235, 512), ("L2 512->256", 512, 256), ("L3 256->128", 256, 128)):
    sets = 4
    xs = [torch.randn(M, (K + 3) // 4 * 4, device=dev, generator=g)[:, :K] for _ in range(sets)]
    w = (torch.rand(N, K, device=dev, generator=g) * 2 - 1) / K**0.5
    b = torch.randn(N, device=dev, generator=g) * 0.1
    dzs = [torch.randn(M, N, device=dev, generator=g) / M for _ in range(sets)]
    wp16, wp32 = ops.weight_prep_f16(w, b), ops.weight_prep(w)
    xps = [ops.split_f16(x) for x in xs]
    dzps = [ops.split_f16(dz) for dz in dzs]
    outs = [ops.pair_empty(M, N, dev) for _ in range(sets)]
    outs32 = [torch.empty(M, N, device=dev) for _ in range(sets)]
    dw = torch.zeros(N, K, device=dev)
    res = {"layer": name, "M": M}
    from cusrl_b200 import _lib
    sweep = {}
    for bn in (256, 128):
        _lib.load().cusrl_b200_f16x3_set_tile(bn)
        row = [timeit([lambda i=i: ops.f16_linear_fwd(xps[i], wp16, b, 1, True, out=outs[i]) for i in range(sets)]),
               timeit([lambda i=i: ops.f16_linear_fwd(xps[i], wp16, b, 1, False, out=outs32[i]) for i in range(sets)])]
        if not name.startswith("L1"):
            dbs = torch.zeros(K, device=dev)
            row.append(timeit([lambda i=i: ops.f16_linear_dgrad(dzps[i], wp16, xps[i], 1, True, db_below=dbs, accumulate=True) for i in range(sets)]))
        sweep[bn] = [round(v, 1) for v in row]
    res["tile_sweep_fwdpair_fwdf32_dgrad_us"] = sweep
    _lib.load().cusrl_b200_f16x3_set_tile(0)
    res["fwd_pair_us"] = timeit([lambda i=i: ops.f16_linear_fwd(xps[i], wp16, b, 1, True, out=outs[i]) for i in range(sets)])
    res["fwd_f32out_us"] = timeit([lambda i=i: ops.f16_linear_fwd(xps[i], wp16, b, 1, False, out=outs32[i]) for i in range(sets)])
    res["fwd_3xtf32_us"] = timeit([lambda i=i: ops.tc_linear_fwd(xs[i], wp32, b, N, 1, 3, out=outs32[i]) for i in range(sets)])
    res["wgrad_f16x3_us"] = timeit([lambda i=i: ops.f16_linear_wgrad(dzps[i], xps[i], dw, True) for i in range(sets)])
    res["wgrad_3xtf32_us"] = timeit([lambda i=i: ops.tc_linear_wgrad(dzs[i], xs[i], dw, None, 3, True) for i in range(sets)])
    if not name.startswith("L1"):
        db = torch.zeros(K, device=dev)
        res["dgrad_pair_us"] = timeit([lambda i=i: ops.f16_linear_dgrad(dzps[i], wp16, xps[i], 1, True, db_below=db, accumulate=True) for i in range(sets)])
        res["dgrad_3xtf32_us"] = timeit([lambda i=i: ops.tc_linear_dgrad(dzs[i], wp32, xs[i], K, 1, 3, db_below=db, accumulate=True) for i in range(sets)])
    res["split_x_us"] = timeit([lambda i=i: ops.split_f16(xs[i], out=xps[i]) for i in range(sets)])
    res["split_dz_us"] = timeit([lambda i=i: ops.split_f16(dzs[i], out=dzps[i]) for i in range(sets)])
    flops = 2.0 * M * N * K
    res["fwd_pair_algorithmic_tflops"] = round(flops / res["fwd_pair_us"] / 1e6, 1)
    res["fwd_pair_hbm_gbs"] = round(M * (K + N) * 4 / res["fwd_pair_us"] / 1e3, 1)
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    del xs, dzs, xps, dzps, outs, outs32
    torch.cuda.empty_cache()
