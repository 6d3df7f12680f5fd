"""Micro-benchmarks of the HBM-bound kernels (CUDA events, L2 flushed between timed launches).

    python tools/kbench.py [--envs 65536] [--steps 24] [--reps 20] [--only gae]

Prints one JSON line per kernel: algorithmic bytes / launch, mean/min time, achieved GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json, else the 6650 GB/s fallback of B200_PROFILING.md).
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from cusrl_b200 import _lib, ops  # noqa: E402


def hbm_peak() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class Timer:
    """Times `fns` (a list of equivalent launches on DIFFERENT buffer sets whose total footprint exceeds
    the 126 MB L2, so every launch streams from HBM) as one CUDA graph: no host launch overhead inside
    the timed region.  Returns (mean, best) seconds per launch."""

    def __init__(self, reps: int):
        self.reps = reps

    def __call__(self, fns, warmup: int = 3) -> tuple[float, float]:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for f in fns:
                f()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for f in fns:
                f()
        for _ in range(warmup):
            graph.replay()
        torch.cuda.synchronize()
        times = []
        for _ in range(self.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) * 1e-3 / len(fns))
        return sum(times) / len(times), min(times)


def report(name, nbytes, mean_s, min_s, peak, which, **extra):
    print(json.dumps({
        "kernel": name, "bytes": nbytes, "mean_us": round(mean_s * 1e6, 2), "min_us": round(min_s * 1e6, 2),
        "gbs_mean": round(nbytes / mean_s / 1e9, 1), "gbs_best": round(nbytes / min_s / 1e9, 1),
        "frac_mean": round(nbytes / mean_s / 1e9 / peak, 3), "frac_best": round(nbytes / min_s / 1e9 / peak, 3),
        "peak_gbs": peak, "peak": which, **extra}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sets", type=int, default=8, help="rotating buffer sets (total footprint must exceed L2)")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    T, N, S = args.steps, args.envs, args.sets
    E = T * N
    B = E // 4
    peak, which = hbm_peak()
    timer = Timer(args.reps)
    dev = "cuda"
    lib = _lib.load()
    want = lambda k: not args.only or args.only in k  # noqa: E731

    sets = []
    for _ in range(S):
        d = {k: torch.randn(T, N, 1, device=dev) for k in ("reward", "value", "nv", "adv", "ret")}
        d["term"] = torch.rand(T, N, 1, device=dev) < 0.01
        d["trunc"] = torch.rand(T, N, 1, device=dev) < 0.001
        d["done"] = d["term"] | d["trunc"]
        d["boot"] = torch.randn(N, 1, device=dev)
        d["mv"] = torch.empty(2, device=dev)
        sets.append(d)

    if want("gae"):
        gae_calls = [lambda d=d: ops.gae(d["reward"], d["done"], d["value"], d["nv"], 0.99, 0.95,
                                         advantage=d["adv"], ret=d["ret"]) for d in sets]
        lib.cusrl_b200_gae_set_variant(0, 0, 2, 2)
        lib.cusrl_b200_gae_set_config(1, 64)
        for schedule, threads in ((1, 64), (1, 32), (0, 64), (0, 32), (1, 64), (0, 64)):
            lib.cusrl_b200_gae_set_schedule(schedule)
            lib.cusrl_b200_gae_set_config(1, threads)
            m, b = timer(gae_calls)
            report("gae", 21 * E, m, b, peak, which, variant="ldg", schedule=schedule, vec=1, threads=threads, T=T, N=N)
        lib.cusrl_b200_gae_set_schedule(0)
        for vec, threads in ((1, 32), (1, 64), (1, 96), (1, 128), (2, 32), (2, 64), (4, 32), (4, 64)):
            lib.cusrl_b200_gae_set_config(vec, threads)
            m, b = timer(gae_calls)
            report("gae", 21 * E, m, b, peak, which, variant="ldg", vec=vec, threads=threads, T=T, N=N)
            if (vec, threads) in ((1, 64), (1, 128)):
                m, b = timer([lambda d=d: ops.gae_fused(d["reward"], d["term"], d["trunc"], d["value"], d["boot"], 0.99, 0.95,
                                                        advantage=d["adv"], ret=d["ret"]) for d in sets])
                report("gae_fused(18B/elt)", 18 * E, m, b, peak, which, variant="ldg", vec=vec, threads=threads, T=T, N=N)
        lib.cusrl_b200_gae_set_config(1, 64)
        # TMA-staged variant: (warps per tile, stages, resident CTAs per SM the grid is sized for); warps 0 = automatic
        for warps, stages, ctas in ((0, 2, 1), (0, 2, 2), (0, 2, 3), (0, 2, 4), (7, 1, 2), (7, 2, 1), (8, 1, 2), (8, 2, 1),
                                    (5, 3, 1), (4, 2, 2), (4, 4, 1), (3, 2, 3), (2, 2, 4), (2, 4, 2), (1, 2, 8), (1, 4, 4), (1, 8, 2)):
            lib.cusrl_b200_gae_set_variant(1, warps, stages, ctas)
            m, b = timer(gae_calls)
            report("gae", 21 * E, m, b, peak, which, variant="tma", warps=warps, stages=stages, ctas_per_sm=ctas, T=T, N=N)
        lib.cusrl_b200_gae_set_variant(*ops.GAE_DEFAULT_VARIANT)
        lib.cusrl_b200_gae_set_schedule(ops.GAE_DEFAULT_SCHEDULE)
    if want("gae") or want("copy"):
        # context: a device-to-device copy moving the same 33 MB (16.5 MB read + 16.5 MB written) on rotating buffers
        half = (21 * E // 2 + 15) // 16 * 16
        srcs = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(S)]
        dsts = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(S)]
        m, b = timer([lambda a=a, c=c: c.copy_(a) for a, c in zip(srcs, dsts)])
        report("torch_copy_same_bytes", 2 * half, m, b, peak, which)
        del srcs, dsts
    if want("next_value"):
        m, b = timer([lambda d=d: ops.next_value(d["value"], d["term"], d["trunc"], d["boot"], out=d["nv"]) for d in sets])
        report("next_value", 10 * E, m, b, peak, which)
    if want("advnorm"):
        m, b = timer([lambda d=d: ops.advantage_stats(d["adv"], out=d["mv"]) for d in sets] * 4)
        report("advantage_stats(2 launches)", 4 * E, m, b, peak, which)
        m, b = timer([lambda d=d: ops.advantage_normalize_(d["adv"], d["mv"]) for d in sets] * 4)
        report("advantage_normalize", 8 * E, m, b, peak, which)
    del sets
    if want("loss"):
        ls = []
        for _ in range(4):
            d = {"mean": torch.randn(B, 12, device=dev), "action": torch.randn(B, 12, device=dev)}
            for k in ("lp", "a", "r", "vo", "v"):
                d[k] = torch.randn(B, 1, device=dev)
            ls.append(d)
        std = torch.ones(12, device=dev)
        outs = [ops.ppo_loss(d["mean"], std, d["action"], d["lp"], d["a"], d["r"], d["vo"], d["v"], 0.2, 1.0, 0.005, 0.5)
                for d in ls]
        del outs
        m, b = timer([lambda d=d: ops.ppo_loss(d["mean"], std, d["action"], d["lp"], d["a"], d["r"], d["vo"], d["v"],
                                               0.2, 1.0, 0.005, 0.5) for d in ls])
        report("ppo_loss(fwd+grads+per-sample, 2 launches)", (112 + 52 + 16) * B, m, b, peak, which, B=B)
        del ls
    if want("gather"):
        obs = torch.randn(E, 240, device=dev)
        action = torch.randn(E, 12, device=dev)
        sc = [torch.randn(E, 1, device=dev) for _ in range(4)]
        idx = torch.randperm(E, device=dev)[:B]
        d_obs, d_act = torch.empty(B, 240, device=dev), torch.empty(B, 12, device=dev)
        d_sc = [torch.empty(B, 1, device=dev) for _ in range(4)]
        fields = [(obs, d_obs), (action, d_act)] + list(zip(sc, d_sc))
        m, b = timer([lambda: ops.gather_rows(fields, idx)] * 2)
        report("gather(obs240+act12+4 scalars)", 2 * (960 + 48 + 16) * B + 8 * B, m, b, peak, which, B=B)
    if want("adam"):
        n = 571801
        p, g, m1, v1 = (torch.randn(n, device=dev) for _ in range(4))
        v1.abs_()
        sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        norm, coef = torch.zeros(1, device=dev), torch.zeros(1, device=dev)

        def step():
            sumsq.zero_()
            ops.grad_sumsq_(g, sumsq)
            ops.clip_coef(sumsq, 1.0, norm, coef)
            ops.adam_step_(p, g, m1, v1, 1, 1e-3, coef=coef)

        m, b = timer([step] * 8)
        report("clip+adam(memset + 3 launches, L2-resident)", 32 * n, m, b, peak, which, n=n)


if __name__ == "__main__":
    main()
