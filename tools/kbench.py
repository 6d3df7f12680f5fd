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
    def __init__(self, reps: int, flush_mb: int = 384):
        self.reps = reps
        self.flush = torch.empty(flush_mb * 1024 * 1024 // 4, device="cuda")

    def __call__(self, fn, warmup: int = 3) -> tuple[float, float]:
        for _ in range(warmup):
            fn()
        times = []
        for _ in range(self.reps):
            self.flush.add_(1.0)  # evict L2 (126 MB) so the timed launch reads HBM
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) * 1e-3)
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
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    T, N = args.steps, args.envs
    E = T * N
    B = E // 4
    peak, which = hbm_peak()
    timer = Timer(args.reps)
    dev = "cuda"
    lib = _lib.load()
    want = lambda k: not args.only or args.only in k  # noqa: E731

    reward, value, nv = (torch.randn(T, N, 1, device=dev) for _ in range(3))
    term = torch.rand(T, N, 1, device=dev) < 0.01
    trunc = torch.rand(T, N, 1, device=dev) < 0.001
    done = term | trunc
    boot = torch.randn(N, 1, device=dev)
    adv, ret = torch.empty_like(value), torch.empty_like(value)

    if want("gae"):
        for vec, threads in ((1, 128), (1, 256), (2, 64), (2, 128), (2, 256), (4, 64), (4, 128)):
            lib.cusrl_b200_gae_set_config(vec, threads)
            m, b = timer(lambda: ops.gae(reward, done, value, nv, 0.99, 0.95, advantage=adv, ret=ret))
            report("gae", 21 * E, m, b, peak, which, vec=vec, threads=threads, T=T, N=N)
            m, b = timer(lambda: ops.gae_fused(reward, term, trunc, value, boot, 0.99, 0.95, advantage=adv, ret=ret))
            report("gae_fused(19B/elt)", 19 * E, m, b, peak, which, vec=vec, threads=threads, T=T, N=N)
        lib.cusrl_b200_gae_set_config(2, 128)
    if want("next_value"):
        m, b = timer(lambda: ops.next_value(value, term, trunc, boot, out=nv))
        report("next_value", 10 * E, m, b, peak, which)
    if want("advnorm"):
        mv = torch.empty(2, device=dev)
        m, b = timer(lambda: ops.advantage_stats(adv, out=mv))
        report("advantage_stats", 4 * E, m, b, peak, which)
        m, b = timer(lambda: ops.advantage_normalize_(adv, mv))
        report("advantage_normalize", 8 * E, m, b, peak, which)
    if want("loss"):
        mean, action = torch.randn(B, 12, device=dev), torch.randn(B, 12, device=dev)
        std = torch.ones(12, device=dev)
        lp, a, r, vo, v = (torch.randn(B, 1, device=dev) for _ in range(5))
        m, b = timer(lambda: ops.ppo_loss(mean, std, action, lp, a, r, vo, v, 0.2, 1.0, 0.005, 0.5))
        report("ppo_loss(fwd+grads+per-sample)", (112 + 52 + 16) * B, m, b, peak, which, B=B)
    if want("gather"):
        obs = torch.randn(E, 240, device=dev)
        action = torch.randn(E, 12, device=dev)
        sc = [torch.randn(E, 1, device=dev) for _ in range(4)]
        idx = torch.randperm(E, device=dev)[:B]
        d_obs, d_act = torch.empty(B, 240, device=dev), torch.empty(B, 12, device=dev)
        d_sc = [torch.empty(B, 1, device=dev) for _ in range(4)]
        fields = [(obs, d_obs), (action, d_act)] + list(zip(sc, d_sc))
        m, b = timer(lambda: ops.gather_rows(fields, idx))
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

        m, b = timer(step)
        report("clip+adam(4 launches)", 32 * n, m, b, peak, which, n=n)


if __name__ == "__main__":
    main()
