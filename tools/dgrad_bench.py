"""Data-gradient GEMM timings at the minibatch size (with fused act' and bias-gradient column sums)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops
dev = "cuda"; M = 393216
for (N, K) in ((256, 512), (128, 256)):
    dys = [torch.randn(M, N, device=dev) for _ in range(2)]; xa = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / N**0.5; wp = ops.weight_prep(w); db = torch.zeros(K, device=dev)
    out = torch.empty(M, K, device=dev)
    for cs in (False, True):
        f = lambda dy: ops.tc_linear_dgrad(dy, wp, xa, K, 1, 3, out=out, db_below=db if cs else None)
        for dy in dys: f(dy)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10): f(dys[i % 2])
        e.record(); torch.cuda.synchronize()
        print(f"dgrad dY[{M},{N}] -> dX[{M},{K}] colsum={cs}: {a.elapsed_time(e)/10*1e3:.1f} us", flush=True)
