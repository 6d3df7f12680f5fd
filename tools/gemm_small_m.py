"""Per-item time of the forward GEMM when everything is L2 resident (small M) vs streaming from HBM (large M)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops, _lib
dev = "cuda"
def t(M, K, N, p, reps=20):
    x = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    w = torch.randn(N, K, device=dev) / K**0.5
    b = torch.randn(N, device=dev)
    wp = ops.weight_prep(w)
    y = torch.empty(M, N, device=dev)
    for _ in range(3):
        ops.tc_linear_fwd(x, wp, b, N, 1, p, out=y)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.tc_linear_fwd(x, wp, b, N, 1, p, out=y)
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3
for M in (75776, 393216):
    for p in (3, 1):
        for (K, N) in ((235, 512), (512, 256), (256, 128)):
            us = t(M, K, N, p)
            items = ((M + 255) // 256) * ((N + 255) // 256)
            per_cluster = -(-items // 74)
            print(f"M={M:7d} K={K} N={N} p={p}: {us:7.1f} us  items/cluster={per_cluster:3d}  us/item={us/per_cluster:6.2f}", flush=True)
