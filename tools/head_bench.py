"""Head kernel timings (CUDA events over back-to-back launches on rotating buffers)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops
dev = "cuda"
M, K = 393216, 128
for No in (12, 1):
    hs = [torch.randn(M, K, device=dev) for _ in range(3)]
    w = torch.randn(No, K, device=dev) / K**0.5; b = torch.randn(No, device=dev)
    dy = torch.randn(M, No, device=dev); dw = torch.zeros(No, K, device=dev); db = torch.zeros(No, device=dev); dbt = torch.zeros(K, device=dev)
    for name, fn, nbytes in (("head_fwd", lambda h: ops.head_fwd(h, w, b), M * K * 4 + M * No * 4),
                             ("head_bwd", lambda h: ops.head_bwd(dy, h, w, 1, dw, db, db_trunk=dbt), 2 * M * K * 4 + M * No * 4)):
        for h in hs: fn(h)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(12): fn(hs[i % 3])
        e.record(); torch.cuda.synchronize()
        us = a.elapsed_time(e) / 12 * 1e3
        print(f"{name} No={No}: {us:.1f} us  {nbytes/us/1e3:.0f} GB/s", flush=True)
