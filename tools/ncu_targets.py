"""One warm-up + one profiled launch of every hot kernel at the BASELINE minibatch size (for `ncu --set full`)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops
dev = "cuda"
M = 393216
g = torch.Generator(device=dev).manual_seed(0)
x512 = torch.randn(M, 512, device=dev, generator=g); x256 = torch.randn(M, 256, device=dev, generator=g)
w = torch.randn(256, 512, device=dev, generator=g) / 22.6; b = torch.randn(256, device=dev, generator=g)
wp = ops.weight_prep(w)
dw = torch.zeros(256, 512, device=dev); db = torch.zeros(256, device=dev); dbb = torch.zeros(512, device=dev)
h = torch.randn(M, 128, device=dev, generator=g); hw = torch.randn(12, 128, device=dev, generator=g) / 11.3
hb = torch.zeros(12, device=dev); dy = torch.randn(M, 12, device=dev, generator=g)
hdw = torch.zeros(12, 128, device=dev); hdb = torch.zeros(12, device=dev); dbt = torch.zeros(128, device=dev)
T, N = 24, 65536
r = torch.randn(T, N, 1, device=dev); v = torch.randn(T, N, 1, device=dev); nv = torch.randn(T, N, 1, device=dev)
d = torch.rand(T, N, 1, device=dev) < 0.01
for _ in range(2):
    y = ops.tc_linear_fwd(x512, wp, b, 256, 1, 3)                                  # gemm_tf32_kernel<256,3,0>
    dx = ops.tc_linear_dgrad(x256, wp, x512, 512, 1, 3, db_below=dbb)              # gemm_tf32_kernel<256,3,1>
    ops.tc_linear_wgrad(x256, x512, dw, None, 3)                                   # wgrad_tf32_kernel<256,3>
    ops.head_fwd(h, hw, hb)
    ops.head_bwd(dy, h, hw, 1, hdw, hdb, db_trunk=dbt)
    ops.gae(r, d, v, nv, 0.99, 0.95)
    torch.cuda.synchronize()
