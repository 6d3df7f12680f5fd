import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops
p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M, K, N = 393216 // 4, 512, 256
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K**0.5; b = torch.randn(N, device="cuda")
wp = ops.weight_prep(w); y = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.tc_linear_fwd(x, wp, b, N, 1, p, out=y)
torch.cuda.synchronize()
