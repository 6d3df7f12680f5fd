"""Launch targets for `ncu --set full -k regex:f16x3`: one bench-minibatch instance of the f16x3 forward (235->512, pair
output), data gradient (256->512 with ELU' and bias-gradient column sums) and weight gradient (393216 x 256 x 512)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops  # noqa: E402

dev, M = "cuda", 393216
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(M, 236, device=dev, generator=g)[:, :235]
w1 = (torch.rand(512, 235, device=dev, generator=g) * 2 - 1) / 15.0
b1 = torch.randn(512, device=dev, generator=g) * 0.1
w2 = (torch.rand(256, 512, device=dev, generator=g) * 2 - 1) / 22.0
b2 = torch.randn(256, device=dev, generator=g) * 0.1
x0p = ops.split_f16(x0)
wp1, wp2 = ops.weight_prep_f16(w1, b1), ops.weight_prep_f16(w2, b2)
dz2 = ops.split_f16(torch.randn(M, 256, device=dev, generator=g) / M)
for _ in range(3):
    a1 = ops.f16_linear_fwd(x0p, wp1, b1, 1, True)            # 235 -> 512, pair out
    a2 = ops.f16_linear_fwd(a1, wp2, b2, 1, True)             # 512 -> 256
    db = torch.zeros(512, device=dev)
    dz1 = ops.f16_linear_dgrad(dz2, wp2, a1, 1, True, db_below=db, accumulate=True)   # 256 -> 512
    dw = torch.zeros(256, 512, device=dev)
    ops.f16_linear_wgrad(dz2, a1, dw, True)
torch.cuda.synchronize()
print("ok")
