"""Does tcgen05.mma.kind::tf32 truncate or round (RN) its fp32 inputs to TF32?  Compare the single-pass kernel's output
with fp64 references built from truncated / RN-rounded inputs."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops, _lib

torch.manual_seed(0)
M, K, N = 512, 256, 256
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") / K**0.5

def trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

def rn(t):  # round-to-nearest-even to 10 explicit mantissa bits
    i = t.view(torch.int32)
    bias = ((i >> 13) & 1) + 0xFFF
    return ((i + bias) & ~0x1FFF).view(torch.float32)

# weight_prep masks W (truncation) in the 'hi' copy; build an UNMASKED operand copy to expose the hardware behaviour on both inputs
wp = ops.weight_prep(w)
wp_raw = {k: v.clone() for k, v in wp.items()}
wp_raw["hi"][:, :K] = w
y = ops.tc_linear_fwd(x, wp_raw, None, N, 0, 1)
torch.cuda.synchronize()
for name, f in (("truncate", trunc), ("round-nearest", rn)):
    ref = f(x).double() @ f(w).double().t()
    print(f"{name:14s}: max|y - ref| = {(y.double() - ref).abs().max().item():.3e}")
print(f"{'exact fp32':14s}: max|y - ref| = {(y.double() - x.double() @ w.double().t()).abs().max().item():.3e}")
