"""Both GAE scan variants at the BASELINE size (24 x 65536) for `ncu --set full -k regex:gae`: two warm-up launches and
one profiled launch each, on buffer sets that rotate so no launch finds its inputs in L2."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import _lib, ops
dev = "cuda"
T, N = 24, 65536
lib = _lib.load()
sets = []
for _ in range(6):
    d = {k: torch.randn(T, N, 1, device=dev) for k in ("reward", "value", "nv", "adv", "ret")}
    d["done"] = torch.rand(T, N, 1, device=dev) < 0.011
    sets.append(d)
for cfg in ((0, 0, 2, 2), (1, 0, 2, 2), (1, 0, 2, 1)):
    lib.cusrl_b200_gae_set_variant(*cfg)
    for d in sets[:3]:
        ops.gae(d["reward"], d["done"], d["value"], d["nv"], 0.99, 0.95, advantage=d["adv"], ret=d["ret"])
    torch.cuda.synchronize()
