"""CUDA-event timing of the sequence-resident LSTM kernels (csrc/lstm_seq.cu) at the recurrent-PPO shapes.

    python tools/lstm_bench.py [--debug BITS]
"""
import argparse, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only-config3", action="store_true", help="only the config-3 minibatch shape (ncu captures)")
args = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda", 0)


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e3


for debug in sorted({0, args.debug}):
    lib.cusrl_b200_lstm_seq_set_debug(debug)
    for T, Nb, H in (((24, 1024, 256),) if args.only_config3 else ((24, 1024, 256), (1, 4096, 256), (24, 4096, 256), (24, 1024, 128))):
        torch.manual_seed(0)
        w_hh = torch.randn(4 * H, H, device=dev) / H**0.5
        b_hh = torch.randn(4 * H, device=dev) * 0.1
        wp = ops.weight_prep_f16(w_hh, b_hh)
        xp = torch.randn(T * Nb, 4 * H, device=dev)
        h0 = torch.randn(Nb, H, device=dev).tanh()
        c0 = torch.randn(Nb, H, device=dev)
        done = torch.rand(T, Nb, device=dev) < 0.02
        private = ops.lstm_seq_private(T, Nb, H, dev)
        out, hin = (torch.empty(T, Nb, H, device=dev) for _ in range(2))
        c_last = torch.empty(Nb, H, device=dev)
        fwd = lambda: ops.lstm_seq_fwd(xp, wp, b_hh, h0, c0, done, private, out, hin, c_last)
        us_f = timeit(fwd, args.reps)
        line = {"debug": debug, "T": T, "Nb": Nb, "H": H, "fwd_us": round(us_f, 1), "fwd_us_per_step": round(us_f / T, 2)}
        if T > 1:
            dout = torch.randn(T, Nb, H, device=dev) / (T * Nb) ** 0.5
            dgates = torch.empty(T, Nb, 4 * H, device=dev)
            bwd = lambda: ops.lstm_seq_bwd(dout, private, done, wp, dgates)
            us_b = timeit(bwd, args.reps)
            line.update(bwd_us=round(us_b, 1), bwd_us_per_step=round(us_b / (T - 1), 2))
        print(json.dumps(line), flush=True)
lib.cusrl_b200_lstm_seq_set_debug(0)
