"""Stand-alone correctness / speed probe of the tcgen05 dense-layer kernels (run under `timeout` on the GPU box)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cusrl_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def check(M, K, N, act, precision, with_bias=True):
    x = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    w = torch.randn(N, K, device=dev) / K**0.5
    b = torch.randn(N, device=dev) if with_bias else None
    wp = ops.weight_prep(w)
    y = ops.tc_linear_fwd(x, wp, b, N, act, precision)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(x.double(), w.double(), None if b is None else b.double())
    if act == 1:
        ref = torch.nn.functional.elu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    f32 = torch.nn.functional.linear(x, w, b)
    if act == 1:
        f32 = torch.nn.functional.elu(f32)
    elif act == 2:
        f32 = torch.relu(f32)
    err32 = (f32.double() - ref).abs().max().item()
    print(f"fwd  M={M} K={K} N={N} act={act} p={precision}: max|err|={err:.3e} (cuBLAS fp32: {err32:.3e}) scale={scale:.2f}", flush=True)
    return err, err32


def check_dgrad(M, N, K, act, precision):
    dy = torch.randn(M, N, device=dev)
    w = torch.randn(N, K, device=dev) / N**0.5
    xa = torch.randn(M, K, device=dev)
    wp = ops.weight_prep(w)
    dx = ops.tc_linear_dgrad(dy, wp, xa, K, act, precision)
    torch.cuda.synchronize()
    ref = dy.double() @ w.double()
    if act == 1:
        ref = ref * torch.where(xa > 0, torch.ones_like(xa), xa + 1).double()
    err = (dx.double() - ref).abs().max().item()
    print(f"dgrad M={M} N={N} K={K} act={act} p={precision}: max|err|={err:.3e} scale={ref.abs().max().item():.2f}", flush=True)


def speed(M, K, N, precision, reps=20):
    x = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    w = torch.randn(N, K, device=dev) / K**0.5
    b = torch.randn(N, device=dev)
    wp = ops.weight_prep(w)
    y = torch.empty(M, N, device=dev)
    for _ in range(3):
        ops.tc_linear_fwd(x, wp, b, N, 1, precision, out=y)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.tc_linear_fwd(x, wp, b, N, 1, precision, out=y)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / reps
    fl = 2.0 * M * K * N
    # cuBLAS fp32 for comparison
    for _ in range(3):
        torch.nn.functional.elu(torch.nn.functional.linear(x, w, b))
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        torch.nn.functional.elu(torch.nn.functional.linear(x, w, b))
    e.record()
    torch.cuda.synchronize()
    ms_ref = a.elapsed_time(e) / reps
    print(f"speed M={M} K={K} N={N} p={precision}: {ms*1e3:.1f} us, {fl/ms/1e9:.1f} TFLOP/s(fp32-equiv), "
          f"{(M*K+M*N)*4/ms/1e6:.0f} GB/s | torch fp32 linear+elu {ms_ref*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode in ("all", "small"):
        from cusrl_b200 import _lib
        if len(sys.argv) > 2:
            pass
        check(128, 32, 128, 0, 1, with_bias=False)
        check(128, 32, 128, 0, 3, with_bias=False)
        check(256, 64, 256, 0, 1)
        check(300, 236, 512, 1, 1)
        check(300, 235, 512, 1, 3)
        check(1000, 512, 256, 1, 3)
        check(1000, 256, 128, 1, 3)
        check_dgrad(1000, 128, 256, 1, 3)
        check_dgrad(777, 256, 512, 1, 3)
        check_dgrad(512, 256, 512, 1, 1)
    if mode in ("all", "speed"):
        from cusrl_b200 import _lib
        if len(sys.argv) > 2:
            pass
        for p in (3, 1):
            speed(393216, 235, 512, p)
            speed(393216, 512, 256, p)
            speed(393216, 256, 128, p)

def check_wgrad(M, N, K, precision):
    dz = torch.randn(M, N, device=dev)
    x = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    dw = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    ops.tc_linear_wgrad(dz, x, dw, db, precision)
    torch.cuda.synchronize()
    ref = dz.double().t() @ x.double()
    err = (dw.double() - ref).abs().max().item()
    f32 = (dz.t() @ x).double()
    print(f"wgrad M={M} N={N} K={K} p={precision}: max|err|={err:.3e} (cuBLAS fp32 {(f32-ref).abs().max().item():.3e}) "
          f"scale={ref.abs().max().item():.1f}; db err={(db.double()-dz.double().sum(0)).abs().max().item():.3e}", flush=True)


def check_head(M, K, No):
    h = torch.randn(M, K, device=dev); w = torch.randn(No, K, device=dev) / K**0.5; b = torch.randn(No, device=dev)
    y = ops.head_fwd(h, w, b)
    ref = torch.nn.functional.linear(h.double(), w.double(), b.double())
    dy = torch.randn(M, No, device=dev)
    dw = torch.zeros(No, K, device=dev); db = torch.zeros(No, device=dev)
    dh = ops.head_bwd(dy, h, w, 1, dw, db)
    torch.cuda.synchronize()
    dh_ref = (dy.double() @ w.double()) * torch.where(h > 0, torch.ones_like(h), h + 1).double()
    print(f"head M={M} K={K} No={No}: fwd err={(y.double()-ref).abs().max().item():.2e} dh err={(dh.double()-dh_ref).abs().max().item():.2e} "
          f"dw err={(dw.double()-dy.double().t()@h.double()).abs().max().item():.2e} db err={(db.double()-dy.double().sum(0)).abs().max().item():.2e}", flush=True)


if __name__ == "__main__" and (len(sys.argv) > 1 and sys.argv[1] == "bwd"):
    check_head(1000, 128, 12)
    check_head(5000, 128, 1)
    check_wgrad(1024, 128, 128, 1)
    check_wgrad(1024, 128, 128, 3)
    check_wgrad(4096, 512, 235, 3)
    check_wgrad(5000, 256, 512, 3)
    check_wgrad(5000, 128, 256, 3)
    check_wgrad(393216, 256, 512, 3)
    import time
    dz = torch.randn(393216, 256, device=dev); x = torch.randn(393216, 512, device=dev)
    dw = torch.zeros(256, 512, device=dev); db = torch.zeros(256, device=dev)
    for p in (1, 3):
        for _ in range(2): ops.tc_linear_wgrad(dz, x, dw, db, p)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): ops.tc_linear_wgrad(dz, x, dw, db, p)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"wgrad speed 393216x256x512 p={p}: {dt*1e6:.0f} us, {2*393216*256*512/dt/1e12:.1f} TFLOP/s", flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): dw2 = dz.t() @ x
    torch.cuda.synchronize(); print(f"torch fp32 wgrad: {(time.perf_counter()-t0)/10*1e6:.0f} us")
