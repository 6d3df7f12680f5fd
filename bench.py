"""Benchmark of the B200 PPO hot path (contract: see the task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--envs 65536] [--rollout 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one PPO iteration on synthetic Anymal-C-rough-shaped data: 24 x (agent.act + agent.step) +
agent.update() (5 epochs x 4 minibatches), i.e. exactly the sections the reference times as "agent"
(cusrl/template/trainer.py:296-321, Perf/agent_fps at :385-393).  metric = env-steps/s = rollout * envs_global /
time.  Strong scaling: the 65536 environments of BASELINE.json are split over the ranks.

Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

OBS, ACT = 235, 12


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mlp", choices=["mlp", "lstm", "rnd"],
                    help="BASELINE.json configuration family: mlp = MLP 512-256-128 PPO (configs 2 and 5; the default 65536 envs "
                         "is the headline), lstm = recurrent PPO, LSTM 2 x 256 (config 3: --envs 4096), rnd = MLP PPO + RND "
                         "intrinsic reward (config 4: --envs 16384)")
    ap.add_argument("--envs", type=int, default=65536, help="global number of environments")
    ap.add_argument("--rollout", type=int, default=24)
    ap.add_argument("--cpu-envs", type=int, default=4096, help="environments of the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference-on-cuda:0 baseline leg")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of timed CPU iterations")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def measured_peaks() -> tuple[dict, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return json.loads(f.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(key: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel at exactly this problem size, from
    the committed `ncu --set full` captures (profiles/ncu_traffic.json names the capture each figure comes from)."""
    f = ROOT / "profiles" / "ncu_traffic.json"
    if not f.exists():
        return None
    entry = json.loads(f.read_text()).get(key)
    return None if entry is None else entry["dram_bytes"]


class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
class RolloutData:
    """Synthetic env stream: obs ~ N(0,1), reward ~ N(0,1), terminated ~ B(0.01),
    truncated ~ B(0.001); generated once, either resident in HBM or in pinned host memory."""

    def __init__(self, T: int, N: int, device, seed: int, pinned_host: bool):
        g = torch.Generator(device=device).manual_seed(seed)
        self.obs = torch.randn(T + 1, N, OBS, device=device, generator=g)
        self.reward = torch.randn(T, N, 1, device=device, generator=g)
        self.terminated = torch.rand(T, N, 1, device=device, generator=g) < 0.01
        self.truncated = torch.rand(T, N, 1, device=device, generator=g) < 0.001
        if pinned_host:
            for k in ("obs", "reward", "terminated", "truncated"):
                setattr(self, k, getattr(self, k).cpu().pin_memory())
        self.T, self.N = T, N


def run_iteration(agent, data: RolloutData, phases: dict | None = None) -> dict:
    """24 x (act, step) + update: the reference's "agent" timer sections.  `phases` (optional) accumulates where the time of
    the iteration goes: CUDA-event time of the rollout and of the update, and the HOST time spent issuing each of them (up
    to the metrics read-back, the one synchronisation point of an iteration) -- host time close to the event time means the
    phase is bound by launch overhead, not by the GPU."""
    if phases is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        t0 = time.perf_counter()
    for t in range(data.T):
        action = agent.act(data.obs[t])
        ready = agent.step(data.obs[t + 1], data.reward[t], data.terminated[t], data.truncated[t])
    assert ready
    del action
    if phases is None:
        return agent.update()
    t1 = time.perf_counter()
    ev[1].record()
    mark = {}
    summary = agent.metrics.summary

    def timed_summary(*a, **k):
        mark["t"] = time.perf_counter()
        return summary(*a, **k)

    agent.metrics.summary = timed_summary
    try:
        out = agent.update()
    finally:
        agent.metrics.summary = summary
    ev[2].record()
    ev[2].synchronize()
    phases["rollout_ms"] = phases.get("rollout_ms", 0.0) + ev[0].elapsed_time(ev[1])
    phases["update_ms"] = phases.get("update_ms", 0.0) + ev[1].elapsed_time(ev[2])
    phases["host_rollout_issue_ms"] = phases.get("host_rollout_issue_ms", 0.0) + (t1 - t0) * 1e3
    phases["host_update_issue_ms"] = phases.get("host_update_issue_ms", 0.0) + (mark.get("t", t1) - t1) * 1e3
    phases["iterations"] = phases.get("iterations", 0) + 1
    return out


def time_iterations(agent, data, steps: int, warmup: int, distributed: bool) -> tuple[float, dict]:
    """(seconds for `steps` iterations as the max over ranks, last metrics); device-timed with CUDA events."""
    import torch.distributed as dist

    for _ in range(warmup):
        metrics = run_iteration(agent, data)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        metrics = run_iteration(agent, data)
    end.record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    seconds = start.elapsed_time(end) * 1e-3
    if distributed:
        t = torch.tensor([seconds], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = t.item()
    return seconds, metrics


# ------------------------------------------------------------------------------------------------
def gae_roofline(T: int, N: int, peaks: dict, which: str, chain: bool = False) -> dict:
    """Achieved HBM bandwidth of the GAE scan kernel, timed live: a CUDA graph of launches over rotating buffer
    sets whose footprint exceeds L2, CUDA events on the launching stream.  chain=False: the separable K1 kernel (21 B per
    element: reward, value, next_value, done in; advantage, return out).  chain=True: the kernel the pre-update actually
    runs, K3 + K1 + K2 statistics in one launch (22 B: reward, value, terminated, truncated in; next_value, advantage,
    return out) -- the same work the reference does in three stages moving 10 + 21 + 4 = 35 B per element."""
    from cusrl_b200 import _lib, ops

    E = T * N
    bytes_per_elt = 22 if chain else 21
    n_sets = max(2, int(300e6 // (bytes_per_elt * E)) + 1)
    sets = []
    for _ in range(n_sets):
        d = {k: torch.randn(T, N, 1, device="cuda") for k in ("reward", "value", "nv", "adv", "ret")}
        d["done"] = torch.rand(T, N, 1, device="cuda") < 0.011
        d["trunc"] = torch.rand(T, N, 1, device="cuda") < 0.001
        d["boot"] = torch.randn(N, 1, device="cuda")
        d["mean_var"] = torch.empty(2, device="cuda")
        sets.append(d)

    def launch_all():
        for d in sets:
            if chain:
                ops.gae_chain(d["reward"], d["done"], d["trunc"], d["value"], d["boot"], 0.99, 0.95, None, 0.0, d["nv"], d["adv"],
                              d["ret"], mean_var=d["mean_var"])
            else:
                ops.gae(d["reward"], d["done"], d["value"], d["nv"], 0.99, 0.95, advantage=d["adv"], ret=d["ret"])

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        launch_all()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        launch_all()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e-3 / n_sets)
    sec = sum(times) / len(times)
    achieved = bytes_per_elt * E / sec / 1e9
    peak = float(peaks["hbm_gbs"])
    if chain:
        return {"kernel": "gae_chain_kernel (K3 next_value + K1 GAE scan + K2 statistics, one launch; includes its one-block "
                          "finalisation)", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(f"gae_chain_f32@T={T},N={N}"), "peak_source": which,
                "bytes_per_launch": bytes_per_elt * E, "us_per_launch": round(sec * 1e6, 2),
                "reference_stage_bytes_per_launch": 35 * E,
                "vs_reference_stage_bytes": round(35 * E / sec / 1e9 / peak, 4)}
    return {"kernel": "gae_kernel", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4),
            "traffic": ncu_traffic(f"gae_f32@T={T},N={N},variant={_lib._gae_variant_from_env()[0]}"), "peak_source": which,
            "bytes_per_launch": 21 * E, "us_per_launch": round(sec * 1e6, 2),
            "variant": dict(zip(("variant", "warps", "stages", "ctas_per_sm"), _lib._gae_variant_from_env()))}


def gemm_roofline(rows: int, peaks: dict, which: str) -> dict:
    """The dominant kernel of the step (profiles/r02_iter_profile.md): the f16x3 tcgen05 dense-layer forward
    `gemm_f16x3_kernel<256,0,1>` (fp16 hi/lo split operands, pair output) on the 512 -> 256 trunk layer of one minibatch,
    timed live with CUDA events over back-to-back launches on rotating activation buffers larger than L2.
    achieved = ALGORITHMIC flops (2 M N K, fp32 semantics) / time; the tensor pipe executes 3x that as fp16 MMAs.
    peak = measured dense bf16 cuBLAS throughput (fp16 and bf16 MMAs run at the same rate); the sustained figure because
    the kernel runs inside a long step."""
    from cusrl_b200 import ops

    M, K, N = rows, 512, 256
    n_sets = max(2, int(300e6 // (M * (K + N) * 4)) + 1)
    xs = [ops.split_f16(torch.randn(M, K, device="cuda")) for _ in range(n_sets)]
    ys = [ops.pair_empty(M, N, "cuda") for _ in range(n_sets)]
    w = torch.randn(N, K, device="cuda") / K**0.5
    b = torch.randn(N, device="cuda")
    wp = ops.weight_prep_f16(w, b)
    for x, y in zip(xs, ys):
        ops.f16_linear_fwd(x, wp, b, 1, True, out=y)
    torch.cuda.synchronize()
    reps = 4
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for x, y in zip(xs, ys):
            ops.f16_linear_fwd(x, wp, b, 1, True, out=y)
    e.record()
    torch.cuda.synchronize()
    sec = a.elapsed_time(e) * 1e-3 / (reps * n_sets)
    flops = 2.0 * M * N * K
    achieved = flops / sec / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    return {"kernel": "gemm_f16x3_kernel<256,0,1> (f16x3 forward, 512->256, fp16 hi/lo pair in and out)", "bound": "tensor",
            "achieved": round(achieved, 1), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "traffic": ncu_traffic(f"gemm_f16x3_kernel<256,0,1>@M={M},K={K},N={N}"),
            "peak_source": f"{which}: bf16_tflops_sustained (fp16 MMA rate)", "flops_per_launch": flops,
            "executed_f16_flops_per_launch": 3 * flops, "executed_frac": round(3 * achieved / peak, 4),
            "hbm_gbs": round(M * (K + N) * 4 / sec / 1e9, 1), "us_per_launch": round(sec * 1e6, 2), "rows": M}


# ------------------------------------------------------------------------------------------------
# The reference arm: the UNMODIFIED reference package (baseline/_ref, see tools/install_reference.py) driven through its
# own public API -- cusrl.preset.ppo.PpoAgentFactory with the Isaac-Velocity-Rough-Anymal-C-v0 values of
# cusrl/zoo/isaaclab/locomotion.py:48-59, agent.act / agent.step / agent.update exactly as cusrl/template/trainer.py:296-321
# calls them.  None of this repository's kernels, modules or oracle code is on that path.
ANYMAL_C_ROUGH = dict(num_steps_per_update=24, actor_hidden_dims=(512, 256, 128), critic_hidden_dims=(512, 256, 128),
                      activation_fn="ELU", lr=1e-3, sampler_epochs=5, sampler_mini_batches=4, orthogonal_init=False,
                      entropy_loss_weight=0.005, desired_kl_divergence=0.015)


def _import_reference():
    sys.path.insert(0, str(ROOT / "tools"))
    from install_reference import import_reference

    # the reference reads the torchrun variables at import (cusrl/utils/config.py:31-38) and would open a process group
    # on its first collective; this arm runs on ONE process (rank 0), so it must not see them
    for key in ("LOCAL_RANK", "RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE"):
        os.environ.pop(key, None)
    return import_reference()


class ReferenceRun:
    """The reference agent on `device` ("cpu" or "cuda:0") with `envs` synthetic environments, same data recipe as
    :class:`RolloutData`."""

    def __init__(self, device: str, envs: int, T: int, seed: int = 1000, config: str = "mlp"):
        cusrl = _import_reference()
        from cusrl.template.environment import EnvironmentSpec

        self.device, self.envs, self.T = torch.device(device), envs, T
        torch.manual_seed(42)
        spec = EnvironmentSpec(num_instances=envs, observation_dim=OBS, action_dim=ACT, reward_dim=1, autoreset=True,
                               final_state_is_missing=True)
        if config == "lstm":   # RecurrentPpoAgentFactory defaults: LSTM 2 x 256 both nets, lr 2e-4 (preset/ppo.py:185-245)
            factory = cusrl.preset.ppo.RecurrentPpoAgentFactory(device=device, num_steps_per_update=T)
        else:
            factory = cusrl.preset.ppo.PpoAgentFactory(device=device, **dict(ANYMAL_C_ROUGH, num_steps_per_update=T))
        if config == "rnd":    # cusrl_test/hook/auxiliary/test_rnd.py:12-19
            factory = factory.to_underlying()
            factory.register_hook(cusrl.hook.RandomNetworkDistillation(cusrl.Mlp.Factory([128, 128]), output_dim=16, reward_scale=0.1),
                                  before="value_computation")
        self.agent = factory(spec)
        self.data = RolloutData(T, envs, self.device, seed=seed, pinned_host=False)

    def iteration(self) -> dict:
        return run_iteration(self.agent, self.data)

    def time(self, iters: int, warmup: int, budget_s: float | None = None) -> tuple[float, float, int]:
        """(env-steps/s, seconds per iteration, timed iterations).  CUDA events on the GPU like the reference's Timer
        (cusrl/utils/timing.py:49-94), perf_counter on the CPU (:32-46).  `budget_s` bounds the timed iterations by the
        duration of the last warm-up iteration so the whole call ends within minutes."""
        cuda = self.device.type == "cuda"
        last = None
        for _ in range(max(warmup, 1)):
            t0 = time.perf_counter()
            self.iteration()
            if cuda:
                torch.cuda.synchronize()
            last = time.perf_counter() - t0
        if budget_s is not None:
            iters = max(1, min(iters, int(budget_s / max(last, 1e-9))))
        if cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        t0 = time.perf_counter()
        for _ in range(iters):
            self.iteration()
        if cuda:
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b) * 1e-3 / iters
        else:
            dt = (time.perf_counter() - t0) / iters
        return self.T * self.envs / dt, dt, iters


def make_b200_agent(C, config: str, device, env):
    """The agent of a BASELINE.json configuration family through this repository's public factories."""
    if config == "lstm":
        return C.RecurrentPpoAgentFactory(device=device).from_environment(env)
    factory = C.anymal_c_rough_ppo(device=device)
    if config == "rnd":
        factory = factory.to_underlying()
        factory.register_hook(C.RandomNetworkDistillation(C.Mlp.Factory([128, 128]), output_dim=16, reward_scale=0.1),
                              before="value_computation")
    return factory.from_environment(env)


def pick_cpu_threads(T: int, max_threads: int, envs: int = 1024, config: str = "mlp") -> tuple[int, dict]:
    """PyTorch-CPU does not scale to every hardware thread on this workload (128 threads were 20x slower than 32 on the
    GPU box's host): calibrate on a small sample (one warm-up + three timed reference iterations per setting) and use the
    fastest setting, so the baseline is the CPU at its best."""
    run = ReferenceRun("cpu", envs if config != "lstm" else min(envs, 256), T, config=config)
    rates = {}
    for th in (8, 16, 32, 64, max_threads):
        if th > max_threads or th in rates:
            continue
        torch.set_num_threads(th)
        rates[th], _, _ = run.time(iters=3, warmup=1)
    best = max(rates, key=rates.get)
    torch.set_num_threads(best)
    return best, {str(k): round(v, 1) for k, v in rates.items()}


def reference_cpu_baseline(envs: int, T: int, iters: int, warmup: int, host_threads: int, budget_s: float | None,
                           config: str = "mlp"):
    """`cpu_baseline` object: the reference's own CPU path on this box's host cores."""
    threads, calib = pick_cpu_threads(T, host_threads, config=config)
    run = ReferenceRun("cpu", envs, T, config=config)
    rate, dt, done = run.time(iters, warmup, budget_s)
    return {"value": round(rate, 1), "unit": "env-steps/s", "cores": threads, "kind": "reference",
            "sample": f"{envs} envs x {T} steps per iteration, {max(warmup, 1)} warm-up + {done} timed iterations "
                      f"({dt:.2f} s each); unmodified reference (baseline/_ref) on device='cpu'",
            "thread_calibration_env_steps_per_s": calib}, dt, done


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    T = args.rollout
    host_threads = os.cpu_count() or 1

    if args.impl == "reference":
        # The reference arm of this tier: the UNMODIFIED reference (baseline/_ref) through its own act/step/update on the
        # box's host cores, rank 0 only, on the SAME workload (all args.envs environments).  One CPU iteration at 65536
        # environments takes tens of seconds, so the number of timed iterations is bounded by --ref-budget seconds
        # (never fewer than one); `steps` reports the iterations actually timed, `steps_requested` what was asked for.
        if rank != 0:
            return
        cpu, dt, done = reference_cpu_baseline(args.envs, T, iters=max(1, args.steps), warmup=1,
                                               host_threads=host_threads, budget_s=args.ref_budget, config=args.config)
        rate = cpu["value"]
        line = {
            "impl": "reference", "metric": "ppo_env_steps_per_sec", "value": rate, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": done, "warmup": 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "cpu_baseline": cpu,
            "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return

    import cusrl_b200 as C
    from cusrl_b200 import ops
    from cusrl_b200.runtime import CONFIG, configure_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (cusrl_b200 has no CPU fallback)")
    distributed = world > 1
    if distributed:
        configure_distributed()
    device = CONFIG.device
    if args.envs % world:
        raise SystemExit("--envs must be divisible by the number of ranks")
    N = args.envs // world
    peaks, which = measured_peaks()

    torch.manual_seed(42 + rank)  # per-rank seed like the reference (utils/misc.py:163)
    env = C.SyntheticEnvironment(N, OBS, ACT, device=device, seed=42 + rank)
    agent = make_b200_agent(C, args.config, device, env)

    # ---- value: inputs resident in HBM
    data = RolloutData(T, N, device, seed=1000 + rank, pinned_host=False)
    launches0 = ops.launch_count()
    with ClockSampler(device.index or 0) as clocks:
        seconds, metrics = time_iterations(agent, data, args.steps, args.warmup, distributed)
    launches = (ops.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    value = args.steps * T * args.envs / seconds

    # ---- e2e: the same iterations driven through agent.act / agent.step with HOST (pinned) buffers
    data_for_phases = data
    e2e = None
    if not args.no_e2e:
        del data
        host = RolloutData(T, N, device, seed=1000 + rank, pinned_host=True)
        e_steps = max(1, min(args.steps, 3))
        e_sec, _ = time_iterations(agent, host, e_steps, 1, distributed)
        # every observation crosses PCIe ONCE: step(t) uploads next_obs[t]; the following act(t+1) receives the same host
        # array and reads the device copy (template/rollout.py); only the first act of an iteration uploads its own
        per_iter_in = (T + 1) * N * OBS * 4 + T * N * (4 + 1 + 1)
        per_iter_out = T * N * ACT * 4 + 4 * 16
        e2e = {"value": round(e_steps * T * args.envs / e_sec, 1), "unit": "env-steps/s",
               "h2d_bytes_per_step": per_iter_in * world, "d2h_bytes_per_step": per_iter_out * world,
               "steps": e_steps, "note": "pinned host tensors through agent.act/agent.step; metrics read back per update"}
        del host

    # ---- where the time goes (2 extra iterations, separately timed): device time per phase vs host issue time per phase
    phases: dict = {}
    if data_for_phases is not None:
        for _ in range(2):
            run_iteration(agent, data_for_phases, phases)
        n_it = phases.pop("iterations")
        phases = {k: round(v / n_it, 3) for k, v in phases.items()}
        if distributed:   # the slowest rank per phase
            t = torch.tensor([phases[k] for k in sorted(phases)], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            phases = {k: round(v, 3) for k, v in zip(sorted(phases), t.tolist())}
        del data_for_phases

    line = None
    if rank == 0:
        roof_gae = gae_roofline(T, N, peaks, which)
        # the same kernel on 16x the columns: a 33 MB launch lasts ~9 us, of which ~2.6 us are launch ramp and drain that
        # any kernel of this size pays (a device copy of the same bytes reaches 0.66 of the 1 GiB copy rate); the larger
        # launch shows the kernel's streaming rate without that fixed cost.  Context only: `frac` above is the metric.
        big = gae_roofline(T, 16 * N, peaks, which)
        roof_gae["same_kernel_16x_columns"] = {k: big[k] for k in ("achieved", "frac", "bytes_per_launch", "us_per_launch")}
        roof_gae["pre_update_chain"] = gae_roofline(T, N, peaks, which, chain=True)
        roof = gemm_roofline(T * N // 4, peaks, which)  # one minibatch of this rank (4 minibatches per epoch)
        # ---- baselines (N = 1 only), both the UNMODIFIED reference from baseline/_ref, never on the product path
        ref_cuda = None
        if not args.no_reference_cuda and world == 1:
            # BASELINE.json's >= 10x denominator: "the reference's own 1-GPU PyTorch PPO" = the reference code with
            # device="cuda" on this GPU, same workload (all args.envs environments), CUDA-event timed like its Timer
            del agent, env
            torch.cuda.empty_cache()
            try:
                run = ReferenceRun("cuda:0", args.envs, T, config=args.config)
                rate, dt, done = run.time(iters=3, warmup=2)
                ref_cuda = {"value": round(rate, 1), "unit": "env-steps/s", "kind": "reference", "ms_per_step": round(dt * 1e3, 2),
                            "speedup_value": round(value / rate, 2),
                            "speedup_e2e": None if e2e is None else round(e2e["value"] / rate, 2),
                            "note": "unmodified reference (baseline/_ref): " + {
                                "mlp": "cusrl.preset.ppo.PpoAgentFactory with the locomotion.py:48-59 values",
                                "lstm": "cusrl.preset.ppo.RecurrentPpoAgentFactory defaults (nn.LSTM / cuDNN 2 x 256)",
                                "rnd": "PpoAgentFactory (locomotion.py:48-59 values) + RandomNetworkDistillation before "
                                       "value_computation"}[args.config]
                                    + f" on cuda:0, {args.envs} envs x {T} steps, same act/step/update calls and data, "
                                      f"2 warm-up + {done} timed iterations, CUDA events"}
                del run
                torch.cuda.empty_cache()
            except Exception as error:  # a baseline must never take the bench line down with it
                ref_cuda = {"value": None, "error": f"{type(error).__name__}: {error}"[:300]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu_envs = min(args.cpu_envs, args.envs) if args.config != "lstm" else min(args.envs, 256)
                cpu, _, _ = reference_cpu_baseline(cpu_envs, T, iters=3, warmup=1, host_threads=host_threads, budget_s=30.0,
                                                   config=args.config)
            except Exception as error:
                cpu = {"value": None, "error": f"{type(error).__name__}: {error}"[:300]}
        line = {
            "metric": "ppo_env_steps_per_sec", "value": round(value, 1), "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(seconds / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "e2e": e2e, "gpu_launches": int(launches), "phases": phases,
            "clocks": clocks.summary(), "roofline": roof, "roofline_gae": roof_gae, "cpu_baseline": cpu, "reference_cuda": ref_cuda,
            "last_metrics": {k: round(v, 6) for k, v in metrics.items() if k.startswith("Agent/")},
        }
    if distributed:
        torch.distributed.barrier()
    if line is not None:
        print(json.dumps(line), flush=True)


def workload_config(args, world: int) -> dict:
    model = {"mlp": "MLP 512-256-128 ELU actor-critic PPO", "lstm": "LSTM 2 x 256 recurrent actor-critic PPO (RecurrentPpoAgentFactory)",
             "rnd": "MLP 512-256-128 ELU actor-critic PPO + RND intrinsic reward (235-128-128-16 nets)"}[args.config]
    return {
        "workload": f"synthetic Anymal-C-rough-shaped obs/act (235/12), {model}, "
                    f"{args.envs} envs x {args.rollout} steps, 5 epochs x 4 minibatches",
        "envs_global": args.envs, "envs_per_rank": args.envs // max(world, 1), "rollout_steps": args.rollout,
        "parallelism": f"dp{world} (env axis), NCCL flat-gradient allreduce per minibatch",
        "l2_policy": "working set (rollout buffer >= 0.4 GB per rank, minibatches >= 100 MB) exceeds the 126 MB L2",
        "timed_region": "24 x (agent.act + agent.step) + agent.update per step; synthetic env data precomputed",
    }


if __name__ == "__main__":
    main()
