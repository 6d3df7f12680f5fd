#!/bin/bash
# One GPU-box session (run under gpurun): every step has its own timeout and log under gpurun_out/.
# Usage: bash tools/gpu_session.sh <tag> <step> [<step> ...]
cd "${GRAFT_REPO_ROOT:-.}"
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
step() {  # name timeout cmd...
  local name=$1 limit=$2; shift 2
  local t0=$(date +%s)
  timeout "$limit" "$@" > "$out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc $(( $(date +%s) - t0 ))s" | tee -a "$out/steps.log"
}
for s in "$@"; do
  case $s in
    tma_tests)  step tma_tests 90 python -m pytest tests/test_gae_tma_gpu.py -x -q -m gpu ;;
    gae_tests)  step gae_tests 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k gae ;;
    kbench_gae) step kbench_gae 240 python tools/kbench.py --only gae ;;
    kbench)     step kbench 300 python tools/kbench.py ;;
    ncu_gae)    step ncu_gae 110 ncu --set full --clock-control none --import-source on -k regex:gae -o "$out/gae" -f python tools/ncu_gae.py ;;
    tests)      step tests 900 python -u -m pytest tests -q -m gpu --timeout 300 -rf ;;
    tests_x)    step tests_x 900 python -u -m pytest tests -x -q -m gpu --timeout 300 ;;
    bench_ref_full) step bench_ref_full 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget 90 ;;
    bench_short) step bench_short 600 python bench.py --steps 3 --warmup 3 ;;
    bench)      step bench 300 python bench.py ;;
    bench_ref)  step bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1 ;;
    smoke)      step smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
    configs)    step configs 400 python tools/config_fps.py ;;
    tests_all)  step tests_all 600 python -u -m pytest tests -q -m gpu --timeout 120 ;;
    graphs_tests) step graphs_tests 300 env CUSRL_B200_TEST_GRAPHS=1 python -u -m pytest tests/test_graphs_gpu.py -x -v -m gpu --timeout 120 ;;
    graphs_fps) step graphs_fps 200 bash -c "python tools/config_fps.py --only mlp_ppo_4096 --steps 3 --warmup 2 --no-graphs --no-fused-rollout; python tools/config_fps.py --only mlp_ppo_4096 --steps 3 --warmup 2 --no-graphs; python tools/config_fps.py --only mlp_ppo_4096 --steps 3 --warmup 2" ;;
    f16tests)   step f16tests 400 python -u -m pytest tests/test_gemm_f16x3_gpu.py -q -m gpu --timeout 60 -x -rf -s ;;
    f16bench)   step f16bench 300 python tools/f16x3_bench.py ;;
    tests_p3)   step tests_p3 900 env CUSRL_B200_GEMM_PRECISION=3 python -u -m pytest tests -q -m gpu --timeout 300 -rf ;;
    bench_p3)   step bench_p3 600 env CUSRL_B200_GEMM_PRECISION=3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda ;;
    tests_p2)   step tests_p2 900 env CUSRL_B200_GEMM_PRECISION=2 python -u -m pytest tests -q -m gpu --timeout 300 -rf ;;
    bench_p2)   step bench_p2 600 env CUSRL_B200_GEMM_PRECISION=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda ;;
    configs_p2) step configs_p2 400 env CUSRL_B200_GEMM_PRECISION=2 python tools/config_fps.py ;;
    ncu_f16)    step ncu_f16 280 ncu --set full --clock-control none --import-source on -k regex:f16x3_kernel --launch-skip 4 --launch-count 4 -o "$out/f16x3" -f python tools/ncu_f16.py ;;
    bench_cfgs) step bench_cfgs 900 bash -c "python bench.py --config mlp --envs 4096 --steps 5 --warmup 3; python bench.py --config lstm --envs 4096 --steps 3 --warmup 2; python bench.py --config rnd --envs 16384 --steps 5 --warmup 3" ;;
    bench2)     step bench2 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ;;
    bench8)     step bench8 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 5 ;;
    bench4)     step bench4 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 5 ;;
    bench2ref)  step bench2ref 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference --ref-budget 40 ;;
    fp16probe)  step fp16probe 300 python tools/fp16_split_probe.py ;;
    iter)       step iter 70 python tools/iter_profile.py ;;
    launches)   step launches 230 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches.csv" python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-cuda ;;
    bench_8k)   step bench_8k 300 python bench.py --envs 8192 --steps 10 --warmup 5 --no-cpu-baseline --no-reference-cuda ;;
    bench_1x)   step bench_1x 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda ;;
    hostprof_8k) step hostprof_8k 200 python tools/host_profile.py 8192 ;;
    iter_8k)    step iter_8k 100 python tools/iter_profile.py --envs 8192 --iters 2 ;;
    sym_tests)  step sym_tests 600 python -u -m pytest tests/test_symmetry_gpu.py tests/test_gemm_f16x3_gpu.py tests/test_lstm_gpu.py -q -m gpu --timeout 200 -rf -x ;;
    iter_lstm)  step iter_lstm 150 python tools/iter_profile.py --envs 4096 --iters 1 --config lstm ;;
    lstm_tests) step lstm_tests 400 python -u -m pytest tests/test_lstm_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu --timeout 120 -rf -x -k "lstm or sequence" ;;
    bench_lstm) step bench_lstm 300 python bench.py --config lstm --envs 4096 --steps 3 --warmup 2 --no-cpu-baseline --no-reference-cuda ;;
    lstm_bench) step lstm_bench 200 python tools/lstm_bench.py --debug 3 ;;
    hostprof_lstm) step hostprof_lstm 200 python tools/host_profile.py 4096 lstm ;;
    gemm_tests) step gemm_tests 600 python -u -m pytest tests/test_gemm_f16x3_gpu.py tests/test_gemm_gpu.py tests/test_agent_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu --timeout 200 -rf -x ;;
    ncu_lstm)   step ncu_lstm 280 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_ --launch-skip 3 --launch-count 2 -o "$out/lstm_seq" -f python tools/lstm_bench.py --reps 1 --only-config3 ;;
    rollout_tests) step rollout_tests 400 python -u -m pytest tests/test_rollout_gpu.py tests/test_lstm_gpu.py tests/test_graphs_gpu.py -q -m gpu --timeout 200 -rf -x ;;
    bench_small) step bench_small 200 bash -c "for e in 4096 8192; do python bench.py --config mlp --envs \$e --steps 10 --warmup 5 --no-cpu-baseline --no-reference-cuda; CUSRL_B200_ROLLOUT_STREAMS=1 python bench.py --config mlp --envs \$e --steps 10 --warmup 5 --no-cpu-baseline --no-reference-cuda; done" ;;   # profiles/r02_rollout_host.md: two-stream vs one-stream rollout step
    *) echo "unknown step $s" ;;
  esac
done
cat "$out/steps.log"
