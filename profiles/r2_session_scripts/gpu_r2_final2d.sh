#!/usr/bin/env bash
# round 2, 2 GPUs, final: whole GPU suite on 2 GPUs, config 5 on 2 ranks with the defaults, bench N=2
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout -s KILL 1700 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest gpu (2 GPUs) rc=$?"; tail -5 gpurun_out/pytest_gpu_2gpu.log | cut -c1-300
run 400 29891 benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2.err | grep "^{" > gpurun_out/config5_n2.jsonl; echo "config 5 N=2 rc=$?"; cut -c1-300 gpurun_out/config5_n2.jsonl; grep -h "RuntimeError:" gpurun_out/cfg5_n2.err | head -1 | cut -c1-120
run 300 29881 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-200
