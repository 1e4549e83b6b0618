#!/usr/bin/env bash
# round 2, 2 GPUs: multi-rank GPU tests (incl. the executor's all-to-all / sampling / pass-through modes), config 5 and 7 on 2 GPUs, bench N=2
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout -s KILL 1200 python -m pytest tests/test_multirank.py -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_mr_n2.log 2>&1; echo "multirank tests rc=$?"; tail -5 gpurun_out/pytest_mr_n2.log | cut -c1-300
: > gpurun_out/baseline_configs_n2.jsonl
run 500 29891 benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2.err | grep "^{" >> gpurun_out/baseline_configs_n2.jsonl; echo "config 5 N=2 rc=$?"
run 300 29892 benchmarks/baseline_configs.py --config 7 --rounds 30 --warmup 5 2> gpurun_out/cfg7_n2.err | grep "^{" >> gpurun_out/baseline_configs_n2.jsonl; echo "config 7 N=2 rc=$?"
cut -c1-400 gpurun_out/baseline_configs_n2.jsonl; tail -3 gpurun_out/cfg5_n2.err | cut -c1-200
run 300 29881 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-200
