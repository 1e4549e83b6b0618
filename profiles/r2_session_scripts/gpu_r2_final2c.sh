#!/usr/bin/env bash
# round 2, 2 GPUs: generic conv models across GPUs (graph captures confined to init_nodes), config 5 on 2 GPUs
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout -s KILL 900 python -m pytest tests/test_multirank.py -m gpu -q --timeout 600 --timeout-method=thread -k "generic_conv or cpp_executor" -p no:cacheprovider > gpurun_out/pytest_mr_generic.log 2>&1; echo "generic multirank tests rc=$?"; tail -5 gpurun_out/pytest_mr_generic.log | cut -c1-300
: > gpurun_out/config5_n2.jsonl
run 600 29891 benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2.err | grep "^{" >> gpurun_out/config5_n2.jsonl; echo "config 5 N=2 rc=$?"
timeout -s KILL 400 python benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n1.err | grep "^{" >> gpurun_out/config5_n2.jsonl; echo "config 5 N=1 rc=$?"
cut -c1-420 gpurun_out/config5_n2.jsonl; tail -3 gpurun_out/cfg5_n2.err | cut -c1-200
