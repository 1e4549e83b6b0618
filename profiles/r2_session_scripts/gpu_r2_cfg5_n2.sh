#!/usr/bin/env bash
# round 2, 2 GPUs: config 5 (default settings) on 2 ranks after priming the generic exchange path in init_nodes
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
: > gpurun_out/config5_n2.jsonl
run 400 29891 benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2.err | grep "^{" >> gpurun_out/config5_n2.jsonl; echo "config 5 N=2 rc=$?"
run 400 29892 benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2b.err | grep "^{" >> gpurun_out/config5_n2.jsonl; echo "config 5 N=2 (again) rc=$?"
cut -c1-420 gpurun_out/config5_n2.jsonl; grep -h "RuntimeError" gpurun_out/cfg5_n2.err gpurun_out/cfg5_n2b.err | head -3 | cut -c1-200
