#!/usr/bin/env bash
# round 2, one GPU: kernel table (all training kernels incl. fused momentum), graph-replay test repeated, config 3 with pipelined read-back
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python benchmarks/check_tc4.py tc8 tc8-tf32 tc3 cluster > gpurun_out/check_tc4_all.log 2>&1; echo "check_tc4 rc=$?"; grep "timing" gpurun_out/check_tc4_all.log | cut -c1-200
for i in 1 2 3 4 5; do
  timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cuda_graph or keyed_perm" -p no:cacheprovider > gpurun_out/pytest_graph_$i.log 2>&1; echo "graph tests run $i rc=$?"; tail -2 gpurun_out/pytest_graph_$i.log | cut -c1-300
  grep -h "Error\|error" gpurun_out/pytest_graph_$i.log | head -5 | cut -c1-400
done
: > gpurun_out/baseline_configs_n1.jsonl
for c in 3 2; do
  timeout -s KILL 300 python benchmarks/baseline_configs.py --config $c --rounds 30 --warmup 5 2> gpurun_out/cfg$c.err | grep "^{" >> gpurun_out/baseline_configs_n1.jsonl; echo "config $c rc=$?"
done
cut -c1-300 gpurun_out/baseline_configs_n1.jsonl
