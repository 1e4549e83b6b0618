#!/usr/bin/env bash
# round 2, 1 GPU, last call: BASELINE configs 2 / 3 / 4 with the final kernel
set -u
mkdir -p gpurun_out
: > gpurun_out/baseline_configs_final.jsonl
for c in 2 3 4; do
  timeout -s KILL 100 python benchmarks/baseline_configs.py --config $c --rounds 30 --warmup 5 2> gpurun_out/cfg$c.err | grep "^{" >> gpurun_out/baseline_configs_final.jsonl; echo "config $c rc=$?"
done
cut -c1-200 gpurun_out/baseline_configs_final.jsonl
