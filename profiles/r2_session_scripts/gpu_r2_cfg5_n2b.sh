#!/usr/bin/env bash
# round 2, 2 GPUs: config 5 on 2 ranks with the multi-rank default (graphs captured in init_nodes, NCHW rows), long run, twice
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
: > gpurun_out/config5_n2.jsonl
for i in 1 2; do
  run 400 $((29890+i)) benchmarks/baseline_configs.py --config 5 --rounds 12 --warmup 6 2> gpurun_out/cfg5_n2_$i.err | grep "^{" >> gpurun_out/config5_n2.jsonl; echo "config 5 N=2 run $i rc=$?"
  grep -h "RuntimeError:" gpurun_out/cfg5_n2_$i.err | head -1 | cut -c1-120
done
cut -c1-420 gpurun_out/config5_n2.jsonl
