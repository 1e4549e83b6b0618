#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
for N in 4 8; do
  run 200 $N $((29910+N)) bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-330
done
