#!/usr/bin/env bash
# round 2, 4 GPUs, final tree: bench N=4
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29881 bench.py --gpus 4 --steps 40 --warmup 3 --no-tf32 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench N=4 rc=$?"; tail -1 gpurun_out/bench_n4.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'], d['test_acc_by_round_tail'][-2:])"
