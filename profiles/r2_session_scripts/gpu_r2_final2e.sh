#!/usr/bin/env bash
# round 2, 2 GPUs, final tree (pipelined tc8 kernel): multi-rank MLP tests, bench N=2
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout -s KILL 900 python -m pytest tests/test_multirank.py -m gpu -q --timeout 600 --timeout-method=thread -k "two_ranks_cuda_equal_single_gpu and not pens and not nccl and not banked and not generic" -p no:cacheprovider > gpurun_out/pytest_mr_final.log 2>&1; echo "multirank tests rc=$?"; tail -3 gpurun_out/pytest_mr_final.log | cut -c1-300
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29881 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'], d['test_acc_by_round_tail'][-2:])"
