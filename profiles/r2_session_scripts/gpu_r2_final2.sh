#!/usr/bin/env bash
# round 2, 2 GPUs: the whole GPU test-suite incl. the multi-rank tests, headline at N=2
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout -s KILL 1700 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest gpu (2 GPUs) rc=$?"; tail -6 gpurun_out/pytest_gpu_2gpu.log | cut -c1-300
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29881 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-200
