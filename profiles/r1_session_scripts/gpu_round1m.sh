#!/usr/bin/env bash
# 2 GPUs: cross-GPU tests (incl. PENS across ranks) + bench N=2 with the final round-1 kernels
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout -s KILL 400 python -m pytest tests/test_multirank.py -m gpu -q --timeout 300 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr_m.log 2>&1; echo "mr rc=$?"; tail -3 gpurun_out/pytest_mr_m.log | cut -c1-300
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_m_n2.json 2> gpurun_out/bench_m_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_m_n2.json | cut -c1-330
