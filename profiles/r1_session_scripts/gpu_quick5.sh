#!/usr/bin/env bash
# 2 GPUs: kernels touched by the coalesced merge pre-pass + bench N=1 / N=2
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "fused or handshake or tc_matches" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_quick.log | cut -c1-300
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-330
