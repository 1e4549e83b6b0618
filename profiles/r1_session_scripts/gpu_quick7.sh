#!/usr/bin/env bash
# tc3 fine-grained phase counters + timings (1 GPU, < 1 min)
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "tc3 or train or stage or fused" > gpurun_out/pytest_q7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q7.log | cut -c1-300
timeout 300 python benchmarks/micro.py train --impl tc3 > gpurun_out/micro_q7.log 2>&1; echo "micro rc=$?"; grep "^{" gpurun_out/micro_q7.log | cut -c1-1000
