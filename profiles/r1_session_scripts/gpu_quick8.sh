#!/usr/bin/env bash
# tc3 exchange variants: correctness (default, then GB_TC3_VARIANT=1) and timings (1 GPU, ~1 min)
set -u
mkdir -p gpurun_out
K="tc3 or train or stage or fused"
timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "$K" > gpurun_out/pytest_q8a.log 2>&1; echo "pytest v0 rc=$?"; tail -2 gpurun_out/pytest_q8a.log | cut -c1-300
GB_TC3_VARIANT=1 timeout -s KILL 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "$K" > gpurun_out/pytest_q8b.log 2>&1; echo "pytest v1 rc=$?"; tail -4 gpurun_out/pytest_q8b.log | cut -c1-300
for v in 0 1; do
  MICRO_TC3_VARIANTS=$v timeout -s KILL 120 python benchmarks/micro.py train --impl tc3 > gpurun_out/micro_q8_$v.log 2>&1; echo "micro v$v rc=$?"; grep "^{" gpurun_out/micro_q8_$v.log | grep -v "CTA 1\|mlp1_stage\|phase cycles per step (thread 0)\"" | cut -c1-1000
done
