#!/usr/bin/env bash
# quick kernel iteration: tc2 tests + micro timers (1 GPU, ~1.5 min)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "tc or fused or handshake or stage" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_quick.log | cut -c1-300
timeout 200 python benchmarks/micro.py train --impl tc2 > gpurun_out/micro_tc2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/micro_tc2.log | cut -c1-900
