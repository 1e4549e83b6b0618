#!/usr/bin/env bash
# tc3 bring-up: operand-layout probes, tc3 tests, timers (1 GPU)
set -u
mkdir -p gpurun_out
timeout 120 python benchmarks/probe_tc2.py > gpurun_out/probe_tc2.log 2>&1; echo "probe rc=$?"; cat gpurun_out/probe_tc2.log | tail -8
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "tc3 or forward_first or handshake" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_quick.log | cut -c1-300
for impl in tc3 tc2; do timeout 200 python benchmarks/micro.py train --impl $impl > gpurun_out/micro_$impl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/micro_$impl.log | cut -c1-900; done
