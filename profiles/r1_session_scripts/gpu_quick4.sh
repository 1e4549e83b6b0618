#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "bank or native_scheduler or simulation_on_gpu" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -20 gpurun_out/pytest_quick.log | cut -c1-300
for impl in banked events; do timeout 300 python benchmarks/many_nodes.py --nodes 4141 --rounds 10 --impl $impl > gpurun_out/many_$impl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/many_$impl.log | cut -c1-500; done
timeout 500 python benchmarks/many_nodes.py --nodes 4141 --rounds 2 --impl reference > gpurun_out/many_reference.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/many_reference.log | cut -c1-500
