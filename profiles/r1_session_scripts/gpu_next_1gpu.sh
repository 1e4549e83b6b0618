#!/usr/bin/env bash
# FIRST GPU call of the next round (1 GPU, ~40 s on the box): does the C++ executor (FIFO slots) match / beat the
# Python executor on the headline?  Also re-validates the GPU test suite of the final round-1 tree.
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu_next.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_next.log | cut -c1-300
for ex in python native; do
  timeout -s KILL 200 python bench.py --steps 50 --warmup 3 --executor $ex > gpurun_out/bench_next_$ex.json 2> gpurun_out/bench_next_$ex.err; echo "bench $ex rc=$?"; tail -1 gpurun_out/bench_next_$ex.json | cut -c1-200; grep -i "error\|Traceback" gpurun_out/bench_next_$ex.err | head -3
done
