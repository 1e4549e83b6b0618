#!/usr/bin/env bash
# round 2, 1 GPU, last call: executor GPU test incl. the momentum / cache-neighbour modes, bench
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_stream_executor.py tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 200 python bench.py --steps 100 --warmup 5 --no-tf32 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-200
