#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "eval_confusion or simulation_on_gpu or native_scheduler" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_quick.log | cut -c1-300
timeout 200 python benchmarks/micro.py eval > gpurun_out/micro_eval.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/micro_eval.log | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-2200; tail -3 gpurun_out/bench_n1.err
