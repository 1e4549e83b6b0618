#!/usr/bin/env bash
# 1 GPU: final validation with the C++ executor enabled + A/B against the Python executor
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_n.log | cut -c1-300
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_n.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_n.log
for ex in native python; do
  timeout -s KILL 200 python bench.py --steps 50 --warmup 3 --executor $ex > gpurun_out/bench_n_$ex.json 2> gpurun_out/bench_n_$ex.err; echo "bench $ex rc=$?"; tail -1 gpurun_out/bench_n_$ex.json | cut -c1-200; grep -i "error\|Traceback" gpurun_out/bench_n_$ex.err | head -3
done
