#!/usr/bin/env bash
# 1 GPU: validation of the final round-1 tree (coalesced TMEM fill / write-back, vectorised loader) + timings
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_l.log | cut -c1-300
timeout 200 python benchmarks/micro.py train --impl tc3 > gpurun_out/micro_l.log 2>&1; echo "micro rc=$?"; grep "^{" gpurun_out/micro_l.log | cut -c1-600
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_l.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_l.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_l_n1.json 2> gpurun_out/bench_l_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_l_n1.json | cut -c1-400
