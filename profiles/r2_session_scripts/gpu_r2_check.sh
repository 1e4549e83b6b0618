#!/usr/bin/env bash
# round 2: GPU test-suite, smoke(), launch list of two benchmark rounds, sanitizer passes
set -u
mkdir -p gpurun_out
timeout -s KILL 700 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout -s KILL 400 ncu --clock-control none --metrics gpu__time_duration.sum -s 120 -c 100 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-tf32 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"; grep -c "mlp1_train" gpurun_out/launches_r2.csv
if [ "${SANITIZE:-1}" = "1" ]; then bash tools/sanitize.sh; fi
