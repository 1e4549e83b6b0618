#!/usr/bin/env bash
# round 2: GPU test suite + headline bench (both executors) on one GPU
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for ex in native python; do
  timeout -s KILL 300 python bench.py --steps ${BENCH_STEPS:-50} --warmup 3 --executor $ex > gpurun_out/bench_$ex.json 2> gpurun_out/bench_$ex.err; echo "bench $ex rc=$?"; tail -1 gpurun_out/bench_$ex.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d[k] for k in ('value', 'ms_per_step', 'dtype', 'gpu_launches')}, 'e2e', d['e2e'] and d['e2e']['value'], 'tf32', d.get('tf32') and d['tf32']['value'], d['details']['executor'], d['test_acc_by_round_tail'])
except Exception as e:
    print('bad json', e)
"; grep -i "error\|Traceback" gpurun_out/bench_$ex.err | head -3
done
