#!/usr/bin/env bash
# round 2, 1 GPU, final: what the driver runs at round end -- GPU suite, smoke(), bench.py both arms
set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench reference rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-260
timeout -s KILL 300 python bench.py > gpurun_out/bench_default_flags.json 2> gpurun_out/bench_default_flags.err; echo "bench (no flags) rc=$?"; tail -1 gpurun_out/bench_default_flags.json | cut -c1-400
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'dtype', 'scaling', 'vs_baseline', 'clocks')}, 'e2e', d['e2e'], 'tf32', d['tf32']['value'])"
