#!/usr/bin/env bash
# round 2, 1 GPU, final tree: ncu capture of the pipelined tc8 kernel, GPU suite, smoke, bench
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout -s KILL 500 $NCU --set full --import-source on -k regex:mlp1_train_tc4 -s 2 -c 2 -o gpurun_out/ncu_tc8_pipe -f python benchmarks/check_tc4.py ncuonly > gpurun_out/ncu_tc8_pipe.log 2>&1; echo "ncu tc8 rc=$?"; tail -2 gpurun_out/ncu_tc8_pipe.log
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, 'e2e', d['e2e']['value'], 'tf32', d['tf32']['value'])"
