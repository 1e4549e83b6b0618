#!/usr/bin/env bash
# Third GPU contact (1 GPU): validate the tc2 training kernel + loader, phase timers, bench, ncu.
set -u
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== micro train =="
for impl in tc2 tc; do timeout 200 python benchmarks/micro.py train --impl $impl > gpurun_out/micro_$impl.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/micro_$impl.log | cut -c1-900; done
timeout 200 python benchmarks/micro.py overlap --impl tc2 > gpurun_out/micro_overlap.log 2>&1; tail -2 gpurun_out/micro_overlap.log
echo "== bench N=1 =="
timeout 600 python bench.py --steps 20 --warmup 3 --curve > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-2500; tail -3 gpurun_out/bench_n1.err
echo "== ncu full: tc2 + stage =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp1_train_tc2|mlp1_stage" -s 6 -c 2 -f -o gpurun_out/prof_train_tc2 python benchmarks/micro.py train --impl tc2 > gpurun_out/ncu_train2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_train2.log
