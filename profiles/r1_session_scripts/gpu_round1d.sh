#!/usr/bin/env bash
# 1 GPU: full gpu tests, bench (native engine vs python engine), ncu of tc2 + launch list
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench N=1 native engine =="
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-2200; tail -3 gpurun_out/bench_n1.err
echo "== bench N=1 python engine =="
timeout 600 python bench.py --steps 30 --warmup 3 --engine python --no-e2e > gpurun_out/bench_n1_py.json 2> gpurun_out/bench_n1_py.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1_py.json | cut -c1-400
echo "== micro =="
timeout 300 python benchmarks/micro.py all --impl tc3 > gpurun_out/micro_all.log 2>&1; tail -14 gpurun_out/micro_all.log | cut -c1-700
echo "== ncu launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full tc2 =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp1_train_tc2|mlp1_stage" -s 6 -c 2 -f -o gpurun_out/prof_train_tc3 python benchmarks/micro.py train --impl tc3 > gpurun_out/ncu_train2.log 2>&1; echo "rc=$?"
