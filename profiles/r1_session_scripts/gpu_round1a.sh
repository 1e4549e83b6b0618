#!/usr/bin/env bash
# First GPU contact: tests, micro-benchmarks, headline bench (both arms), launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest gpu (cluster impl is the safe default via env) =="
timeout 600 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
echo "== micro (cluster) =="
timeout 300 python benchmarks/micro.py all --impl cluster > gpurun_out/micro_cluster.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/micro_cluster.log
echo "== bench native (cluster) =="
timeout 600 python bench.py --steps 10 --warmup 3 --train-impl cluster > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; echo "rc=$?"; tail -3 gpurun_out/bench_cluster.json; tail -5 gpurun_out/bench_cluster.err
echo "== bench reference =="
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -2 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
echo "== tc probe / debug =="
timeout 120 python benchmarks/probe_tc.py > gpurun_out/probe_tc.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/probe_tc.log
timeout 120 python benchmarks/debug_tc.py > gpurun_out/debug_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/debug_tc.log
echo "== micro (tc) =="
timeout 200 python benchmarks/micro.py train --impl tc > gpurun_out/micro_tc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/micro_tc.log
