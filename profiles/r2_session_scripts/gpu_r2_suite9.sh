#!/usr/bin/env bash
# round 2, one GPU: asynchronous all-to-all (config 7) through the C++ executor vs per event, GPU suite with the new executor modes
set -u
mkdir -p gpurun_out
: > gpurun_out/config7.jsonl
timeout -s KILL 300 python benchmarks/baseline_configs.py --config 7 --rounds 30 --warmup 5 2> gpurun_out/cfg7.err | grep "^{" >> gpurun_out/config7.jsonl; echo "config 7 (executor) rc=$?"
timeout -s KILL 300 python benchmarks/baseline_configs.py --config 7 --rounds 30 --warmup 5 --no-executor 2> gpurun_out/cfg7_ev.err | grep "^{" >> gpurun_out/config7.jsonl; echo "config 7 (per event) rc=$?"
cut -c1-330 gpurun_out/config7.jsonl
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-220
