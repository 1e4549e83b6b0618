#!/usr/bin/env bash
# round 2, one GPU: graph-replay tests repeated in ONE process (pool reuse after earlier handlers died), full GPU suite, smoke
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "cuda_graph or keyed_perm" -p no:cacheprovider > gpurun_out/pytest_graph.log 2>&1; echo "graph tests rc=$?"; tail -3 gpurun_out/pytest_graph.log | cut -c1-300; grep -h "Error" gpurun_out/pytest_graph.log | head -5 | cut -c1-400
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 400 python benchmarks/check_graph_step.py > gpurun_out/check_graph_step.jsonl 2> gpurun_out/check_graph_step.err; echo "check_graph_step rc=$?"; tail -2 gpurun_out/check_graph_step.jsonl; grep -i "warn\|fail" gpurun_out/check_graph_step.err | head -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
