#!/usr/bin/env bash
# round 2, one GPU: graph-step diagnostics, GPU suite, headline bench (request-leg elision A/B), config 5 with device-side sample order
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python benchmarks/check_graph_step.py > gpurun_out/check_graph_step.jsonl 2> gpurun_out/check_graph_step.err; echo "check_graph_step rc=$?"; cat gpurun_out/check_graph_step.jsonl; tail -3 gpurun_out/check_graph_step.err
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
GOSSIPY_EXEC_ELIDE=0 timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-tf32 > gpurun_out/bench_n1_noelide.json 2> gpurun_out/bench_n1_noelide.err; echo "bench (no elision) rc=$?"
for f in gpurun_out/bench_n1.json gpurun_out/bench_n1_noelide.json; do tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$f', {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d.get('e2e') and d['e2e']['value'], 'tf32', d.get('tf32') and d['tf32']['value'], d['test_acc_by_round_tail'][-2:])
except Exception as e:
    print('bad json', e)
"; done
: > gpurun_out/config5.jsonl
for gflag in 0 1; do
  GOSSIPY_CUDA_GRAPHS=$gflag timeout -s KILL 400 python benchmarks/baseline_configs.py --config 5 --rounds 5 --warmup 2 2> gpurun_out/cfg5_g$gflag.err | grep "^{" | sed "s/^{/{\"cuda_graphs\": $gflag, /" >> gpurun_out/config5.jsonl; echo "config 5 graphs=$gflag rc=$?"
done
cut -c1-300 gpurun_out/config5.jsonl
