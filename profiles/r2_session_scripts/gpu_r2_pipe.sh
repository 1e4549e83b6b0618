#!/usr/bin/env bash
# round 2, one GPU: forward MMAs pipelined with the W += G pass -- accuracy checks, kernel tests, timing, headline
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python benchmarks/check_tc4.py tc8 > gpurun_out/check_tc4_pipe.log 2>&1; echo "check_tc4 rc=$?"; grep "update vs fp64\|timing" gpurun_out/check_tc4_pipe.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if d['check'].startswith('update'): print('steps', d['steps'], 'kernel_err', d['kernel_max_err'], 'torch_err', d['torch_fp32_max_err'])
    else: print({k: d[k] for k in d if k in ('impl', 'ms_per_update', 'us_per_step_marginal', 'us_per_step', 'fixed_us_per_launch')})
"
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "mlp1 or flagship or momentum or determin or fused_merge or partition or handshake or tf32" -p no:cacheprovider > gpurun_out/pytest_pipe.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/pytest_pipe.log | cut -c1-300
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_pipe.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'tf32', d['tf32']['value'], d['test_acc_by_round_tail'][-2:])"
