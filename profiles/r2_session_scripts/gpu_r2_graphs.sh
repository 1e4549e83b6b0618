#!/usr/bin/env bash
# round 2, one GPU: CUDA-graph replay of the generic step (tests + config 5 with / without), accuracy band on identical shards
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cuda_graph" -p no:cacheprovider > gpurun_out/pytest_graph.log 2>&1; echo "graph tests rc=$?"; tail -5 gpurun_out/pytest_graph.log
: > gpurun_out/config5_graphs.jsonl
for gflag in 0 1; do
  GOSSIPY_CUDA_GRAPHS=$gflag timeout -s KILL 400 python benchmarks/baseline_configs.py --config 5 --rounds 5 --warmup 2 2> gpurun_out/cfg5_g$gflag.err | grep "^{" | sed "s/^{/{\"cuda_graphs\": $gflag, /" >> gpurun_out/config5_graphs.jsonl; echo "config 5 graphs=$gflag rc=$?"
done
cut -c1-420 gpurun_out/config5_graphs.jsonl
tail -3 gpurun_out/cfg5_g1.err
timeout -s KILL 300 python -m pytest tests/test_acc_band_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_accband.log 2>&1; echo "acc band test rc=$?"; tail -3 gpurun_out/pytest_accband.log
timeout -s KILL 900 python benchmarks/acc_band.py --seeds 5 --rounds 50 2> gpurun_out/acc_band.err | grep "^{" > gpurun_out/acc_band.json; echo "acc_band rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/acc_band.json'))
    print({k: d[k] for k in ('inside_band', 'max_gap', 'max_gap_over_tolerance', 'rounds', 'seeds')})
    print('ours', d['ours_mean'][::7]); print('ref ', d['ref_mean'][::7])
except Exception as e:
    print('acc_band: bad output', e)
PY
