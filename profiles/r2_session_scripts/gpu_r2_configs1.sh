#!/usr/bin/env bash
# round 2, one GPU: the BASELINE.json configurations, the weak-scaling base point, accuracy bands vs the reference
set -u
mkdir -p gpurun_out
: > gpurun_out/baseline_configs_n1.jsonl
for c in 1 2 3 4 5 6; do
  extra=""; [ $c = 5 ] && extra="--rounds 3 --warmup 1"
  timeout -s KILL 400 python benchmarks/baseline_configs.py --config $c $extra 2> gpurun_out/cfg$c.err | grep "^{" >> gpurun_out/baseline_configs_n1.jsonl; echo "config $c rc=$?"
done
cut -c1-330 gpurun_out/baseline_configs_n1.jsonl
timeout -s KILL 300 python benchmarks/many_nodes.py > gpurun_out/many_nodes.log 2>&1; echo "many_nodes rc=$?"; grep "^{" gpurun_out/many_nodes.log | cut -c1-300
timeout -s KILL 900 python benchmarks/acc_band.py --seeds 5 --rounds 50 > gpurun_out/acc_band.json 2> gpurun_out/acc_band.err; echo "acc_band rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/acc_band.json'))
    print({k: d[k] for k in ('inside_band', 'max_gap', 'max_gap_over_tolerance', 'rounds', 'seeds')})
    print('ours', d['ours_mean'][::7]); print('ref ', d['ref_mean'][::7])
except Exception as e:
    print('acc_band: bad output', e)
PY
