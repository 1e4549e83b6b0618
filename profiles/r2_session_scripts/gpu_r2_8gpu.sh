#!/usr/bin/env bash
# round 2, N GPUs (default 8): headline at N, BASELINE configs 3 / 4 / weak scaling, NCCL harness, peer merge vs NCCL
set -u
N=${NGPU:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 240 29881 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
tail -1 gpurun_out/bench_n${N}.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e'] and d['e2e']['value'], 'tf32', d.get('tf32') and d['tf32']['value'], d['details']['executor'], d['test_acc_by_round_tail'])
except Exception as e:
    print('bad json', e)
"; grep -i "error\|Traceback" gpurun_out/bench_n${N}.err | head -3
if [ "${PY_TRANSPORTS:-0}" = 1 ]; then
for tr in p2p nccl; do
  run 300 29871 bench.py --gpus $N --steps 30 --warmup 3 --executor python --transport $tr --no-e2e --no-tf32 > gpurun_out/bench_n${N}_py_$tr.json 2> gpurun_out/bench_n${N}_py_$tr.err; echo "bench N=$N python executor, transport $tr rc=$?"
  tail -1 gpurun_out/bench_n${N}_py_$tr.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step')}, d['details']['transport'], d['test_acc_by_round_tail'][-2:])
except Exception as e:
    print('bad json', e)
"
done
fi
: > gpurun_out/baseline_configs_n$N.jsonl
p=29890
for c in 3 4 6; do
  p=$((p+1)); run 240 $p benchmarks/baseline_configs.py --config $c 2> gpurun_out/cfg${c}_n$N.err | grep "^{" >> gpurun_out/baseline_configs_n$N.jsonl; echo "config $c N=$N rc=$?"
done
cut -c1-330 gpurun_out/baseline_configs_n$N.jsonl
run 300 29885 baseline/nccl_harness.py --steps 5 --warmup 3 > gpurun_out/nccl_harness_n$N.json 2> gpurun_out/nccl_harness_n$N.err; echo "harness N=$N rc=$?"; grep "^{" gpurun_out/nccl_harness_n$N.json | cut -c1-300
run 300 29887 baseline/nccl_harness.py --steps 5 --warmup 3 --all2all > gpurun_out/nccl_harness_a2a_n$N.json 2> gpurun_out/nccl_harness_a2a_n$N.err; echo "harness a2a N=$N rc=$?"; grep "^{" gpurun_out/nccl_harness_a2a_n$N.json | cut -c1-300
run 300 29883 benchmarks/peer_merge.py > gpurun_out/peer_merge_w$N.log 2>&1; echo "peer_merge rc=$?"; grep "^{" gpurun_out/peer_merge_w$N.log | cut -c1-230
