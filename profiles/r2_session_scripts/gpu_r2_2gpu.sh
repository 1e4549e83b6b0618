#!/usr/bin/env bash
# round 2, N GPUs (default 2): multi-rank correctness on the GPUs, headline at N, the NCCL(+cuBLAS) harness, peer-merge vs NCCL
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout -s KILL 600 python -m pytest tests/test_multirank.py -m gpu -q --timeout 400 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr_n$N.log 2>&1; echo "multirank tests rc=$?"; tail -5 gpurun_out/pytest_mr_n$N.log | cut -c1-300
fi
for ex in native python; do
  run 300 29881 bench.py --gpus $N --steps 40 --warmup 3 --executor $ex --no-tf32 > gpurun_out/bench_n${N}_$ex.json 2> gpurun_out/bench_n${N}_$ex.err; echo "bench N=$N $ex rc=$?"
  tail -1 gpurun_out/bench_n${N}_$ex.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e'] and d['e2e']['value'], d['details']['executor'], d['test_acc_by_round_tail'])
except Exception as e:
    print('bad json', e)
"; grep -i "error\|Traceback" gpurun_out/bench_n${N}_$ex.err | head -3
done
timeout -s KILL 300 python baseline/nccl_harness.py --steps 5 --warmup 3 > gpurun_out/nccl_harness_n1.json 2> gpurun_out/nccl_harness_n1.err; echo "harness N=1 rc=$?"; cut -c1-400 gpurun_out/nccl_harness_n1.json; tail -3 gpurun_out/nccl_harness_n1.err
run 300 29885 baseline/nccl_harness.py --steps 5 --warmup 3 > gpurun_out/nccl_harness_n$N.json 2> gpurun_out/nccl_harness_n$N.err; echo "harness N=$N rc=$?"; grep "^{" gpurun_out/nccl_harness_n$N.json | cut -c1-400; tail -3 gpurun_out/nccl_harness_n$N.err
run 300 29887 baseline/nccl_harness.py --steps 5 --warmup 3 --all2all > gpurun_out/nccl_harness_a2a_n$N.json 2> gpurun_out/nccl_harness_a2a_n$N.err; echo "harness a2a N=$N rc=$?"; grep "^{" gpurun_out/nccl_harness_a2a_n$N.json | cut -c1-300
run 400 29883 benchmarks/peer_merge.py > gpurun_out/peer_merge_w$N.log 2>&1; echo "peer_merge rc=$?"; grep "^{" gpurun_out/peer_merge_w$N.log | cut -c1-230
