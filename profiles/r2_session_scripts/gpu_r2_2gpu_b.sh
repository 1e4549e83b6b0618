#!/usr/bin/env bash
# round 2, 2 GPUs, second pass: multi-rank GPU tests (bank, nccl transport, partitioned / UPDATE_MERGE executor, arena growth),
# headline at N=2 (native executor; Python executor over p2p and over the NCCL transport), two-shot all-reduce, peer-kernel ncu, sanitizer
set -u
N=2
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout -s KILL 900 python -m pytest tests/test_multirank.py -m gpu -q --timeout 600 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr_n$N.log 2>&1; echo "multirank tests rc=$?"; tail -8 gpurun_out/pytest_mr_n$N.log | cut -c1-300
run 300 29881 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
for tr in p2p nccl; do
  run 300 29871 bench.py --gpus $N --steps 30 --warmup 3 --executor python --transport $tr --no-e2e --no-tf32 > gpurun_out/bench_n${N}_py_$tr.json 2> gpurun_out/bench_n${N}_py_$tr.err; echo "bench N=$N python executor, transport $tr rc=$?"
done
for f in gpurun_out/bench_n${N}.json gpurun_out/bench_n${N}_py_p2p.json gpurun_out/bench_n${N}_py_nccl.json; do tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$f', {k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d.get('e2e') and d['e2e']['value'], d['details'].get('transport'), d['test_acc_by_round_tail'][-2:])
except Exception as e:
    print('bad json', e)
"; done
run 400 29883 benchmarks/peer_merge.py > gpurun_out/peer_merge_w$N.log 2>&1; echo "peer_merge rc=$?"; grep "^{" gpurun_out/peer_merge_w$N.log | grep -i "allreduce\|baseline: nccl" | cut -c1-230
NCU="ncu --clock-control none"
timeout -s KILL 300 python benchmarks/peer_single_process.py > gpurun_out/peer_single_process.log 2>&1; echo "peer single rc=$?"; cut -c1-200 gpurun_out/peer_single_process.log | tail -8
timeout -s KILL 400 $NCU --set full --import-source on -k regex:merge_ -c 4 -o gpurun_out/ncu_peer -f python benchmarks/peer_single_process.py > gpurun_out/ncu_peer.log 2>&1; echo "ncu peer rc=$?"; tail -3 gpurun_out/ncu_peer.log
timeout -s KILL 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_multirank.py -k "cpp_executor_two_ranks_cuda" > gpurun_out/sanitize_memcheck_2gpu.log 2>&1; echo "memcheck(2 GPUs, flag protocol) rc=$?"
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_memcheck_2gpu.log; grep "ERROR SUMMARY" gpurun_out/sanitize_memcheck_2gpu.log | sort | uniq -c | head -5; tail -3 gpurun_out/sanitize_memcheck_2gpu.log
ls -la gpurun_out/*.ncu-rep
