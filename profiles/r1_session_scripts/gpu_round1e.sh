#!/usr/bin/env bash
# 8 GPUs: cross-GPU tests, scaling bench N=2,4,8, peer-merge bandwidth, NVLS all-reduce, all2all example
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
echo "== multirank test (2 GPUs) =="
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q --timeout 400 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_mr.log | cut -c1-400
echo "== peer merge (2 ranks) =="
run 300 2 29701 benchmarks/peer_merge.py > gpurun_out/peer_merge_w2.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/peer_merge_w2.log | cut -c1-300; grep -v "^{" gpurun_out/peer_merge_w2.log | tail -5 | cut -c1-300
echo "== peer merge (8 ranks) =="
run 300 8 29702 benchmarks/peer_merge.py > gpurun_out/peer_merge_w8.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/peer_merge_w8.log | cut -c1-300; grep -v "^{" gpurun_out/peer_merge_w8.log | tail -5 | cut -c1-300
for N in 2 4 8; do
  echo "== bench N=$N =="
  run 400 $N $((29710+N)) bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-700; grep -v Warning gpurun_out/bench_n$N.err | tail -3 | cut -c1-300
done
echo "== all2all synchronous (8 ranks, NVLS) =="
GOSSIPY_SYNC=1 GOSSIPY_ROUNDS=20 run 300 8 29730 examples/main_all2all.py > gpurun_out/all2all_sync_w8.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/all2all_sync_w8.log | cut -c1-300
