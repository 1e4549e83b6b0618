#!/usr/bin/env bash
# SECOND GPU call of the next round (--gpus 2): the multi-rank launch path of the C++ executor (flag waits / signals,
# PeerSync into the kernels) has only run on CPU so far; NCCL-only baselines of the peer merge; weak scaling A/B.
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
timeout -s KILL 500 python -m pytest tests/test_multirank.py -m gpu -q --timeout 300 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr_next.log 2>&1; echo "mr rc=$?"; tail -4 gpurun_out/pytest_mr_next.log | cut -c1-300
for ex in python native; do
  run 300 29881 bench.py --gpus 2 --steps 50 --warmup 3 --executor $ex > gpurun_out/bench_next_n2_$ex.json 2> gpurun_out/bench_next_n2_$ex.err; echo "bench N=2 $ex rc=$?"; tail -1 gpurun_out/bench_next_n2_$ex.json | cut -c1-200
done
run 300 29883 benchmarks/peer_merge.py > gpurun_out/peer_merge_next_w2.log 2>&1; echo "peer_merge rc=$?"; grep "^{" gpurun_out/peer_merge_next_w2.log | cut -c1-220
