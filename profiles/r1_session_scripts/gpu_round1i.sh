#!/usr/bin/env bash
# 8 GPUs, final-ish validation: cross-GPU test, bench N=2,4,8 (tc3 + tc eval + double-buffered inputs), BASELINE configs 2,3,4 on 8 GPUs
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q --timeout 400 --timeout-method=thread --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_mr.log 2>&1; echo "mr rc=$?"; tail -3 gpurun_out/pytest_mr.log | cut -c1-300
for N in 2 4 8; do
  run 400 $N $((29810+N)) bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-330; grep -v Warning gpurun_out/bench_n$N.err | grep -i "error\|Traceback" | head -3
done
for c in 2 3 4; do
  R=20; [ $c -eq 4 ] && R=10
  run 400 8 $((29830+c)) benchmarks/baseline_configs.py --config $c --rounds $R --warmup 2 > gpurun_out/cfg${c}_n8.log 2>&1; echo "cfg $c rc=$?"; grep "^{" gpurun_out/cfg${c}_n8.log | tail -1 | cut -c1-500
done
