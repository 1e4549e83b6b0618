#!/usr/bin/env bash
# Second GPU contact (2 GPUs): all gpu tests incl. cross-GPU, bench N=1/N=2, ncu captures.
set -u
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
echo "== bench N=1 (auto/tc) =="
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.json | cut -c1-1500; tail -3 gpurun_out/bench_n1.err
echo "== bench N=2 =="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-1500; tail -5 gpurun_out/bench_n2.err
echo "== ncu: launch list of 2 rounds =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu: full capture of the tc training kernel =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp1_train_tc -s 3 -c 1 -f -o gpurun_out/prof_train_tc python benchmarks/micro.py train --impl tc > gpurun_out/ncu_train.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_train.log
ls -la gpurun_out
