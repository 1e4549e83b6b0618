#!/usr/bin/env bash
# round 2: ncu captures (one GPU for the training / evaluation kernels; the peer kernels need 2 visible GPUs in ONE process)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list of two benchmark rounds (cold-cache, serialised: compare shares)
timeout -s KILL 400 $NCU --metrics gpu__time_duration.sum -s 600 -c 160 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-tf32 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
# the training kernel (3xTF32, 8-CTA cluster), the loader and the evaluation kernel
timeout -s KILL 500 $NCU --set full --import-source on -k regex:mlp1_train_tc4 -s 2 -c 2 -o gpurun_out/ncu_tc8 -f python benchmarks/check_tc4.py ncuonly > gpurun_out/ncu_tc8.log 2>&1; echo "tc8 rc=$?"
timeout -s KILL 300 $NCU --set full --import-source on -k regex:mlp1_stage4 -s 2 -c 1 -o gpurun_out/ncu_stage4 -f python benchmarks/check_tc4.py ncuonly > gpurun_out/ncu_stage4.log 2>&1; echo "stage4 rc=$?"
timeout -s KILL 300 $NCU --set full --import-source on -k regex:mlp1_eval_tc -s 1 -c 2 -o gpurun_out/ncu_eval -f python benchmarks/micro.py eval > gpurun_out/ncu_eval.log 2>&1; echo "eval rc=$?"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout -s KILL 300 python benchmarks/peer_single_process.py > gpurun_out/peer_single_process.log 2>&1; echo "peer single rc=$?"; cat gpurun_out/peer_single_process.log | cut -c1-200
  timeout -s KILL 400 $NCU --set full --import-source on -k regex:merge_ -c 4 -o gpurun_out/ncu_peer -f python benchmarks/peer_single_process.py > gpurun_out/ncu_peer.log 2>&1; echo "ncu peer rc=$?"
fi
ls -la gpurun_out/*.ncu-rep
