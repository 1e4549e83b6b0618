#!/usr/bin/env bash
# one ncu --set full capture of the current tc3 training kernel + its loader (1 GPU)
set -u
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"mlp1_train_tc3|mlp1_stage" -s 6 -c 2 -f -o gpurun_out/prof_train_tc3_r1l python benchmarks/micro.py train --impl tc3 > gpurun_out/ncu_train_r1l.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_train_r1l.log | cut -c1-200; ls -la gpurun_out/prof_train_tc3_r1l.ncu-rep
