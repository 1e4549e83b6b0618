#!/usr/bin/env bash
# round 2: bring-up of the fp32-equivalent training kernel (accuracy vs fp64, timing, phase counters)
set -u
mkdir -p gpurun_out
timeout -s KILL ${T4_TIMEOUT:-240} python benchmarks/check_tc4.py "$@" > gpurun_out/check_tc4.log 2>&1; echo "check rc=$?"
cut -c1-1200 gpurun_out/check_tc4.log | tail -60
