#!/usr/bin/env bash
# round 2, one GPU: channels-last rows for conv models (tests, ResNet-20 step timing, config 5 with / without)
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cuda_graph or channels_last" -p no:cacheprovider > gpurun_out/pytest_cl.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/pytest_cl.log | cut -c1-300; grep -h "Error" gpurun_out/pytest_cl.log | head -5 | cut -c1-300
for cl in 0 1; do
  GOSSIPY_CHANNELS_LAST=$cl timeout -s KILL 400 python benchmarks/check_graph_step.py 2> gpurun_out/check_graph_step_cl$cl.err | grep resnet | sed "s/^{/{\"channels_last\": $cl, /" > gpurun_out/check_graph_step_cl$cl.jsonl; echo "check_graph_step cl=$cl rc=$?"; cat gpurun_out/check_graph_step_cl$cl.jsonl
done
: > gpurun_out/config5_cl.jsonl
for cl in 0 1; do
  GOSSIPY_CHANNELS_LAST=$cl timeout -s KILL 500 python benchmarks/baseline_configs.py --config 5 --rounds 20 --warmup 8 2> gpurun_out/cfg5_cl$cl.err | grep "^{" | sed "s/^{/{\"channels_last\": $cl, /" >> gpurun_out/config5_cl.jsonl; echo "config 5 channels_last=$cl rc=$?"
done
cut -c1-260 gpurun_out/config5_cl.jsonl
GOSSIPY_CHANNELS_LAST=1 timeout -s KILL 400 python benchmarks/profile_config5.py > gpurun_out/profile_config5_cl.txt 2> gpurun_out/profile_config5_cl.err; echo "profile rc=$?"; head -14 gpurun_out/profile_config5_cl.txt | cut -c1-180
