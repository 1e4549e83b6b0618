#!/usr/bin/env bash
# round 2, 2 GPUs: multi-rank checkpoint / resume of the C++ executor on the GPUs
set -u
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_multirank.py -m gpu -q -x -k "checkpoint_with_two_ranks_cuda" -p no:cacheprovider > gpurun_out/pytest_mr_ckpt.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_mr_ckpt.log | cut -c1-300
