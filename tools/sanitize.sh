#!/usr/bin/env bash
# compute-sanitizer passes over the kernel tests (run on a GPU box: gpurun -- bash tools/sanitize.sh).
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the cluster / DSMEM and
# tcgen05 kernels are excluded from racecheck -- it does not model async-proxy / mbarrier ordering).
set -u
mkdir -p gpurun_out
K='merge or optim or segment or logreg or sequential or keyed or eval'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "tc2 and 96" > gpurun_out/sanitize_memcheck_tc2.log 2>&1; echo "memcheck(tc2) rc=$?"
tail -5 gpurun_out/sanitize_memcheck_tc2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "logreg or sequential or eval_confusion" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/sanitize_racecheck.log
