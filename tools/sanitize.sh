#!/usr/bin/env bash
# compute-sanitizer passes over the kernel tests (run on a GPU box: gpurun [--gpus 2] -- bash tools/sanitize.sh).
# memcheck: out-of-bounds / misaligned accesses, incl. the tcgen05 / cluster training kernels (tc8 = the headline
# kernel, tc3) and the fused MERGE_UPDATE + ready/done flag protocol; racecheck: shared-memory hazards of the plain
# shared-memory kernels AND of tc8 / tc3 (racecheck does not model async-proxy / mbarrier ordering: hazards it reports
# between st.async / tcgen05 / bulk-copy traffic and generic accesses are listed, not failed on: e.g. it flags the
# evaluation kernel's cp.async.bulk refill of a stage against the lo-image readers of the stage's previous life, which
# are ordered by split[] -> tcgen05.commit -> empty[] -> producer, a chain of mbarriers it does not follow).
# With two GPUs the cross-GPU handshake (2 ranks, C++ executor + Python executor) runs under memcheck as well.
set -u
mkdir -p gpurun_out
PY="python -m pytest -m gpu -q -x -p no:cacheprovider"
K='merge or optim or segment or logreg or sequential or keyed or eval'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 $PY tests/test_kernels_gpu.py -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 $PY tests/test_kernels_gpu.py -k "fp32_equivalent or first_step or fused_merge_update or handshake or determin or partition_scaled or (tf32_kernels and 96)" > gpurun_out/sanitize_memcheck_train.log 2>&1; echo "memcheck(training kernels + flag protocol) rc=$?"
tail -3 gpurun_out/sanitize_memcheck_train.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 $PY tests/test_kernels_gpu.py -k "logreg or sequential or (eval_confusion and simt)" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool racecheck $PY tests/test_kernels_gpu.py -k "determin or (fp32_equivalent and 96) or (eval_confusion and tc)" > gpurun_out/sanitize_racecheck_train.log 2>&1; echo "racecheck(tc8, tc3) rc=$? (report only)"
grep -c "Race reported\|hazard" gpurun_out/sanitize_racecheck_train.log; tail -3 gpurun_out/sanitize_racecheck_train.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 1200 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 $PY tests/test_multirank.py -k "cpp_executor_two_ranks_cuda" > gpurun_out/sanitize_memcheck_2gpu.log 2>&1; echo "memcheck(2 GPUs, flag protocol) rc=$?"
  grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_memcheck_2gpu.log; tail -3 gpurun_out/sanitize_memcheck_2gpu.log
fi
