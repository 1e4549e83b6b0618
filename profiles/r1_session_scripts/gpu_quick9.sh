#!/usr/bin/env bash
# programmatic dependent launch of the tc3 training kernel behind its loader: correctness + A/B timing (1 GPU)
set -u
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 --timeout-method=thread --tb=short -rf -p no:cacheprovider -k "tc3 or train or stage or fused or handshake" > gpurun_out/pytest_q9.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_q9.log | cut -c1-300
for v in 0 1; do
  GB_TC3_PDL=$v timeout -s KILL 120 python benchmarks/micro.py train --impl tc3 > gpurun_out/micro_q9_$v.log 2>&1; echo "micro pdl=$v rc=$?"; grep "^{" gpurun_out/micro_q9_$v.log | grep "mlp1_train\"\|fixed cost" | cut -c1-400
done
timeout -s KILL 200 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_q9.json 2> gpurun_out/bench_q9.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_q9.json | cut -c1-330
