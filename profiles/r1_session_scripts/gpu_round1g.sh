#!/usr/bin/env bash
# 1 GPU: smoke(), BASELINE configs 1-5 at N=1, ncu capture of tc3, compute-sanitizer passes
set -u
mkdir -p gpurun_out
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
echo "== baseline configs N=1 =="
for c in 1 2 3 4 5; do
  R=10; [ $c -eq 5 ] && R=3; [ $c -eq 4 ] && R=5
  timeout 600 python benchmarks/baseline_configs.py --config $c --rounds $R --warmup 2 > gpurun_out/cfg${c}_n1.log 2>&1; echo "cfg $c rc=$?"; tail -1 gpurun_out/cfg${c}_n1.log | cut -c1-600
done
echo "== ncu full tc3 =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp1_train_tc3 -s 3 -c 1 -f -o gpurun_out/prof_train_tc3 python benchmarks/micro.py train --impl tc3 > gpurun_out/ncu_train3.log 2>&1; echo "rc=$?"
echo "== sanitizer =="
bash tools/sanitize.sh
