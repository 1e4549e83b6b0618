#!/usr/bin/env bash
# round 2, one GPU: the three node classes of main_giaretta_2019 at 4 141 nodes (banked engine vs per-event execution), GPU suite
set -u
mkdir -p gpurun_out
: > gpurun_out/many_nodes_variants.jsonl
for node in gossip passthrough cacheneigh; do
  timeout -s KILL 300 python benchmarks/many_nodes.py --node $node --impl banked 2> gpurun_out/mn_$node.err | grep "^{" >> gpurun_out/many_nodes_variants.jsonl; echo "banked $node rc=$?"
done
for node in passthrough cacheneigh; do
  timeout -s KILL 300 python benchmarks/many_nodes.py --node $node --impl events --rounds 5 2> gpurun_out/mn_ev_$node.err | grep "^{" >> gpurun_out/many_nodes_variants.jsonl; echo "events $node rc=$?"
done
cut -c1-330 gpurun_out/many_nodes_variants.jsonl
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
