#!/usr/bin/env bash
# round 2, 2 GPUs in ONE process: the peer kernels (pair merge pulling cuda:1's row, k-way merge) timed and captured by ncu
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout -s KILL 300 python benchmarks/peer_single_process.py > gpurun_out/peer_single_process.log 2>&1; echo "peer single rc=$?"; grep "^{" gpurun_out/peer_single_process.log | cut -c1-200
PEER_FLOATS=$((1<<24)) timeout -s KILL 500 $NCU --set full --import-source on -k regex:merge_ -c 6 -o gpurun_out/ncu_peer -f python benchmarks/peer_single_process.py > gpurun_out/ncu_peer.log 2>&1; echo "ncu peer rc=$?"; tail -3 gpurun_out/ncu_peer.log
ls -la gpurun_out/ncu_peer.ncu-rep
