"""One process per rank (torch.distributed): every rank replays the host-side simulation and executes
only the device work of the nodes it owns; models cross ranks through shared arenas with the
ready/done flag handshake.  The multi-rank result must equal the single-process result.

CPU (gloo, POSIX shared memory) runs everywhere; the CUDA variant (NCCL bootstrap, CUDA-IPC arenas,
peer loads over NVLink inside the merge / training kernels) needs >= 2 GPUs."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_mr_worker.py")
KINDS = "pegasos,mlp_pushpull,limited_pull,partitioned,all2all,all2all_sync"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, device, rounds=3, kinds=KINDS, transport=None, arena_rows=None, checkpoint=False, metrics_every=None,
         placement=None):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    if placement:
        env["MR_PLACEMENT"] = placement
    if metrics_every:
        env["MR_METRICS_EVERY"] = str(metrics_every)
    if checkpoint:
        env["MR_CHECKPOINT"] = "1"
    if transport:
        env["GOSSIPY_B200_TRANSPORT"] = transport
    if arena_rows:
        env["GOSSIPY_B200_ARENA_ROWS"] = str(arena_rows)
    if world == 1:
        cmd = [sys.executable, WORKER, kinds, device, str(rounds)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER, kinds, device,
               str(rounds)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, "worker failed:\n" + out.stdout[-2000:] + out.stderr[-4000:]
    return json.loads(lines[-1][len("RESULT "):])


def _compare(a, b, rel, skip=()):
    for kind in a:
        x, y = a[kind], b[kind]
        for f in ("sent", "failed", "size", "ages", "cache_left", "best"):
            if f in skip:
                continue
            assert x[f] == y[f], (kind, f, x[f], y[f])
        for f in ("glob", "loc"):
            assert len(x[f]) == len(y[f])
            for (t1, m1), (t2, m2) in zip(x[f], y[f]):
                assert t1 == t2 and m1.keys() == m2.keys()
                for k in m1:
                    assert m1[k] == pytest.approx(m2[k], abs=2e-2), (kind, f, k)
        assert x["sums"].keys() == y["sums"].keys()
        for n in x["sums"]:
            assert x["sums"][n] == pytest.approx(y["sums"][n], rel=rel, abs=rel), (kind, n)


def test_two_ranks_cpu_equal_single_process():
    single = _run(1, "cpu")
    multi = _run(2, "cpu")
    _compare(single, multi, rel=1e-5)


def test_nccl_transport_two_and_three_ranks_cpu_equal_single_process():
    """``transport="nccl"`` (send / recv of the row + the same kernels on a staging copy: the NCCL-only baseline of the
    engine, here over gloo) gives the results of the shared-arena transport and of a single process."""
    kinds = "mlp_pushpull,limited_pull,partitioned"
    single = _run(1, "cpu", kinds=kinds)
    for world in (2, 3):
        _compare(single, _run(world, "cpu", kinds=kinds, transport="nccl"), rel=1e-5)


def test_loopback_transport_runs_everything_in_every_process():
    """``transport="loopback"``: no inter-rank transport, every process of the job hosts all nodes itself."""
    kinds = "mlp_pushpull,x_mlp_pushpull"
    _compare(_run(1, "cpu", kinds=kinds), _run(2, "cpu", kinds=kinds, transport="loopback"), rel=1e-6)


def test_pens_two_and_three_ranks_cpu_equal_single_process():
    """PENS: the top-m choice is made on the owner from device results and broadcast; both steps run."""
    single = _run(1, "cpu", rounds=9, kinds="pens")
    assert any(v for v in single["pens"]["best"].values()), "step 2 was never reached"
    # (multi-rank nodes hold their step-1 candidates as local scratch rows instead of CACHE entries)
    _compare(single, _run(2, "cpu", rounds=9, kinds="pens"), rel=1e-5, skip=("cache_left",))
    _compare(single, _run(3, "cpu", rounds=9, kinds="pens"), rel=1e-5, skip=("cache_left",))


XKINDS = "x_mlp_pushpull,x_limited_push,x_update_pull,x_update_merge,x_passthrough,x_sampled,x_cacheneigh,x_momentum,x_all2all,x_part_mlp,x_part_logreg"


def test_cpp_executor_two_and_three_ranks_cpu_equal_single_process():
    """csrc/exec with several ranks: replicated books, each rank launches its own nodes, snapshot slots in the
    symmetric arenas with the ready/done handshake (here: shared-memory flags, host-side waits)."""
    kinds = XKINDS + ",x_part_update,x_sampled_update"    # (UPDATE of partitioned / sampled models: CPU only so far)
    single = _run(1, "cpu", rounds=4, kinds=kinds)
    assert all(v["cpp_executor"] for v in single.values())
    for world in (2, 3):
        multi = _run(world, "cpu", rounds=4, kinds=kinds)
        assert all(v["cpp_executor"] for v in multi.values())
        _compare(single, multi, rel=1e-5)


def test_checkpoint_with_several_ranks_resumes_exactly():
    """save / load with two ranks: every rank writes a complete checkpoint (the owners' in-flight snapshot slots and cached
    models are gathered), only owners restore row values, a barrier separates restoring from reading; the interrupted run
    equals the uninterrupted single-process run -- Python executor and C++ executor (delays, caches, partitioned models)."""
    kinds = "mlp_pushpull,limited_pull,x_update_pull,x_limited_push,x_all2all,x_cacheneigh,x_part_logreg,x_part_update,x_sampled_update"
    single = _run(1, "cpu", rounds=6, kinds=kinds)
    _compare(single, _run(1, "cpu", rounds=6, kinds=kinds, checkpoint=True), rel=1e-5)
    _compare(single, _run(2, "cpu", rounds=6, kinds=kinds, checkpoint=True), rel=1e-5)


def test_explicit_placements_give_the_single_process_result():
    """``runtime.Placement``: round-robin and load-balanced node -> rank maps (installed before ``init_nodes``, which keeps
    them) instead of the default blocks; Python executor, C++ executor and the bank."""
    kinds = "mlp_pushpull,x_mlp_pushpull,x_cacheneigh,bank_pegasos"
    single = _run(1, "cpu", rounds=4, kinds=kinds)
    _compare(single, _run(2, "cpu", rounds=4, kinds=kinds, placement="round_robin"), rel=1e-5)
    _compare(single, _run(3, "cpu", rounds=4, kinds=kinds, placement="by_load"), rel=1e-5)


def test_metrics_exchanged_every_k_rounds_give_the_same_report():
    """``metrics_sync_every = k``: one all-reduce of the evaluation results per k rounds (and at the end / before a
    checkpoint) -- the report is the one of the per-round exchange."""
    kinds = "mlp_pushpull,x_mlp_pushpull,pegasos"
    single = _run(1, "cpu", rounds=7, kinds=kinds)
    _compare(single, _run(2, "cpu", rounds=7, kinds=kinds, metrics_every=3), rel=1e-5)
    _compare(single, _run(2, "cpu", rounds=7, kinds=kinds, metrics_every=3, checkpoint=True), rel=1e-5)


def test_symmetric_arenas_grow_when_they_run_out():
    """Segments of 4 rows: the shared arenas (and the C++ executor's snapshot pools on them) must add segments
    collectively, in the middle of a round, without changing the results."""
    kinds = "mlp_pushpull,x_mlp_pushpull"
    single = _run(1, "cpu", kinds=kinds)
    _compare(single, _run(2, "cpu", kinds=kinds, arena_rows=4), rel=1e-5)


BKINDS = "bank_pegasos,bank_adaline_pushpull,bank_passthrough,bank_cacheneigh"


def test_banked_engine_two_and_three_ranks_cpu_equal_single_process():
    """engine/bank.py with several ranks: snapshots pushed into the slot bank of the receiver's rank, replicated slot
    pools, a barrier after every phase with cross-rank pushes."""
    single = _run(1, "cpu", rounds=5, kinds=BKINDS)
    assert all(v["banked"] for v in single.values())
    for world in (2, 3):
        multi = _run(world, "cpu", rounds=5, kinds=BKINDS)
        assert all(v["banked"] for v in multi.values())
        _compare(single, multi, rel=1e-5)


def test_generic_conv_model_two_ranks_cpu_equal_single_process():
    single = _run(1, "cpu", rounds=4, kinds="cnn_pushpull")
    _compare(single, _run(2, "cpu", rounds=4, kinds="cnn_pushpull"), rel=1e-4)


def test_banked_checkpoint_with_several_ranks_resumes_exactly():
    """save / load of the banked engine with two and three ranks: snapshots on the wire (and cached models) live in the bank
    of their receiver's rank, are gathered into every rank's checkpoint and restored by their owners only."""
    single = _run(1, "cpu", rounds=6, kinds=BKINDS)
    _compare(single, _run(1, "cpu", rounds=6, kinds=BKINDS, checkpoint=True), rel=1e-5)
    for world in (2, 3):
        multi = _run(world, "cpu", rounds=6, kinds=BKINDS, checkpoint=True)
        assert all(v["banked"] for v in multi.values())
        _compare(single, multi, rel=1e-5)


@pytest.mark.gpu
def test_checkpoint_with_two_ranks_cuda_resumes_exactly():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    kinds = "x_update_pull,x_all2all"
    single = _run(1, "cuda:0", rounds=6, kinds=kinds)
    _compare(single, _run(2, "cuda", rounds=6, kinds=kinds, checkpoint=True), rel=2e-3)


@pytest.mark.gpu
def test_generic_conv_model_two_ranks_cuda_equal_single_gpu():
    """Generic models across GPUs (eager steps: graph replay and channels-last rows are single-rank features)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = _run(1, "cuda:0", rounds=5, kinds="cnn_pushpull")
    _compare(single, _run(2, "cuda", rounds=5, kinds="cnn_pushpull"), rel=5e-2)


@pytest.mark.gpu
def test_banked_engine_two_ranks_cuda_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = _run(1, "cuda:0", rounds=5, kinds=BKINDS)
    multi = _run(2, "cuda", rounds=5, kinds=BKINDS)
    assert all(v["banked"] for v in multi.values())
    _compare(single, multi, rel=1e-4)


@pytest.mark.gpu
def test_two_ranks_cuda_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = _run(1, "cuda:0")
    multi = _run(2, "cuda")
    _compare(single, multi, rel=2e-3)


@pytest.mark.gpu
def test_nccl_transport_two_ranks_cuda_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    kinds = "mlp_pushpull,limited_pull,partitioned"
    _compare(_run(1, "cuda:0", kinds=kinds), _run(2, "cuda", kinds=kinds, transport="nccl"), rel=2e-3)


def test_pens_native_scheduler_two_ranks_cpu_equal_single_process():
    """PENS under the C++ scheduler: the restricted peer lists of step 2 derive from the replicated selection
    counters, so every rank computes the same schedule."""
    single = _run(1, "cpu", rounds=9, kinds="pens_native")
    assert any(v for v in single["pens_native"]["best"].values()), "step 2 was never reached"
    assert single["pens_native"]["cpp_executor"], "step 2 should have moved to the C++ executor"
    multi = _run(2, "cpu", rounds=9, kinds="pens_native")
    assert multi["pens_native"]["cpp_executor"]
    _compare(single, multi, rel=1e-5, skip=("cache_left",))
    # interrupted in step 1 (candidates cached on their receivers' ranks) and in step 2 (slots of the C++ executor)
    _compare(single, _run(2, "cpu", rounds=9, kinds="pens_native", checkpoint=True), rel=1e-5, skip=("cache_left",))
    single = _run(1, "cpu", rounds=14, kinds="pens_native")
    _compare(single, _run(2, "cpu", rounds=14, kinds="pens_native", checkpoint=True), rel=1e-5, skip=("cache_left",))


@pytest.mark.gpu
def test_pens_two_ranks_cuda_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = _run(1, "cuda:0", rounds=9, kinds="pens")
    multi = _run(2, "cuda", rounds=9, kinds="pens")
    _compare(single, multi, rel=2e-3, skip=("cache_left",))


@pytest.mark.gpu
def test_cpp_executor_two_ranks_cuda_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = _run(1, "cuda:0", rounds=4, kinds=XKINDS)
    multi = _run(2, "cuda", rounds=4, kinds=XKINDS)
    assert all(v["cpp_executor"] for v in multi.values())
    _compare(single, multi, rel=2e-3)
    # tiny arena segments: CUDA-IPC segments are added collectively mid-run
    _compare(single, _run(2, "cuda", rounds=4, kinds=XKINDS, arena_rows=8), rel=2e-3)
