#!/usr/bin/env python
"""GPU validation of the paths that were written after round 2's GPU budget was spent and are therefore opt-in on CUDA.

Run on a B200 (one GPU is enough; `gpurun --timeout 900 -- python tools/validate_optins_gpu.py`):

  1. C++ executor, UPDATE mode of partitioned and sampled models  (GOSSIPY_EXEC_PART_UPDATE=1)  vs the per-event executor
  2. PENS: hand-over of step 2 to the C++ executor                (GOSSIPY_EXEC_PENS_STEP2=1)    vs the per-event executor
  3. banked engine: pinned staging of the index vectors           (GOSSIPY_BANK_PINNED_IDX=1)    vs pageable copies, with timing

Each check prints OK / MISMATCH; exit code 0 only if all agree.  When they do: drop the gates in
`engine/stream_exec.py::eligible`, make the pinned staging the default in `engine/bank.py::_idx`, and add `x_part_update`,
`x_sampled_update` to the GPU kinds of `tests/test_multirank.py`.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GOSSIPY_EXEC_PART_UPDATE"] = "1"
os.environ["GOSSIPY_EXEC_PENS_STEP2"] = "1"


def executor_update_modes(dev):
    import torch
    import gossipy_b200 as g
    import test_stream_executor as T
    ok = True
    for kw in (dict(model="mlp", protocol="PUSH_PULL", partitioned=4, mode="UPDATE", n=8),
               dict(model="logreg", protocol="PUSH", partitioned=4, mode="UPDATE", faults=True, tokenized=True),
               dict(model="mlp", protocol="PUSH_PULL", sampled=.2, mode="UPDATE", n=8),
               dict(model="logreg", protocol="PULL", sampled=.5, mode="UPDATE", faults=True, sync=False)):
        a, ra = T._sim(False, device=dev, **kw)
        b, rb = T._sim(True, device=dev, **kw)
        torch.cuda.synchronize()
        try:
            assert "_stream_exec" in b.__dict__, "not taken by the executor"
            T._same(a, ra, b, rb, tol=1e-5)
            print("OK       executor UPDATE", kw)
        except AssertionError as exc:
            ok = False
            print("MISMATCH executor UPDATE", kw, exc)
        g.CACHE.clear()
    return ok


def pens_handover(dev):
    import torch
    import gossipy_b200 as g
    import test_native_scheduler as T
    g.GlobalSettings().set_device(dev)
    ok = True
    for faults in (False, True):
        ref, rep_ref, _ = T._pens_sim("native", rounds=9, step1_rounds=3, executor=False, faults=faults)
        rows_ref = {i: n.model_handler.row.clone() for i, n in ref.nodes.items()}
        g.CACHE.clear()
        sim, rep, _ = T._pens_sim("native", rounds=9, step1_rounds=3, faults=faults)
        torch.cuda.synchronize()
        good = "_stream_exec" in sim.__dict__ and all(
            torch.allclose(n.model_handler.row, rows_ref[i], rtol=1e-5, atol=1e-6) for i, n in sim.nodes.items())
        good = good and rep._sent_messages == rep_ref._sent_messages
        ok = ok and good
        print("OK      " if good else "MISMATCH", "PENS hand-over, faults =", faults)
        g.CACHE.clear()
    g.GlobalSettings().set_device("cpu")
    return ok


def bank_pinned():
    """Two fresh processes (the switch is read at import): same curve, time per round."""
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, GOSSIPY_BANK_PINNED_IDX=flag)
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "many_nodes.py"), "--nodes", "4141", "--rounds", "20",
                            "--impl", "banked", "--device", "cuda"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=800)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        out[flag] = line[-1] if line else "FAILED: " + r.stderr[-400:]
        print("bank, pinned index vectors =", flag, "|", out[flag][:260], "| %.0f s" % (time.time() - t0))
    import json
    try:
        a, b = json.loads(out["0"]), json.loads(out["1"])
        same = a["last_eval"] == b["last_eval"] and a["sent"] == b["sent"]
        print("OK      " if same else "MISMATCH", "bank results; rounds/s %.1f -> %.1f" % (a["rounds_per_s"], b["rounds_per_s"]))
        return same
    except Exception as exc:      # noqa: BLE001
        print("MISMATCH bank:", exc)
        return False


def main():
    import torch
    if not torch.cuda.is_available():
        sys.exit("needs a GPU")
    dev = "cuda:0"
    results = [executor_update_modes(dev), pens_handover(dev), bank_pinned()]
    print("ALL OK" if all(results) else "FAILURES", results)
    sys.exit(0 if all(results) else 1)


if __name__ == "__main__":
    main()
