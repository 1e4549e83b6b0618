"""The experiment scripts (one per reference ``main_*.py``) run end to end on CPU at reduced size."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, **env):
    e = dict(os.environ, GOSSIPY_DEVICE="cpu", GOSSIPY_ROUNDS="3", GOSSIPY_NODES="10", GOSSIPY_SAMPLES="300",
             GOSSIPY_EPOCHS="1", OMP_NUM_THREADS="2")
    e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], capture_output=True, text=True,
                         timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "rounds evaluated: 3" in out.stdout, out.stdout[-1500:]
    return out.stdout


@pytest.mark.parametrize("script", ["main_ormandi_2013.py", "main_giaretta_2019.py", "main_berta_2014.py",
                                    "main_hegedus_2020.py", "main_danner_2023.py", "main_all2all.py"])
def test_reference_scripts_python_engine(script):
    _run(script)


def test_tokenized_partitioned_script_native_engine():
    from gossipy_b200.ops.native import native_available
    if not native_available():
        pytest.skip("extension not built")
    _run("main_hegedus_2021.py", GOSSIPY_ENGINE="native", GOSSIPY_ROUNDS=3)


def test_all2all_synchronous_rounds_and_pens():
    out = _run("main_all2all.py", GOSSIPY_SYNC=1, GOSSIPY_NODES=4)
    assert "sent=36" in out                     # 3 rounds x 4 nodes x 3 peers
    _run("main_onoszko_2021.py", GOSSIPY_NODES=4)
    from gossipy_b200.ops.native import native_available
    if native_available():                      # PENS on the C++ schedule (step switch between two pieces of a round)
        _run("main_onoszko_2021.py", GOSSIPY_NODES=4, GOSSIPY_ENGINE="native")


def test_danner_script_with_the_cpp_executor():
    """main_danner_2023 (LimitedMergeTMH, churn, delays) is eligible for csrc/exec."""
    from gossipy_b200.ops.native import native_available
    if not native_available():
        pytest.skip("extension not built")
    a = _run("main_danner_2023.py", GOSSIPY_ENGINE="native", GOSSIPY_EXECUTOR="native")
    b = _run("main_danner_2023.py", GOSSIPY_ENGINE="native")
    assert [l for l in a.splitlines() if l.startswith("last evaluation")] == \
        [l for l in b.splitlines() if l.startswith("last evaluation")]
