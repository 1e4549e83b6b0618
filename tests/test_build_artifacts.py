"""The built extension carries sm_100a code on the hardware paths the design names (checked without a GPU: cuobjdump)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def census():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    import sass_census
    k = sass_census.census()
    if k is None:
        pytest.skip("extension not built")
    return k


def _col(row, prefix):
    return sum(v for c, v in row.items() if c.startswith(prefix))


def test_training_and_evaluation_kernels_use_tcgen05_tmem_and_bulk_copies(census):
    tc = {k: v for k, v in census.items() if "mlp1_train_tc4_kernel" in k or "mlp1_train_tc3_kernel" in k
          or "mlp1_eval_tc_kernel" in k}
    assert len(tc) >= 6, sorted(census)
    for name, row in tc.items():
        assert _col(row, "UTC*MMA") > 0, name                   # tcgen05.mma
        assert _col(row, "LDTM") > 0, name                      # tcgen05.ld (accumulators come back from TMEM)
        assert _col(row, "UBLKCP") > 0, name                    # cp.async.bulk operand staging
        assert _col(row, "SYNCS") > 0 and _col(row, "UTCBAR") > 0, name      # mbarriers, tcgen05.commit
        assert _col(row, "generic LD/ST") == 0, name            # no pointer lost its address space
        assert _col(row, "HMMA") == 0, name                     # no legacy mma.sync
    train = [v for k, v in tc.items() if "mlp1_train_tc4_kernel" in k]
    assert all(_col(v, "STAS") > 0 and _col(v, "UCGABAR") > 0 for v in train)     # st.async over DSMEM, cluster barriers
    assert any(_col(v, "STTM") >= 8 for v in train)             # fp32 master weights / low parts written to TMEM


def test_collective_kernels_use_multimem(census):
    for name in ("allreduce_mean_kernel<true>", "allreduce_rs_ag_kernel<true>"):
        rows = [v for k, v in census.items() if name in k]
        assert rows and all(_col(v, "LDGMC") > 0 for v in rows), name        # multimem.ld_reduce through the switch


def test_every_kernel_family_is_in_the_binary(census):
    names = " ".join(census)
    for fam in ("merge_pair_kernel", "merge_kway_kernel", "merge_segments_kernel", "merge_indexed_kernel", "logreg_train_kernel",
                "linear_seq_kernel", "bank_deliver_kernel", "kmeans_match_merge_kernel", "mf_update_kernel", "mlp1_stage4_kernel",
                "keyed_perm_kernel", "flag_wait_kernel", "rank_barrier_kernel"):
        assert fam in names, fam
