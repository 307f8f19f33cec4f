"""N>1 path on real GPUs: one process per GPU, NCCL all-reduce of partial tables inside the library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from parseable_b200 import _lib as L
    return L.load().pq_device_count()


def test_two_rank_allreduce_matches_oracle(small_files, tmp_path, built):
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 2
    idfile = str(tmp_path / "nccl_id")
    from parseable_b200 import synth
    split = []
    for tag, rate, rg in (("nn", 0.0, 5), ("nulls", 0.02, 7)):     # one row group each: rank 0's shard is NULL-free, rank 1's is not
        p = str(tmp_path / f"one_{tag}.parquet")
        synth.write_logs16(p, n_row_groups=1, first_rg=rg, rows_per_group=50_000, null_rate=rate)
        split.append(p)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "scripts", "mgpu_check.py"), str(r), str(n), idfile,
                               small_files["nulls"], small_files["nn"], "--"] + split, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(n)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o[-3000:]}"
        assert "parity OK" in o
