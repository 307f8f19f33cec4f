#!/bin/bash
# 2-rank NCCL parity check outside pytest (bounded): tests/scripts/mgpu_run.sh
set -u
cd "$(dirname "$0")/../.."
mkdir -p /tmp/pqb
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from parseable_b200 import synth
for tag, rate in (("nn", 0.0), ("nulls", 0.02)):
    p = f"/tmp/pqb/small_{tag}.parquet"
    if not os.path.exists(p):
        synth.write_logs16(p, n_row_groups=3, rows_per_group=70_000, null_rate=rate)
    p = f"/tmp/pqb/one_{tag}.parquet"
    if not os.path.exists(p):
        synth.write_logs16(p, n_row_groups=1, first_rg=7 if rate else 5, rows_per_group=50_000, null_rate=rate)
PY
rm -f /tmp/pqb/nccl_id
timeout -s KILL ${1:-90} python tests/scripts/mgpu_check.py 0 2 /tmp/pqb/nccl_id /tmp/pqb/small_nulls.parquet /tmp/pqb/small_nn.parquet -- /tmp/pqb/one_nn.parquet /tmp/pqb/one_nulls.parquet > /tmp/pqb/r0.log 2>&1 &
P0=$!
timeout -s KILL ${1:-90} python tests/scripts/mgpu_check.py 1 2 /tmp/pqb/nccl_id /tmp/pqb/small_nulls.parquet /tmp/pqb/small_nn.parquet -- /tmp/pqb/one_nn.parquet /tmp/pqb/one_nulls.parquet > /tmp/pqb/r1.log 2>&1 &
P1=$!
wait $P0; echo "rank0 rc=$?"; wait $P1; echo "rank1 rc=$?"
tail -5 /tmp/pqb/r0.log; tail -5 /tmp/pqb/r1.log
