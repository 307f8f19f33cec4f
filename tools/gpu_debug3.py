import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from parseable_b200 import synth
from parseable_b200.query import *
os.makedirs("/tmp/pqb", exist_ok=True)
path = "/tmp/pqb/dbg3.parquet"
if not os.path.exists(path):
    synth.write_logs16(path, n_row_groups=1, rows_per_group=40_000, columns=["level", "latency_ms", "status"])
ora = Oracle.from_parquet(path)
prov = StandardTableProvider([path], schema=ora.table.schema)
checks = {
 "c2": [(col("level") == "ERROR") & (col("latency_ms") > 100)],
 "two_or_and": [((col("level") == "ERROR") | (col("level") == "FATAL")) & ((col("status") == 500) | (col("status") == 503))],
}
for name, flt in checks.items():
    got = [prov.scan(filters=flt, count_only=True).metrics["rows_selected"] for _ in range(2)]
    print(name, got, ora.count(flt), flush=True)
