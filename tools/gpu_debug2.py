import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from parseable_b200 import synth
from parseable_b200.query import *
os.makedirs("/tmp/pqb", exist_ok=True)
path = "/tmp/pqb/dbg2.parquet"
if not os.path.exists(path):
    synth.write_logs16(path, n_row_groups=3, rows_per_group=70_000)
ora = Oracle.from_parquet(path)
prov = StandardTableProvider([path], schema=ora.table.schema)
checks = {
 "lvl_or": [(col("level") == "ERROR") | (col("level") == "FATAL")],
 "status_or": [(col("status") == 500) | (col("status") == 503)],
 "not_region": [~(col("region") == "region-00")],
 "bytes": [col("bytes") > 10],
 "two_or_and": [((col("level") == "ERROR") | (col("level") == "FATAL")) & ((col("status") == 500) | (col("status") == 503))],
 "deep": [((col("level") == "ERROR") | (col("level") == "FATAL")) & ((col("status") == 500) | (col("status") == 503)) & ~(col("region") == "region-00") & (col("bytes") > 10)],
 "c2": [(col("level") == "ERROR") & (col("latency_ms") > 100)],
 "or2col": [(col("level") == "FATAL") | (col("bytes") < 1000)],
}
for name, flt in checks.items():
    got = [prov.scan(filters=flt, count_only=True).metrics["rows_selected"] for _ in range(3)]
    print(os.environ.get("PQB_SYNC_CTL", "0"), name, got, ora.count(flt), flush=True)
