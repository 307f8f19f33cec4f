"""Race / work-distribution stress for k_scan: same queries under forced grid sizes
(PQB_GRID) so that CTAs take several items each in every order, compared with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/scripts/ -> repo root
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from parseable_b200 import synth
from parseable_b200.query import *
path = "/tmp/pqb/stress.parquet"
os.makedirs("/tmp/pqb", exist_ok=True)
if not os.path.exists(path):
    synth.write_logs16(path, n_row_groups=2, rows_per_group=70_000, columns=["level", "latency_ms", "status", "region", "bytes"])
ora = Oracle.from_parquet(path)
prov = StandardTableProvider([path], schema=ora.table.schema)
qs = {"c2": [(col("level") == "ERROR") & (col("latency_ms") > 100)], "c1": [(col("level") == "ERROR")]}
bad = 0
for g in sys.argv[1:] or ["1", "2", "3", "5", "8"]:
    os.environ["PQB_GRID"] = g
    for name, f in qs.items():
        exp = ora.count(f)
        got = [prov.scan(filters=f, count_only=True).metrics["rows_selected"] for _ in range(6)]
        ok = all(x == exp for x in got)
        bad += not ok
        print("grid", g, name, "ok" if ok else f"MISMATCH {got} != {exp}", flush=True)
sys.exit(1 if bad else 0)
