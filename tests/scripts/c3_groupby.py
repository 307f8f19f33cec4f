"""BASELINE.json configs[2]: SELECT host, COUNT(*), SUM(bytes) GROUP BY host over the bench table
(100 M rows, 10 000 distinct hosts), table resident in HBM; plus configs-style filtered group-by.
Checks one file against the oracle, then times the whole table."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/scripts/ -> repo root
sys.path.insert(0, ROOT)
import pyarrow as pa
import bench
from oracle.oracle import Oracle
from parseable_b200.query import *
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import assert_tables_equal

def main():
    nrg = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_ROW_GROUPS
    files = bench.ensure_data(nrg)
    cols = ["host", "bytes", "level", "status"]
    schema = {"host": pa.string(), "bytes": pa.int64(), "level": pa.string(), "status": pa.int64()}
    queries = {
        "c3: GROUP BY host -> COUNT(*), SUM(bytes)": (["host"], [count_star(), sum_("bytes")], []),
        "GROUP BY level, status -> COUNT(*), SUM/MIN/MAX(bytes)": (["level", "status"], [count_star(), sum_("bytes"), min_("bytes"), max_("bytes")], []),
        "WHERE level='ERROR' GROUP BY host -> COUNT(*), SUM(bytes)": (["host"], [count_star(), sum_("bytes")], [col("level") == "ERROR"]),
    }
    # parity on the first file (16 row groups)
    ora = Oracle.from_parquet(files[0], columns=cols)
    prov1 = StandardTableProvider([files[0]], schema=schema)
    for name, (keys, aggs, flt) in queries.items():
        got = prov1.aggregate(keys, aggs, flt)
        assert_tables_equal(got.table(), ora.group_by(keys, aggs, flt), keys)
        print("parity ok:", name, flush=True)
    table = DeviceTable(files, cols)
    prov = StandardTableProvider(table, schema=schema)
    for name, (keys, aggs, flt) in queries.items():
        for _ in range(3): r = prov.aggregate(keys, aggs, flt)
        t0 = time.perf_counter(); n = 10
        for _ in range(n): r = prov.aggregate(keys, aggs, flt)
        dt = (time.perf_counter() - t0) / n
        print(f"{name}: {table.rows/dt/1e9:.1f} G rows/s, {dt*1e3:.3f} ms/step (k_scan {r.metrics['scan_kernel_ms']:.3f} ms, device {r.metrics['device_ms']:.3f} ms), groups {r.metrics['groups']}", flush=True)


if __name__ == "__main__":
    main()
