"""Ad-hoc GPU check used during development (run under gpurun)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyarrow as pa
from oracle.oracle import Oracle
from parseable_b200 import synth
from parseable_b200.query import *

def main():
    os.makedirs("/tmp/pqb", exist_ok=True)
    cols = ["p_timestamp", "latency_ms", "bytes", "status", "cpu", "duration_s", "level", "host", "message"]
    for null_rate, tag in ((0.0, "nn"), (0.02, "nulls")):
        path = f"/tmp/pqb/dbg_{tag}.parquet"
        synth.write_logs16(path, n_row_groups=3, rows_per_group=70_000, null_rate=null_rate, columns=cols)
        ora = Oracle.from_parquet(path, columns=cols)
        prov = StandardTableProvider([path], schema=ora.table.schema)
        checks = [
            ("C2", [(col("level") == "ERROR") & (col("latency_ms") > 100)]),
            ("status", [col("status") == 200]),
            ("or", [(col("level") == "FATAL") | (col("bytes") < 1000)]),
            ("not", [~(col("level") == "INFO")]),
            ("cpu", [col("cpu") > 0.5]),
            ("isnull", [col("host").is_null()]),
            ("like", [col("message").like("%timeout-xyzzy%")]),
            ("none", []),
        ]
        for name, flt in checks:
            try:
                t = time.time()
                got = prov.scan(filters=flt, count_only=True)
                want = ora.count(flt)
                ok = got.metrics["rows_selected"] == want
                print(f"[{tag}] {name}: gpu={got.metrics['rows_selected']} oracle={want} {'OK' if ok else 'MISMATCH'} "
                      f"scan_ms={got.metrics['scan_kernel_ms']:.3f} dev_ms={got.metrics['device_ms']:.3f} wall={time.time()-t:.3f}")
            except Exception as e:
                print(f"[{tag}] {name}: EXC {e}")
        # row ids
        try:
            flt = [(col("level") == "ERROR") & (col("latency_ms") > 100)]
            got = prov.scan(filters=flt).table()
            ids = got["__row_id"].to_numpy() if got.num_rows else []
            want = ora.row_ids(flt)
            print(f"[{tag}] row_ids: n={len(ids)} want={len(want)} {'OK' if list(ids)==list(want) else 'MISMATCH'}")
        except Exception as e:
            print(f"[{tag}] row_ids EXC {e}")
        aggsets = [
            (["host"], [count_star(), sum_("bytes")], []),
            (["host", "status"], [count_star(), sum_("bytes"), min_("latency_ms"), max_("latency_ms"), sum_("duration_s"), max_("cpu")], []),
            (["level"], [count_star(), avg("latency_ms"), count("cpu")], [col("status") == 200]),
            ([], [count_star(), sum_("bytes"), min_("cpu")], [col("level") == "ERROR"]),
        ]
        for keys, aggs, flt in aggsets:
            try:
                t = time.time()
                r = prov.aggregate(keys, aggs, flt)
                res = r.table()
                exp = ora.group_by(keys, aggs, flt)
                if keys:
                    sk = [(k, "ascending") for k in keys]
                    res, exp = res.sort_by(sk), exp.sort_by(sk)
                ok = res.num_rows == exp.num_rows
                bad = []
                for name in exp.column_names:
                    a, b = res[name].to_pylist(), exp[name].to_pylist()
                    if a != b:
                        # f64 sums: tolerance
                        import math
                        if all((x is None and y is None) or (x is not None and y is not None and math.isclose(x, y, rel_tol=1e-9, abs_tol=0)) for x, y in zip(a, b)) and len(a)==len(b):
                            continue
                        bad.append(name)
                print(f"[{tag}] agg {keys} rows={res.num_rows}/{exp.num_rows} {'OK' if ok and not bad else 'MISMATCH '+str(bad)} "
                      f"scan_ms={r.metrics['scan_kernel_ms']:.3f} dev_ms={r.metrics['device_ms']:.3f} wall={time.time()-t:.3f}")
                if bad:
                    for name in bad[:2]:
                        a, b = res[name].to_pylist(), exp[name].to_pylist()
                        d = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:5]
                        print("   ", name, d)
            except Exception as e:
                traceback.print_exc()
                print(f"[{tag}] agg {keys}: EXC {e}")

if __name__ == "__main__":
    main()
