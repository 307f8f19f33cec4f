"""Development probe (GPU box): parity of the BASELINE configs against the oracle on one file, then
timing of the same queries over a resident table.  Not a bench line: bench.py is the contract.

    python tests/scripts/perf_probe.py [row_groups=96] [steps=20]
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COLS = ["p_timestamp", "level", "latency_ms", "host", "bytes", "status", "duration_s", "cpu"]
DIR = os.environ.get("PQB_PROBE_DIR", "/tmp/pqb_probe")
RGS_PER_FILE = 8


def _gen(args):
    path, first, n = args
    from parseable_b200 import synth
    if not os.path.exists(path):
        synth.write_logs16(path, n_row_groups=n, first_rg=first, columns=COLS)
    return path


def ensure(nrg):
    os.makedirs(DIR, exist_ok=True)
    jobs, g = [], 0
    while g < nrg:
        n = min(RGS_PER_FILE, nrg - g)
        jobs.append((os.path.join(DIR, f"probe_{g:05d}_{n}.parquet"), g, n))
        g += n
    missing = [j for j in jobs if not os.path.exists(j[0])]
    if missing:
        t = time.time()
        with mp.get_context("spawn").Pool(max(1, min(len(missing), (os.cpu_count() or 2) - 1, 64))) as pool:
            pool.map(_gen, missing, chunksize=1)
        print(f"generated {len(missing)} files in {time.time() - t:.1f}s", flush=True)
    return [j[0] for j in jobs]


def main():
    nrg = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    only = sys.argv[3] if len(sys.argv) > 3 else ""          # substring filter on the query names
    parity = os.environ.get("PROBE_PARITY", "1") != "0"
    import numpy as np
    import pyarrow as pa
    from oracle.oracle import Oracle
    from parseable_b200.query import (DeviceTable, StandardTableProvider, col, count_star, max_, min_, sum_)
    from test_gpu_parity import assert_tables_equal
    files = ensure(nrg)
    schema = {"p_timestamp": pa.timestamp("ms"), "level": pa.string(), "latency_ms": pa.int64(), "host": pa.string(),
              "bytes": pa.int64(), "status": pa.int64(), "duration_s": pa.float64(), "cpu": pa.float64()}
    c2 = [(col("level") == "ERROR") & (col("latency_ms") > 100)]
    aggq = {
        "C3 GROUP BY host -> COUNT, SUM(bytes)": (["host"], [count_star(), sum_("bytes")], []),
        "C4 GROUP BY host,status -> 6 aggs": (["host", "status"], [count_star(), sum_("bytes"), min_("latency_ms"), max_("latency_ms"),
                                                                   sum_("duration_s"), max_("cpu")], []),
        "GROUP BY level,status -> COUNT, SUM/MIN/MAX(bytes)": (["level", "status"], [count_star(), sum_("bytes"), min_("bytes"), max_("bytes")], []),
        "WHERE level='ERROR' GROUP BY host -> COUNT, SUM(bytes)": (["host"], [count_star(), sum_("bytes")], [col("level") == "ERROR"]),
        "global SUM(bytes), MAX(cpu) WHERE status=500": ([], [count_star(), sum_("bytes"), max_("cpu")], [col("status") == 500]),
    }
    aggq = {k: v for k, v in aggq.items() if only in k}
    # ---- parity on the first file ----
    if parity:
        ora = Oracle.from_parquet(files[0], columns=COLS)
        p1 = StandardTableProvider([files[0]], schema=schema)
        r = p1.scan(filters=c2)
        ids = np.concatenate([b.column(0).to_numpy() for b in r.batches]) if r.batches else np.array([], np.int64)
        assert np.array_equal(ids, ora.row_ids(c2)), "C2 row ids differ from the oracle"
        print("parity ok: C2 row ids", len(ids), flush=True)
        for name, (keys, aggs, flt) in aggq.items():
            got = p1.aggregate(keys, aggs, flt)
            assert_tables_equal(got.table(), ora.group_by(keys, aggs, flt), keys)
            print("parity ok:", name, flush=True)
    # ---- timing, table resident ----
    t0 = time.perf_counter()
    table = DeviceTable(files, COLS)
    print(f"table open: {1e3 * (time.perf_counter() - t0):.1f} ms, {table.rows} rows", flush=True)
    prov = StandardTableProvider(table, schema=schema)

    def run(name, fn):
        for _ in range(3):
            r = fn()
        ms = []
        for _ in range(steps):
            t = time.perf_counter()
            r = fn()
            ms.append(1e3 * (time.perf_counter() - t))
        ms.sort()
        m = r.metrics
        print(f"{name}: p50 {ms[len(ms) // 2]:.3f} ms = {table.rows / ms[len(ms) // 2] / 1e6:.1f} G rows/s | scan {m['scan_kernel_ms']:.3f} ms "
              f"device {m['device_ms']:.3f} host {m['host_ms']:.3f} | algo {m['algorithmic_bytes'] / 1e6:.1f} MB -> "
              f"{m['algorithmic_bytes'] / max(m['scan_kernel_ms'], 1e-6) / 1e6:.0f} GB/s | sel {m['rows_selected']} groups {m['groups']} launches {m['kernel_launches']}",
              flush=True)
        return r

    if only in "C2 filter -> row ids":
        run("C2 filter -> row ids", lambda: prov.scan(filters=c2))
    if only in "C2 filter count only":
        run("C2 filter count only", lambda: prov.scan(filters=c2, count_only=True))
    for name, (keys, aggs, flt) in aggq.items():
        run(name, lambda: prov.aggregate(keys, aggs, flt))
    table.close()


if __name__ == "__main__":
    main()
