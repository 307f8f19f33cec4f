"""C5 of BASELINE.json (configs[4]) on one GPU: a time range that cuts the first and the last row group AND
`message LIKE '%timeout-xyzzy%'` (0.1 % of the rows) -> compacted projection of {p_timestamp, host, message}.
Parity of the whole answer against the oracle on the first and the last file (the two the time range cuts),
row count against the generator's exact 0.1 %, then timing over the resident table.  Not a bench line.

    python tests/scripts/c5_probe.py [row_groups=384] [steps=20]
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COLS = ["p_timestamp", "host", "message"]
DIR = os.environ.get("PQB_PROBE_DIR", "/tmp/pqb_probe_c5")
RGS_PER_FILE = 8


def _gen(args):
    path, first, n = args
    from parseable_b200 import synth
    if not os.path.exists(path):
        synth.write_logs16(path, n_row_groups=n, first_rg=first, columns=COLS)
    return path


def ensure(nrg):
    os.makedirs(DIR, exist_ok=True)
    jobs = [(os.path.join(DIR, f"c5_{g:05d}.parquet"), g, min(RGS_PER_FILE, nrg - g)) for g in range(0, nrg, RGS_PER_FILE)]
    missing = [j for j in jobs if not os.path.exists(j[0])]
    if missing:
        t = time.time()
        with mp.get_context("spawn").Pool(max(1, min(len(missing), (os.cpu_count() or 2) - 1, 64))) as pool:
            pool.map(_gen, missing, chunksize=1)
        print(f"generated {len(missing)} files in {time.time() - t:.1f}s", flush=True)
    return [j[0] for j in jobs]


def main():
    nrg = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import numpy as np
    import pyarrow as pa
    from oracle.oracle import Oracle
    from parseable_b200 import synth
    from parseable_b200.query import DeviceTable, StandardTableProvider, Timestamp, col
    files = ensure(nrg)
    schema = {"p_timestamp": pa.timestamp("ms"), "host": pa.string(), "message": pa.string()}
    lo = synth.TS_BASE - (nrg - 1) * synth.RG_TS_STRIDE_MS - 6_000     # inside the last row group
    hi = synth.TS_BASE - 6_000                                          # inside the first one
    flt = [col("message").like(f"%{synth.TOKEN}%"), col("p_timestamp") >= Timestamp(lo), col("p_timestamp") < Timestamp(hi)]

    def as_py(t):
        return {c: (t[c].cast(pa.string()) if pa.types.is_dictionary(t[c].type) else t[c]).to_pylist() for c in COLS}

    # ---- parity on the two files the range cuts ----
    for f in (files[0], files[-1]):
        ora = Oracle.from_parquet(f, columns=COLS)
        ids = ora.row_ids(flt)
        exp = as_py(ora.table.take(pa.array(ids)).select(COLS))
        res = StandardTableProvider([f], schema=schema).scan(projection=COLS, filters=flt, row_ids=True)
        got = res.table()
        assert np.array_equal(got["__row_id"].to_numpy(), ids), f
        assert as_py(got) == exp, f
        n_like = ora.count(flt[:1])
        assert 0 < len(ids) < n_like, (len(ids), n_like)                # the time range really cuts this file
        print(f"parity ok: {os.path.basename(f)}: {len(ids)} rows of {n_like} LIKE matches inside the range", flush=True)

    t0 = time.perf_counter()
    table = DeviceTable(files, COLS)
    print(f"table open: {1e3 * (time.perf_counter() - t0):.1f} ms, {table.rows} rows", flush=True)
    prov = StandardTableProvider(table, schema=schema)
    r = prov.scan(filters=flt[:1], count_only=True)
    assert r.metrics["rows_selected"] == (synth.ROW_GROUP // 1000) * nrg, (r.metrics["rows_selected"], table.rows)   # the generator's exact 0.1 % of every row group
    for name, fn in (("C5 time range + LIKE -> {p_timestamp, host, message}", lambda: prov.scan(projection=COLS, filters=flt)),
                     ("C5 time range + LIKE -> row ids", lambda: prov.scan(filters=flt)),
                     ("C5 time range + LIKE -> count", lambda: prov.scan(filters=flt, count_only=True))):
        for _ in range(3):
            r = fn()
        ms = []
        for _ in range(steps):
            t = time.perf_counter()
            r = fn()
            ms.append(1e3 * (time.perf_counter() - t))
        ms.sort()
        m = r.metrics
        p50 = ms[len(ms) // 2]
        print(f"{name}: p50 {p50:.3f} ms = {table.rows / p50 / 1e6:.1f} G rows/s | scan {m['scan_kernel_ms']:.3f} ms device {m['device_ms']:.3f} "
              f"host {m['host_ms']:.3f} | algo {m['algorithmic_bytes'] / 1e6:.1f} MB -> {m['algorithmic_bytes'] / max(m['scan_kernel_ms'], 1e-6) / 1e6:.0f} GB/s "
              f"| sel {m['rows_selected']} launches {m['kernel_launches']}", flush=True)
    table.close()


if __name__ == "__main__":
    main()
