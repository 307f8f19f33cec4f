#!/usr/bin/env python
"""bench.py — rows/s of the Parseable query hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over the whole synthetic table:
    configs[1]: 1-GPU Parquet scan+filter, 100M rows x 16 cols,
                WHERE level='ERROR' AND latency_ms>100  (+ the injected p_timestamp range, which
                footer statistics decide), output = selected row ordinals.
value  = rows scanned / step time with the encoded column chunks already resident in HBM
         (decode -> filter -> compaction -> row ids on the host), CUDA work timed per step by the
         library with CUDA events, step timed on the host around a device-synchronising call.
e2e    = same query through the same C-ABI call with the Parquet file images in page-locked HOST
         memory: footer parse, page walk, H2D of the referenced chunks, kernels, D2H every step.
--impl reference: the CPU restatement of the path (pyarrow Parquet decode + oracle.c scalar
         semantics, one worker per host core) on a bounded sample of the same files.

Launch: python bench.py [--gpus N --steps K --warmup W]; under torchrun one rank per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROW_GROUP = 262_144
N_ROW_GROUPS = 382                 # 100 139 008 rows ("100M")
RGS_PER_FILE = 16                  # one Parquet file per ingest minute batch in Parseable; 24 files here
DATA_DIR = os.environ.get("PQB_DATA_DIR", "/tmp/pqb_bench")
QUERY_COLS = ["p_timestamp", "level", "latency_ms"]
METRIC = "rows/sec filter+group-by over synthetic log Parquet; % HBM roofline"
WORKLOAD = "C2 scan+filter: 100M rows x 16 cols logs16, WHERE level='ERROR' AND latency_ms>100 -> row ids"


def _gen_one(args):
    path, first, n = args
    from parseable_b200 import synth
    if os.path.exists(path):
        return path
    synth.write_logs16(path, n_row_groups=n, first_rg=first)
    return path


def ensure_data(n_row_groups: int = N_ROW_GROUPS) -> list[str]:
    """Generate the 16-column logs16 files on this box (no dataset shipping)."""
    os.makedirs(DATA_DIR, exist_ok=True)
    jobs = []
    g = 0
    while g < n_row_groups:
        n = min(RGS_PER_FILE, n_row_groups - g)
        jobs.append((os.path.join(DATA_DIR, f"logs16_{g:05d}_{n}.parquet"), g, n))
        g += n
    missing = [j for j in jobs if not os.path.exists(j[0])]
    if missing:
        import multiprocessing as mp
        workers = max(1, min(len(missing), (os.cpu_count() or 2) - 1, 48))
        t = time.time()
        with mp.get_context("spawn").Pool(workers) as pool:
            pool.map(_gen_one, missing, chunksize=1)
        print(f"[bench] generated {len(missing)} files with {workers} workers in {time.time()-t:.1f}s", file=sys.stderr)
    return [j[0] for j in jobs]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def filters():
    from parseable_b200.query import col
    return [(col("level") == "ERROR") & (col("latency_ms") > 100)]


def time_range():
    from parseable_b200 import synth
    from parseable_b200.query import TimeRange
    # covers every row of the table -> decided by footer statistics, like a dashboard query over "last N hours"
    return TimeRange(synth.TS_BASE - (N_ROW_GROUPS + 2) * synth.RG_TS_STRIDE_MS, synth.TS_BASE + 1)


def all_filters():
    from parseable_b200.query import DEFAULT_TIMESTAMP_KEY, Timestamp, col
    tr = time_range()
    return filters() + [col(DEFAULT_TIMESTAMP_KEY) >= Timestamp(tr.start_ms), col(DEFAULT_TIMESTAMP_KEY) < Timestamp(tr.end_ms)]


# ------------------------------------------------------------------ CPU arms
def _cpu_worker(paths):
    import pyarrow as pa
    from oracle.oracle import Oracle
    rows = sel = 0
    for p in paths:
        o = Oracle.from_parquet(p, columns=QUERY_COLS)
        rows += o.n
        ids = o.row_ids(all_filters())
        sel += len(ids)
    return rows, sel


def cpu_port_throughput(files: list[str], workers: int):
    """oracle port over `files`, one process per worker; returns (rows/s, rows, selected, seconds)."""
    import multiprocessing as mp
    shards = [files[i::workers] for i in range(workers)]
    shards = [s for s in shards if s]
    with mp.get_context("spawn").Pool(len(shards)) as pool:
        pool.map(_cpu_worker, [[s[0]] for s in shards][:1])      # warm the pool / page cache / imports
        t = time.time()
        res = pool.map(_cpu_worker, shards)
        dt = time.time() - t
    rows = sum(r[0] for r in res)
    return rows / dt, rows, sum(r[1] for r in res), dt


def acero_throughput(files: list[str]):
    """Declared stand-in of BASELINE.md §3: pyarrow/Acero scan with the same predicate, all threads."""
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.dataset as ds
    pa.set_cpu_count(os.cpu_count() or 1)
    pa.set_io_thread_count(os.cpu_count() or 1)
    d = ds.dataset(files, format="parquet")
    tr = time_range()
    expr = ((pc.field("level") == "ERROR") & (pc.field("latency_ms") > 100) &
            (pc.field("p_timestamp") >= pa.scalar(tr.start_ms, pa.timestamp("ms"))) &
            (pc.field("p_timestamp") < pa.scalar(tr.end_ms, pa.timestamp("ms"))))
    d.to_table(columns=["latency_ms"], filter=expr)             # warm
    t = time.time()
    tb = d.to_table(columns=["latency_ms"], filter=expr)
    dt = time.time() - t
    rows = sum(f.metadata.num_rows for f in d.get_fragments())
    return rows / dt, rows, tb.num_rows, dt


def groupby_section(files, steps: int = 8, warmup: int = 3):
    """Secondary numbers, N=1 only (not the headline): BASELINE.json configs[2] and two log-analytics
    group-bys over the same 100 M-row files, table resident.  Parity of exactly these queries against
    the oracle: tests/scripts/c3_groupby.py and tests/test_gpu_parity.py."""
    import pyarrow as pa
    from parseable_b200.query import (DeviceTable, StandardTableProvider, col, count_star, max_, min_, sum_)
    cols = ["host", "bytes", "level", "status"]
    schema = {"host": pa.string(), "bytes": pa.int64(), "level": pa.string(), "status": pa.int64()}
    table = DeviceTable(files, cols)
    prov = StandardTableProvider(table, schema=schema)
    queries = [
        ("C3: SELECT host, COUNT(*), SUM(bytes) GROUP BY host (10k groups)", ["host"], [count_star(), sum_("bytes")], []),
        ("SELECT level, status, COUNT(*), SUM/MIN/MAX(bytes) GROUP BY level, status (25 groups)", ["level", "status"],
         [count_star(), sum_("bytes"), min_("bytes"), max_("bytes")], []),
        ("SELECT host, COUNT(*), SUM(bytes) WHERE level='ERROR' GROUP BY host", ["host"], [count_star(), sum_("bytes")],
         [col("level") == "ERROR"]),
    ]
    out = []
    for name, keys, aggs, flt in queries:
        for _ in range(warmup):
            r = prov.aggregate(keys, aggs, flt)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = prov.aggregate(keys, aggs, flt)
        dt = (time.perf_counter() - t0) / steps
        out.append({"query": name, "value": table.rows / dt, "unit": "rows/s", "ms_per_step": dt * 1e3,
                    "k_scan_ms": r.metrics["scan_kernel_ms"], "device_ms": r.metrics["device_ms"], "groups": r.metrics["groups"]})
        print(f"[bench] group-by: {name}: {table.rows / dt / 1e9:.1f} G rows/s ({dt * 1e3:.2f} ms/step, k_scan {r.metrics['scan_kernel_ms']:.2f} ms)",
              file=sys.stderr)
    table.close()
    return out


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    files = ensure_data()
    cores = os.cpu_count() or 1
    # bounded sample: as many whole files as keep one step around 10-20 s of CPU work
    sample = files[: max(1, min(len(files), cores // 2 if cores >= 8 else 2))]
    # one step is one pass of the CPU port over the sample (seconds, not milliseconds): one warm-up pass
    # (imports, page cache) and at most --steps timed passes inside a ~150 s budget, so the arm ends
    # within a few minutes whatever K the GPU arm was given
    vals = []
    warm = min(args.warmup, 1)
    t_start = time.time()
    i = 0
    while len(vals) < max(1, args.steps):
        last = cpu_port_throughput(sample, workers=min(cores, len(sample)))
        if i >= warm:
            vals.append(last)
        i += 1
        if vals and time.time() - t_start > 150.0:
            break
    steps_done = len(vals)
    v = sum(x[0] for x in vals) / len(vals)
    ms = 1000.0 * sum(x[3] for x in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": steps_done, "steps_requested": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i64/utf8-dictionary", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": vals[-1][1]},
        "cpu_baseline": {"value": v, "unit": "rows/s", "cores": min(cores, len(sample)), "kind": "port",
                         "sample": f"{len(sample)} of {len(files)} files ({vals[-1][1]} rows), pyarrow decode + oracle.c, one process per file"},
        "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # ~0.2 s timed: several clock samples land inside it
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--row-groups", type=int, default=N_ROW_GROUPS, help="smaller tables for development runs")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-groupby", action="store_true", help="skip the secondary group-by measurements (N=1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # data: rank 0 generates, everyone waits
    if rank == 0:
        files = ensure_data(args.row_groups)
    barrier()
    files = ensure_data(args.row_groups)

    import ctypes as C
    from parseable_b200 import _lib as L
    from parseable_b200.query import DeviceTable, HostFile, StandardTableProvider
    import pyarrow as pa
    lib = L.load()
    dev = (C.c_int * 1)(local_rank)
    rc = lib.pq_init(dev, 1)
    if rc != 0:
        raise SystemExit(f"pq_init failed: {lib.pq_last_error(None)}")
    if world > 1:
        ident = [None]
        if rank == 0:
            buf = C.create_string_buffer(L.PQ_COMM_ID_BYTES)
            assert lib.pq_comm_unique_id(buf) == 0
            ident[0] = buf.raw
        dist.broadcast_object_list(ident, src=0)
        assert lib.pq_comm_init_rank(ident[0], world, rank) == 0, lib.pq_last_error(None)

    schema = {"p_timestamp": pa.timestamp("ms"), "level": pa.string(), "latency_ms": pa.int64()}
    flt = all_filters()
    # weak scaling: every rank owns N_ROW_GROUPS row groups (the same synthetic files stand in for its
    # shard of a world x 100M-row table; no data-path collective in a filter scan)
    table = DeviceTable(files, QUERY_COLS)
    prov = StandardTableProvider(table, schema=schema)
    rows_per_step = table.rows

    def step_resident():
        r = prov.scan(filters=flt)
        return r

    for _ in range(args.warmup):
        r = step_resident()
    sel_expected = sum(b.num_rows for b in r.batches)
    # full-size properties of the result (the oracle cannot decode 100 M rows in the tests' time budget):
    # row ids strictly ascending and in range, COUNT-only scan agrees, a checksum of per-file checksums agrees
    checks = {}
    if rank == 0:
        import numpy as np
        ids = np.concatenate([b.column(0).to_numpy() for b in r.batches]) if r.batches else np.array([], np.int64)
        assert len(ids) == sel_expected and (len(ids) == 0 or (ids[0] >= 0 and ids[-1] < rows_per_step))
        assert bool(np.all(np.diff(ids) > 0)), "row ids are not strictly ascending"
        assert prov.scan(filters=flt, count_only=True).metrics["rows_selected"] == sel_expected
        per_file = [StandardTableProvider([f], schema=schema).scan(filters=flt, count_only=True).metrics["rows_selected"] for f in files]
        assert sum(per_file) == sel_expected, (sum(per_file), sel_expected)
        checks = {"row_ids_strictly_ascending": True, "count_only_agrees": True, "sum_of_per_file_counts_agrees": True,
                  "selected_rows": int(sel_expected)}
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches = 0
    scan_ms = []
    dev_ms = []
    host_ms = []
    algo_bytes = r.metrics["algorithmic_bytes"]
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        r = step_resident()
        step_ms.append(1000.0 * (time.perf_counter() - ts))
        launches += r.metrics["kernel_launches"]
        scan_ms.append(r.metrics["scan_kernel_ms"])
        dev_ms.append(r.metrics["device_ms"])
        host_ms.append(r.metrics["host_ms"])
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    assert sum(b.num_rows for b in r.batches) == sel_expected
    d2h_res = r.metrics["d2h_bytes"]
    t_local = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    dt = float(t_local.item())
    ms_per_step = 1000.0 * dt / args.steps
    value = rows_per_step * world / (dt / args.steps)
    if rank == 0:
        print(f"[bench] resident step: wall {ms_per_step:.3f} ms = pq_query_open {sum(host_ms)/len(host_ms):.3f} ms "
              f"(device {sum(dev_ms)/len(dev_ms):.3f} ms, k_scan {sum(scan_ms)/len(scan_ms):.3f} ms) + binding/Arrow import",
              file=sys.stderr)

    # ---- e2e: host buffers (page-locked file images), H2D + D2H inside every step ----
    e2e = None
    if not args.skip_e2e:
        table.close()
        hfs = [HostFile(path=p, pinned=True) for p in files]
        prov_e = StandardTableProvider(hfs, schema=schema)
        for _ in range(2):
            re_ = prov_e.scan(filters=flt)
        barrier()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            re_ = prov_e.scan(filters=flt)
        barrier()
        dte = time.perf_counter() - t0
        te = torch.tensor([dte], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dte = float(te.item())
        assert sum(b.num_rows for b in re_.batches) == sel_expected
        if rank == 0:
            print(f"[bench] e2e step: wall {1000.0*dte/k:.3f} ms = pq_query_open {re_.metrics['host_ms']:.3f} ms "
                  f"(footers+page walk+H2D {re_.metrics['upload_ms']:.3f} ms, device {re_.metrics['device_ms']:.3f} ms) + binding",
                  file=sys.stderr)
        e2e = {"value": rows_per_step * world / (dte / k), "unit": "rows/s", "h2d_bytes_per_step": re_.metrics["h2d_bytes"],
               "d2h_bytes_per_step": re_.metrics["d2h_bytes"], "ms_per_step": 1000.0 * dte / k, "steps": k}
        for h in hfs:
            h.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    k_ms = sum(scan_ms) / len(scan_ms)
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_scan_dram_bytes_per_launch")
    except Exception:
        pass
    groupby = None
    if world == 1 and not args.skip_groupby:
        try:
            table.close()   # the C2 table: make room and keep the pool small
            groupby = groupby_section(files)
        except Exception as e:  # secondary numbers must never cost the headline line
            groupby = [{"error": repr(e)}]
    cpu_baseline = None
    if not args.skip_cpu:
        cores = os.cpu_count() or 1
        sample = files[: max(1, min(len(files), cores // 2 if cores >= 8 else 2))]
        v, rows, sel, secs = cpu_port_throughput(sample, workers=min(cores, len(sample)))
        # the CPU port's answer over its sample is also the checker of the GPU arm at full row-group size
        gpu_sel = sum(StandardTableProvider([f], schema=schema).scan(filters=flt, count_only=True).metrics["rows_selected"] for f in sample)
        assert gpu_sel == sel, f"GPU selected {gpu_sel} rows over the CPU sample, the oracle port {sel}"
        checks["oracle_port_count_over_cpu_sample_agrees"] = True
        cpu_baseline = {"value": v, "unit": "rows/s", "cores": min(cores, len(sample)), "kind": "port",
                        "sample": f"{len(sample)} of {len(files)} files ({rows} rows, {secs:.1f} s), pyarrow decode + oracle.c, one process per file"}
        try:
            av, arows, asel, asecs = acero_throughput(sample)
            cpu_baseline["acero_standin"] = {"value": av, "unit": "rows/s", "cores": cores, "rows": arows, "seconds": asecs,
                                             "note": "pyarrow/Acero dataset scan, not DataFusion (BASELINE.md §3)"}
        except Exception as e:  # pragma: no cover
            cpu_baseline["acero_standin"] = {"error": repr(e)}
    line = {
        "metric": METRIC,
        "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i64/utf8-dictionary (bit-packed indices)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": rows_per_step, "row_groups_per_gpu": args.row_groups,
                   "selected_rows": sel_expected, "l2": "inputs (encoded chunks read per step) larger than L2; no explicit flush",
                   "parallelism": f"row-group shards x{world}, no data-path collective for a filter scan"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_scan", "kernel_ms": k_ms, "algorithmic_bytes": algo_bytes,
                     "peak_kind": peak_kind, "device_ms_per_step": sum(dev_ms) / len(dev_ms)},
        "e2e": e2e, "gpu_launches": launches, "clocks": clk, "cpu_baseline": cpu_baseline,
        "d2h_bytes_per_step_resident": d2h_res,
    }
    try:   # spread of the timed steps (median + p10 / p90)
        q = sorted(step_ms)
        line["step_ms_quantiles"] = {"p10": q[len(q) // 10], "p50": q[len(q) // 2], "p90": q[(len(q) * 9) // 10]}
    except Exception:
        pass
    line["checks"] = checks
    if groupby is not None:
        line["groupby"] = groupby
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
