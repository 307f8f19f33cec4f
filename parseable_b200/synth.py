"""Deterministic synthetic log tables written the way Parseable writes them.

This is test/bench infrastructure (SURVEY.md §8d, table `logs16`).  The file
layout follows the reference's writer properties:

* row group 262 144 rows          (/root/reference/src/cli.rs:425-431)
* `p_timestamp` first, DELTA_BINARY_PACKED, rows newest-first
                                   (src/parseable/streams.rs:584-590, src/utils/arrow/mod.rs:120-161)
* every other column dictionary-encoded with PLAIN fallback, data page v1,
  20 000-row page limit, 1 MiB dictionary limit (parquet-rs defaults that
  `WriterProperties::builder()` keeps, streams.rs:584)
* codec from P_PARQUET_COMPRESSION_ALGO (src/cli.rs:441-448); headline runs use
  UNCOMPRESSED, the Parseable default is LZ4_RAW.

RNG: ``numpy.random.Generator(PCG64(20260922 + row_group_index))`` so any row
group can be regenerated anywhere independently.
"""
from __future__ import annotations

import os
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

ROW_GROUP = 262_144
SEED0 = 20260922
TS_BASE = 1_700_000_000_000
RG_TS_STRIDE_MS = 60_000          # one minute of wall time per row group
TOKEN = "timeout-xyzzy"

LEVELS = ["DEBUG", "INFO", "WARN", "ERROR", "FATAL"]
LEVEL_P = [0.30, 0.55, 0.09, 0.05, 0.01]
STATUS = np.array([200, 301, 404, 500, 503], dtype=np.int64)
STATUS_P = [0.80, 0.05, 0.09, 0.05, 0.01]

I64_COLS = ["p_timestamp", "latency_ms", "bytes", "status"]
F64_COLS = ["cpu", "mem_gb", "duration_s", "score"]
STR_COLS = ["level", "host", "service", "region", "method", "path", "pod", "message"]
LOGS16_COLUMNS = I64_COLS + F64_COLS + STR_COLS


def _zipf_cdf(n: int, s: float = 1.1) -> np.ndarray:
    w = 1.0 / np.arange(1, n + 1) ** s
    return np.cumsum(w / w.sum())


_HOST_CDF = _zipf_cdf(10_000)
_DICTS: dict[str, np.ndarray] = {}


def _dictionary(name: str) -> np.ndarray:
    """Fixed value universe of each Utf8 column (seed independent of row group)."""
    if name in _DICTS:
        return _DICTS[name]
    if name == "level":
        d = LEVELS
    elif name == "host":
        d = [f"host-{i:05d}" for i in range(10_000)]
    elif name == "service":
        d = [f"svc-{i:03d}" for i in range(100)]
    elif name == "region":
        d = [f"region-{i:02d}" for i in range(16)]
    elif name == "method":
        d = ["GET", "POST", "PUT", "DELETE", "PATCH", "HEAD", "OPTIONS", "TRACE"]
    elif name == "path":
        d = [f"/api/v1/resource/{i:04d}" for i in range(1_000)]
    elif name == "pod":
        d = [f"pod-{i:04d}-{(i * 2654435761) & 0xffff:04x}" for i in range(5_000)]
    elif name == "message":
        rng = np.random.Generator(np.random.PCG64(SEED0 - 1))
        words = ["request", "completed", "failed", "retry", "upstream", "cache", "miss",
                 "hit", "db", "query", "slow", "user", "session", "token", "expired",
                 "connection", "reset", "peer", "handler", "panic", "ok", "queued"]
        d = []
        for i in range(8192):
            k = int(rng.integers(3, 8))
            ws = [words[int(j)] for j in rng.integers(0, len(words), k)]
            # exactly 8 of 8192 templates carry the token; their row share is
            # steered to 0.1 % in _message_idx()
            if i < 8:
                ws.insert(int(rng.integers(0, k)), TOKEN)
            d.append(f"[{i:04d}] " + " ".join(ws))
    else:
        raise KeyError(name)
    _DICTS[name] = np.array(d, dtype=object)
    return _DICTS[name]


def _message_idx(rng: np.random.Generator, n: int) -> np.ndarray:
    idx = rng.integers(8, 8192, n).astype(np.int32)
    k = n // 1000                      # exactly 0.1 % of the rows carry the token
    pos = rng.choice(n, size=k, replace=False)
    idx[pos] = rng.integers(0, 8, k).astype(np.int32)
    return idx


def _timestamps(rng: np.random.Generator, n: int, start: int) -> np.ndarray:
    """Non-increasing, piecewise constant: one stamp per ingest batch
    (src/utils/arrow/mod.rs:95-97), batches ~U[1,1000] rows, gaps ~U[1,50] ms."""
    nb = n // 200 + 16
    while True:
        lens = rng.integers(1, 1001, nb)
        if lens.sum() >= n:
            break
        nb *= 2
    gaps = rng.integers(1, 51, len(lens))
    stamps = start - np.cumsum(gaps)
    return np.repeat(stamps, lens)[:n].astype(np.int64)


def _dict_array(idx: np.ndarray, universe: np.ndarray, mask: np.ndarray | None) -> pa.Array:
    """Dictionary in first-occurrence order containing only used values, which is
    what parquet-rs builds from a plain Utf8 array."""
    if mask is not None:
        used_idx = idx[~mask]
    else:
        used_idx = idx
    uniq, first = np.unique(used_idx, return_index=True)
    order = np.argsort(first, kind="stable")
    uniq = uniq[order]
    remap = np.zeros(len(universe), dtype=np.int32)
    remap[uniq] = np.arange(len(uniq), dtype=np.int32)
    indices = pa.array(remap[idx], type=pa.int32(), mask=mask)
    return pa.DictionaryArray.from_arrays(indices, pa.array(universe[uniq], type=pa.string()))


def logs16_schema() -> pa.Schema:
    fields = [pa.field("p_timestamp", pa.timestamp("ms"), True)]
    fields += [pa.field(c, pa.int64(), True) for c in I64_COLS[1:]]
    fields += [pa.field(c, pa.float64(), True) for c in F64_COLS]
    fields += [pa.field(c, pa.string(), True) for c in STR_COLS]
    return pa.schema(fields)


def logs16_row_group(g: int, n: int = ROW_GROUP, null_rate: float = 0.0,
                     columns: list[str] | None = None) -> pa.Table:
    """Row group ``g`` of the logs16 table as an Arrow table (Utf8 columns are
    dictionary arrays).  Every column draws from its own stream so a column
    subset regenerates identical values."""
    cols = columns or LOGS16_COLUMNS
    arrays, names = [], []

    def stream(tag: int) -> np.random.Generator:
        return np.random.Generator(np.random.PCG64([SEED0 + g, tag]))

    def nulls(tag: int) -> np.ndarray | None:
        if null_rate <= 0.0:
            return None
        return stream(1000 + tag).random(n) < null_rate

    lat = None
    for name in cols:
        tag = LOGS16_COLUMNS.index(name)
        rng = stream(tag)
        m = nulls(tag)
        if name == "p_timestamp":
            v = _timestamps(rng, n, TS_BASE - g * RG_TS_STRIDE_MS)
            arr = pa.array(v, type=pa.timestamp("ms"), mask=m)
        elif name in ("latency_ms", "duration_s"):
            if lat is None:
                lrng = stream(LOGS16_COLUMNS.index("latency_ms"))
                lat = np.floor(lrng.lognormal(3.5, 1.2, n)).clip(0, 60_000).astype(np.int64)
            if name == "latency_ms":
                arr = pa.array(lat, type=pa.int64(), mask=m)
            else:
                arr = pa.array(lat.astype(np.float64) / 1000.0, type=pa.float64(), mask=m)
        elif name == "bytes":
            arr = pa.array(rng.integers(0, 65_536, n).astype(np.int64), mask=m)
        elif name == "status":
            arr = pa.array(STATUS[rng.choice(5, size=n, p=STATUS_P)], mask=m)
        elif name == "cpu":
            arr = pa.array(rng.random(n), mask=m)
        elif name == "mem_gb":
            arr = pa.array(rng.random(n) * 64.0, mask=m)
        elif name == "score":
            arr = pa.array(rng.standard_normal(n), mask=m)
        elif name == "level":
            arr = _dict_array(rng.choice(5, size=n, p=LEVEL_P).astype(np.int32),
                              _dictionary(name), m)
        elif name == "host":
            arr = _dict_array(np.searchsorted(_HOST_CDF, rng.random(n)).astype(np.int32),
                              _dictionary(name), m)
        elif name == "message":
            arr = _dict_array(_message_idx(rng, n), _dictionary(name), m)
        elif name in STR_COLS:
            u = _dictionary(name)
            arr = _dict_array(rng.integers(0, len(u), n).astype(np.int32), u, m)
        else:
            raise KeyError(name)
        arrays.append(arr)
        names.append(name)
    return pa.table(arrays, names=names)


def parseable_writer_kwargs(schema_names: list[str], compression: str = "NONE",
                            time_col: str = "p_timestamp") -> dict:
    """pyarrow equivalents of Stream::parquet_writer_props
    (src/parseable/streams.rs:572-631) on top of parquet-rs defaults."""
    kw = dict(
        compression=compression,
        use_dictionary=[c for c in schema_names if c != time_col],
        data_page_version="1.0",
        data_page_size=1 << 20,
        dictionary_pagesize_limit=1 << 20,
        write_batch_size=1024,
        max_rows_per_page=20_000,
        write_statistics=True,
        store_schema=True,
    )
    if time_col in schema_names:
        kw["column_encoding"] = {time_col: "DELTA_BINARY_PACKED"}
        kw["sorting_columns"] = [pq.SortingColumn(schema_names.index(time_col), descending=True,
                                                  nulls_first=False)]
    return kw


def write_logs16(path: str, n_row_groups: int, first_rg: int = 0, rows_per_group: int = ROW_GROUP,
                 null_rate: float = 0.0, compression: str = "NONE",
                 columns: list[str] | None = None) -> int:
    """Write row groups [first_rg, first_rg+n_row_groups) to one Parquet file; returns rows."""
    cols = columns or LOGS16_COLUMNS
    schema = pa.schema([pa.field(f.name, pa.dictionary(pa.int32(), pa.string()) if f.name in STR_COLS
                                 else f.type, True)
                        for f in (logs16_schema().field(c) for c in cols)])
    tmp = path + ".part"            # same .part→rename atomicity as streams.rs:771-787
    rows = 0
    with pq.ParquetWriter(tmp, schema, **parseable_writer_kwargs(cols, compression)) as w:
        for g in range(first_rg, first_rg + n_row_groups):
            t = logs16_row_group(g, rows_per_group, null_rate, cols)
            t = t.cast(schema)
            w.write_table(t, row_group_size=rows_per_group)
            rows += t.num_rows
    os.replace(tmp, path)
    return rows
