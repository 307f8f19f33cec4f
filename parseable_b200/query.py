"""Host-side mirror of Parseable's query surface over the C ABI.

Names follow the reference (paths relative to /root/reference):

* ``Query`` / ``execute``           src/query/mod.rs:143-157, 260-343
* ``StandardTableProvider.scan``   src/query/stream_schema_provider.rs:526-659
  (``projection``, ``filters``, ``limit``) plus ``aggregate`` for the
  FilterExec + AggregateExec stack DataFusion puts above the scan
* ``TimeRange`` filter injection   src/query/mod.rs:774-833 (``p_timestamp >= start AND p_timestamp < end``)

Everything here only builds a ``PqQueryDesc`` and hands it to
``libparseable_b200.so``; results come back through the Arrow C Data Interface
into pyarrow.  There is no CPU execution path in this module.
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass, field
from typing import Any, Iterable, Sequence

import pyarrow as pa

from . import _lib as L

DEFAULT_TIMESTAMP_KEY = "p_timestamp"  # src/event/mod.rs DEFAULT_TIMESTAMP_KEY


class QueryError(RuntimeError):
    """ExecuteError / DataFusionError::External of the reference (src/query/mod.rs:904-917)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{L.ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.message = msg


# ----------------------------------------------------------------------------- expressions
@dataclass
class Expr:
    kind: str                      # 'col' 'lit' 'cmp' 'and' 'or' 'not' 'is_null' 'is_not_null' 'like'
    args: tuple = ()
    op: int = 0
    flags: int = 0

    def _bin(self, other, op):
        return Expr("cmp", (self, _lit(other)), op)

    def __eq__(self, o): return self._bin(o, L.PQ_EQ)       # type: ignore[override]
    def __ne__(self, o): return self._bin(o, L.PQ_NE)       # type: ignore[override]
    def __lt__(self, o): return self._bin(o, L.PQ_LT)
    def __le__(self, o): return self._bin(o, L.PQ_LE)
    def __gt__(self, o): return self._bin(o, L.PQ_GT)
    def __ge__(self, o): return self._bin(o, L.PQ_GE)
    def __and__(self, o): return Expr("and", (self, o))
    def __or__(self, o): return Expr("or", (self, o))
    def __invert__(self): return Expr("not", (self,))
    def __hash__(self): return id(self)
    def is_null(self): return Expr("is_null", (self,))
    def is_not_null(self): return Expr("is_not_null", (self,))

    def like(self, pattern: str, negated=False, case_insensitive=False):
        f = (L.PQ_LIKE_NEGATED if negated else 0) | (L.PQ_LIKE_CASE_INSENSITIVE if case_insensitive else 0)
        return Expr("like", (self, Expr("lit", (pattern,))), flags=f)

    def ilike(self, pattern: str, negated=False):
        return self.like(pattern, negated, True)


def col(name: str) -> Expr:
    return Expr("col", (name,))


def lit(v: Any) -> Expr:
    return Expr("lit", (v,))


def _lit(v) -> Expr:
    return v if isinstance(v, Expr) else Expr("lit", (v,))


@dataclass
class Timestamp:
    """A TimestampMillisecond literal (what transform() injects, stream_schema_provider.rs:722-748)."""
    ms: int


_FLIP = {L.PQ_LT: L.PQ_GT, L.PQ_GT: L.PQ_LT, L.PQ_LE: L.PQ_GE, L.PQ_GE: L.PQ_LE, L.PQ_EQ: L.PQ_EQ, L.PQ_NE: L.PQ_NE}


@dataclass
class Agg:
    fn: str           # count_star count sum min max avg
    column: str | None = None


_AGG_CODE = {"count_star": L.PQ_AGG_COUNT_STAR, "count": L.PQ_AGG_COUNT, "sum": L.PQ_AGG_SUM,
             "min": L.PQ_AGG_MIN, "max": L.PQ_AGG_MAX, "avg": L.PQ_AGG_AVG}


@dataclass(frozen=True)
class DateBin:
    """GROUP BY DATE_BIN(width, column, origin): the counts / histogram API of the reference
    (src/query/mod.rs:623-680 builds `DATE_BIN('1m', p_timestamp, TIMESTAMP '1970-01-01 00:00:00+00')`)."""
    width_ms: int
    column: str = DEFAULT_TIMESTAMP_KEY
    origin_ms: int = 0

    @property
    def name(self) -> str:
        return f"date_bin({self.column})"


_INTERVALS = {"s": 1000, "m": 60_000, "h": 3_600_000, "d": 86_400_000}


def date_bin(width: str | int, column: str = DEFAULT_TIMESTAMP_KEY, origin_ms: int = 0) -> DateBin:
    """``date_bin("5m")``: widths like the reference's '1m' | '5m' | '1h' | '1d', or milliseconds."""
    if isinstance(width, str):
        width = int(width[:-1]) * _INTERVALS[width[-1]]
    return DateBin(int(width), column, origin_ms)


def count_star(): return Agg("count_star")
def count(c): return Agg("count", c)
def sum_(c): return Agg("sum", c)
def min_(c): return Agg("min", c)
def max_(c): return Agg("max", c)
def avg(c): return Agg("avg", c)


# ----------------------------------------------------------------------------- descriptor builder
class _Desc:
    """Keeps every ctypes object alive for the duration of the call."""

    def __init__(self):
        self.keep: list = []
        self.columns: list[str] = []

    def col_index(self, name: str) -> int:
        if name not in self.columns:
            self.columns.append(name)
        return self.columns.index(name)

    def literal(self, v) -> L.PqLiteral:
        out = L.PqLiteral()
        if v is None:
            out.type = L.PQ_T_NULL
        elif isinstance(v, bool):
            out.type, out.i64 = L.PQ_T_BOOL, int(v)
        elif isinstance(v, Timestamp):
            out.type, out.i64 = L.PQ_T_TS_MS, int(v.ms)
        elif isinstance(v, int):
            out.type, out.i64 = L.PQ_T_I64, v
        elif isinstance(v, float):
            out.type, out.f64 = L.PQ_T_F64, v
        elif isinstance(v, (str, bytes)):
            b = v.encode() if isinstance(v, str) else v
            buf = C.create_string_buffer(b, len(b) + 1)
            self.keep.append(buf)
            out.type = L.PQ_T_UTF8
            out.str = C.cast(buf, C.c_char_p)
            out.str_len = len(b)
        else:
            raise TypeError(f"unsupported literal {v!r}")
        return out

    def compile_pred(self, e: Expr, ops: list):
        if e.kind in ("and", "or"):
            self.compile_pred(e.args[0], ops)
            self.compile_pred(e.args[1], ops)
            ops.append(L.PqPredOp(kind=L.PQ_OP_AND if e.kind == "and" else L.PQ_OP_OR))
        elif e.kind == "not":
            self.compile_pred(e.args[0], ops)
            ops.append(L.PqPredOp(kind=L.PQ_OP_NOT))
        elif e.kind in ("is_null", "is_not_null"):
            c = e.args[0]
            if c.kind != "col":
                raise QueryError(L.PQ_ERR_UNSUPPORTED, "IS NULL on a non-column expression")
            ops.append(L.PqPredOp(kind=L.PQ_OP_IS_NULL if e.kind == "is_null" else L.PQ_OP_IS_NOT_NULL,
                                  col=self.col_index(c.args[0])))
        elif e.kind == "like":
            c, p = e.args
            if c.kind != "col":
                raise QueryError(L.PQ_ERR_UNSUPPORTED, "LIKE on a non-column expression")
            ops.append(L.PqPredOp(kind=L.PQ_OP_LIKE, col=self.col_index(c.args[0]), flags=e.flags,
                                  lit=self.literal(p.args[0])))
        elif e.kind == "cmp":
            a, b = e.args
            op = e.op
            if a.kind == "lit" and b.kind == "col":
                a, b, op = b, a, _FLIP[op]
            if a.kind != "col" or b.kind != "lit":
                raise QueryError(L.PQ_ERR_UNSUPPORTED, "only column <op> literal comparisons are pushed to the GPU")
            ops.append(L.PqPredOp(kind=L.PQ_OP_CMP, col=self.col_index(a.args[0]), cmp=op, lit=self.literal(b.args[0])))
        elif e.kind == "lit":
            ops.append(L.PqPredOp(kind=L.PQ_OP_CONST, lit=self.literal(e.args[0])))
        else:
            raise QueryError(L.PQ_ERR_UNSUPPORTED, f"expression {e.kind} in a predicate")


_ARROW_TO_PQ = {pa.int64(): L.PQ_T_I64, pa.float64(): L.PQ_T_F64, pa.string(): L.PQ_T_UTF8,
                pa.large_string(): L.PQ_T_UTF8, pa.bool_(): L.PQ_T_BOOL, pa.timestamp("ms"): L.PQ_T_TS_MS}


def _pq_type(t: pa.DataType | None) -> int:
    if t is None:
        return L.PQ_T_NULL
    if pa.types.is_dictionary(t):
        t = t.value_type
    if pa.types.is_timestamp(t):
        return L.PQ_T_TS_MS
    return _ARROW_TO_PQ.get(t, L.PQ_T_NULL)


# ----------------------------------------------------------------------------- files / tables
class HostFile:
    """A Parquet file image in host memory (page-locked when ``pinned``) or a path."""

    def __init__(self, path: str | None = None, data: bytes | None = None, pinned: bool = False):
        self.path = path
        self._pinned_ptr = None
        self._buf = None
        self.size = 0
        if data is not None or pinned:
            if data is None:
                with open(path, "rb") as f:
                    data = f.read()
            self.size = len(data)
            if pinned:
                lib = L.load()
                p = lib.pq_host_alloc(self.size)
                if not p:
                    raise QueryError(L.PQ_ERR_OOM, "pq_host_alloc failed")
                C.memmove(p, data, self.size)
                self._pinned_ptr = p
            else:
                self._buf = C.create_string_buffer(data, len(data))

    def as_pq(self) -> L.PqFile:
        f = L.PqFile()
        if self._pinned_ptr:
            f.buf, f.size = self._pinned_ptr, self.size
        elif self._buf is not None:
            f.buf, f.size = C.cast(self._buf, C.c_void_p), self.size
        else:
            self._path_b = self.path.encode()
            f.path = self._path_b
        return f

    def close(self):
        if self._pinned_ptr:
            L.load().pq_host_free(self._pinned_ptr)
            self._pinned_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _files_array(files: Sequence[HostFile | str]):
    hfs = [f if isinstance(f, HostFile) else HostFile(path=f) for f in files]
    arr = (L.PqFile * len(hfs))(*[h.as_pq() for h in hfs])
    return hfs, arr


def _staging_image(batches: list) -> "HostFile":
    """In-RAM staging record batches -> one Parquet file image (newest row first)."""
    import io

    import pyarrow.parquet as pq

    from . import synth
    rev = [b.take(pa.array(range(b.num_rows - 1, -1, -1), pa.int64())) for b in reversed(batches)]
    t = pa.Table.from_batches(rev)
    buf = io.BytesIO()
    kw = synth.parseable_writer_kwargs(t.column_names, time_col=DEFAULT_TIMESTAMP_KEY)
    pq.write_table(t, buf, row_group_size=synth.ROW_GROUP, **kw)
    return HostFile(data=buf.getvalue())


def field_stats(provider: "StandardTableProvider", field: str, max_field_statistics: int = 50, filters: Iterable[Expr] = ()):
    """Field statistics of one column like the reference's per-upload job (src/storage/field_stats.rs:298-330):
    ``GROUP BY field -> COUNT(*)`` runs on the GPU over every row; the window functions of the SQL (SUM / COUNT OVER (),
    ROW_NUMBER() OVER (ORDER BY value_count DESC)) range over the grouped result and stay above the scan.
    Returns (total_count, distinct_count, [(value, count), ...] for the max_field_statistics most frequent values)."""
    t = provider.aggregate([field], [count_star()], list(filters)).table()
    vals, cnts = t[field].to_pylist(), t["count(*)"].to_pylist()
    order = sorted(range(len(vals)), key=lambda i: -cnts[i])          # stable: ties keep the scan's (dictionary) order
    return sum(cnts), len(vals), [(vals[i], cnts[i]) for i in order[:max_field_statistics]]


class DeviceTable:
    """Encoded column chunks resident in HBM (pq_table_open): the hot tier of
    src/hottier.rs, one level closer to the kernels."""

    def __init__(self, files: Sequence[HostFile | str], columns: Sequence[str], shard_index=0, shard_count=1):
        lib = L.load()
        self._hfs, arr = _files_array(files)
        names = (C.c_char_p * len(columns))(*[c.encode() for c in columns])
        h = C.c_void_p()
        rc = lib.pq_table_open(arr, len(self._hfs), names, len(columns), shard_index, shard_count, C.byref(h))
        if rc != L.PQ_OK:
            raise QueryError(rc, (lib.pq_last_error(None) or b"").decode())
        self.handle = h
        self.columns = list(columns)

    @property
    def rows(self) -> int:
        return L.load().pq_table_rows(self.handle)

    @property
    def device_bytes(self) -> int:
        return L.load().pq_table_device_bytes(self.handle)

    def close(self):
        if self.handle:
            L.load().pq_table_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- execution
@dataclass
class QueryResult:
    batches: list[pa.RecordBatch]
    metrics: dict
    fields: list[str] = field(default_factory=list)
    json_text: bytes | None = None      # the result as JSON text formatted on the GPU (pq_query_json), when asked for

    def to_json(self, with_fields: bool = False, fill_null: bool = False):
        """QueryResponse::to_json (src/response.rs:31-58) over the GPU-formatted records: the list of row objects,
        optionally with every field present (NULL filled in) and wrapped with the field list."""
        import json as _json
        if self.json_text is None:
            raise ValueError("run the query with json='array' or json='lines'")
        txt = self.json_text.decode()
        rows = _json.loads(txt) if txt.startswith("[") else [_json.loads(line) for line in txt.splitlines() if line]
        if fill_null:
            for r in rows:
                for f in self.fields:
                    r.setdefault(f, None)
        return {"fields": self.fields, "records": rows} if with_fields else rows

    def table(self) -> pa.Table:
        if not self.batches:
            return pa.table({})
        return pa.Table.from_batches(self.batches)


class StandardTableProvider:
    """scan()/aggregate() over a file list or a resident DeviceTable.

    ``schema`` maps column name -> Arrow type the plan expects (the table schema
    Parseable keeps per stream); columns the plan names but a file lacks read as NULL.
    """

    def __init__(self, source: DeviceTable | Sequence[HostFile | str], schema: pa.Schema | dict | None = None,
                 shard_index: int = 0, shard_count: int = 1, staging_batches: Sequence[pa.RecordBatch] = (),
                 staging_parquet: Sequence[str] = ()):
        # ---- get_staging_execution_plan (stream_schema_provider.rs:242-298): data still in staging ----
        # staging Parquet files: newest first by file name, scanned like any other file; in-RAM staging batches: reversed
        # (batch order and row order, `reversed_mem_table` :686-695) and turned into ONE in-memory Parquet image with the
        # stream's writer properties -- the conversion Parseable itself runs when it flushes staging
        # (streams.rs:572-631) -- so that the GPU path stays the only reader.  They come first in the file list, like
        # the reference's plan order (staging arrow, staging parquet, then hot tier / object store).
        extra: list = []
        if staging_batches:
            extra.append(_staging_image(list(staging_batches)))
        extra += sorted(staging_parquet, reverse=True)
        if extra:
            if isinstance(source, DeviceTable):
                raise QueryError(L.PQ_ERR_INVALID_ARG, "staging data joins a file list, not a resident table")
            source = extra + list(source)
        self.source = source
        if isinstance(schema, pa.Schema):
            schema = {f.name: f.type for f in schema}
        self.schema = schema or {}
        self.shard_index, self.shard_count = shard_index, shard_count

    # -- TableProvider::scan -------------------------------------------------
    def scan(self, projection: Sequence[str] | None = None, filters: Iterable[Expr] = (), limit: int | None = None,
             count_only: bool = False, row_ids: bool | None = None, batch_size: int = 0, flags: int = 0,
             poll: bool = False, json: str | None = None) -> QueryResult:
        """``projection``: the columns to return for the selected rows (TableProvider::scan's projection);
        without one the scan returns the selected row ordinals (``__row_id``).  ``row_ids=True`` appends
        ``__row_id`` to a projection."""
        f = 0
        if row_ids is None:
            row_ids = not projection
        if count_only:
            f |= L.PQ_QUERY_COUNT_ONLY
        elif row_ids:
            f |= L.PQ_QUERY_EMIT_ROW_IDS
        return self._run(list(filters), [], [], list(projection or []), limit, batch_size, f | flags, poll=poll, json=json)

    # -- FilterExec + AggregateExec folded into the same call ----------------
    def aggregate(self, group_by: Sequence[str], aggs: Sequence[Agg], filters: Iterable[Expr] = (),
                  batch_size: int = 0, flags: int = 0, json: str | None = None) -> QueryResult:
        return self._run(list(filters), list(group_by), list(aggs), [], None, batch_size, flags, json=json)

    def count_distinct(self, group_by: Sequence[str], column: str, filters: Iterable[Expr] = ()) -> pa.Table:
        """``SELECT keys, COUNT(DISTINCT column)`` (Parseable's alerts use it: src/alerts/alert_enums.rs:216-223).
        The distinct values of a dictionary-encoded column are its interned ids: the GPU runs
        ``GROUP BY keys, column -> COUNT(*)`` (every row, one pass) and the per-key number of non-NULL ``column`` groups is
        counted over that small result above the scan.  NULLs do not count, an empty input yields 0 for the global form."""
        t = self.aggregate(list(group_by) + [column], [count_star()], filters).table()
        name = f"count(distinct {column})"
        if not group_by:
            n = sum(1 for v in t[column].to_pylist() if v is not None) if t.num_rows else 0
            return pa.table({name: pa.array([n], pa.int64())})
        if t.num_rows == 0:
            return pa.table({**{k: t[k] for k in group_by}, name: pa.array([], pa.int64())})
        seen: dict = {}
        for row in zip(*[t[k].to_pylist() for k in group_by], t[column].to_pylist()):
            key, v = row[:-1], row[-1]
            seen[key] = seen.get(key, 0) + (0 if v is None else 1)
        cols = {k: pa.array([key[i] for key in seen], t[k].type) for i, k in enumerate(group_by)}
        cols[name] = pa.array(list(seen.values()), pa.int64())
        return pa.table(cols)

    def _run(self, filters, group_by, aggs, projection, limit, batch_size, flags, poll: bool = False, json: str | None = None) -> QueryResult:
        lib = L.load()
        d = _Desc()
        ops: list = []
        pred = None
        for e in filters:                      # conjunction(filters), stream_schema_provider.rs:126
            pred = e if pred is None else Expr("and", (pred, e))
        if pred is not None:
            d.compile_pred(pred, ops)
        gb = [d.col_index(c.column if isinstance(c, DateBin) else c) for c in group_by]
        gx = None
        if any(isinstance(c, DateBin) for c in group_by):
            gx = (L.PqKeyExpr * len(group_by))()
            for i, c in enumerate(group_by):
                if isinstance(c, DateBin):
                    gx[i].kind, gx[i].width_ms, gx[i].origin_ms = L.PQ_KEY_DATE_BIN, c.width_ms, c.origin_ms
        ag = []
        for a in aggs:
            ag.append(L.PqAgg(fn=_AGG_CODE[a.fn], col=d.col_index(a.column) if a.column is not None else -1))
        proj = [d.col_index(c) for c in projection]

        desc = L.PqQueryDesc()
        hfs = None
        if isinstance(self.source, DeviceTable):
            desc.table = self.source.handle
        else:
            hfs, arr = _files_array(self.source)
            desc.files, desc.n_files = arr, len(hfs)
        names = [c.encode() for c in d.columns]
        cols = (L.PqColumn * max(1, len(names)))()
        for i, n in enumerate(names):
            cols[i].name = n
            cols[i].type = _pq_type(self.schema.get(d.columns[i]))
        desc.columns, desc.n_columns = cols, len(names)
        if ops:
            arr_ops = (L.PqPredOp * len(ops))(*ops)
            desc.pred, desc.n_pred = arr_ops, len(ops)
        if gb:
            arr_gb = (C.c_int32 * len(gb))(*gb)
            desc.group_by, desc.n_group_by = arr_gb, len(gb)
            if gx is not None:
                desc.group_exprs = gx
        if ag:
            arr_ag = (L.PqAgg * len(ag))(*ag)
            desc.aggs, desc.n_aggs = arr_ag, len(ag)
        if proj:
            arr_pj = (C.c_int32 * len(proj))(*proj)
            desc.projection, desc.n_projection = arr_pj, len(proj)
        desc.limit = -1 if limit is None else int(limit)
        desc.batch_size = batch_size
        desc.shard_index, desc.shard_count = self.shard_index, self.shard_count
        desc.flags = flags

        h = C.c_void_p()
        rc = lib.pq_query_open(C.byref(desc), C.byref(h))
        if rc != L.PQ_OK:
            raise QueryError(rc, (lib.pq_last_error(None) or b"").decode())
        try:
            # every batch through ONE Arrow C stream (pq_query_stream): what arrow-rs does with
            # ArrowArrayStreamReader; pq_query_next stays for consumers that poll batch by batch
            batches = []
            while poll:   # poll_next, one batch per call
                arr_c, sch_c = L.ArrowArray(), L.ArrowSchema()
                rc = lib.pq_query_next(h, 0, C.byref(arr_c), C.byref(sch_c))
                if rc == L.PQ_END_OF_STREAM:
                    break
                if rc != L.PQ_OK:
                    raise QueryError(rc, (lib.pq_last_error(h) or b"").decode())
                batches.append(pa.RecordBatch._import_from_c(C.addressof(arr_c), C.addressof(sch_c)))
            if not poll:
                stream_c = L.ArrowArrayStream()
                rc = lib.pq_query_stream(h, 0, C.byref(stream_c))
                if rc != L.PQ_OK:
                    raise QueryError(rc, (lib.pq_last_error(h) or b"").decode())
                try:
                    batches = list(pa.RecordBatchReader._import_from_c(C.addressof(stream_c)))
                except pa.ArrowException as e:
                    raise QueryError(L.PQ_ERR_CUDA, (lib.pq_last_error(h) or str(e).encode()).decode()) from e
            json_text = None
            if json is not None:      # the same result as JSON text, formatted on the GPU
                jp, jn = C.c_void_p(), C.c_uint64()
                rc = lib.pq_query_json(h, L.PQ_JSON_LINES if json == "lines" else 0, C.byref(jp), C.byref(jn))
                if rc != L.PQ_OK:
                    raise QueryError(rc, (lib.pq_last_error(h) or b"").decode())
                json_text = C.string_at(jp.value, jn.value) if jn.value else b""
            m = L.PqMetrics()
            lib.pq_query_metrics(h, C.byref(m))
        finally:
            lib.pq_query_close(h)
        return QueryResult(batches, m.as_dict(), [f.name for f in batches[0].schema] if batches else [], json_text)


def flatten_objects_for_count(objects: list[dict]) -> list[dict]:
    """src/query/mod.rs:858-902 (kept "for later" by the reference; its six unit tests pin it): JSON rows that all carry
    the one same ``COUNT...`` key -- per-partition / per-node COUNT results -- fold into one row with their sum; anything
    else passes through untouched.  On the GPU path the same fold is the all-reduce of the partial tables."""
    if not objects:
        return objects
    first_key = next(iter(objects[0]), None)
    if all(all(k.startswith("COUNT") for k in o) for o in objects) and all(all(k == first_key for k in o) for o in objects):
        return [{first_key: sum(int(v) for o in objects for v in o.values())}]
    return objects


# ----------------------------------------------------------------------------- Query / execute
@dataclass
class TimeRange:
    start_ms: int
    end_ms: int


class Query:
    """``SELECT <cols | aggs> FROM <stream> [WHERE ...] [GROUP BY ...] [LIMIT n]`` — the subset of SQL the
    GPU path executes.  The reference hands SQL to DataFusion's planner (src/query/mod.rs:261-264);
    that planner is out of scope, so this small recursive-descent parser only exists
    to let tests and the bench state their queries the way Parseable users do."""

    def __init__(self, sql: str, time_range: TimeRange | None = None):
        self.sql = sql
        self.time_range = time_range
        self._parse(sql)

    # --- tokenizer / parser ---
    _TOK = re.compile(r"\s*(?:(-?\d+\.\d+(?:[eE][-+]?\d+)?|-?\d+)|'((?:[^']|'')*)'|\"([^\"]+)\"|([A-Za-z_][A-Za-z_0-9]*)|(<=|>=|<>|!=|[=<>(),*]))")

    def _parse(self, sql: str):
        toks, pos = [], 0
        sql = sql.strip().rstrip(";")
        while pos < len(sql):
            m = self._TOK.match(sql, pos)
            if not m:
                raise QueryError(L.PQ_ERR_INVALID_ARG, f"cannot tokenise SQL at: {sql[pos:pos+20]!r}")
            pos = m.end()
            if m.group(1) is not None:
                toks.append(("num", m.group(1)))
            elif m.group(2) is not None:
                toks.append(("str", m.group(2).replace("''", "'")))
            elif m.group(3) is not None:
                toks.append(("id", m.group(3)))
            elif m.group(4) is not None:
                w = m.group(4)
                toks.append(("kw", w.upper()) if w.upper() in _KEYWORDS else ("id", w))
            else:
                toks.append(("op", m.group(5)))
        self._t, self._i = toks, 0
        self._expect("kw", "SELECT")
        self.select: list = []
        while True:
            self.select.append(self._select_item())
            if not self._accept("op", ","):
                break
        self._expect("kw", "FROM")
        self.stream = self._next("id")[1]
        self.where = None
        self.group_by: list[str] = []
        self.limit = None
        if self._accept("kw", "WHERE"):
            self.where = self._or()
        if self._accept("kw", "GROUP"):
            self._expect("kw", "BY")
            while True:
                self.group_by.append(self._next("id")[1])
                if not self._accept("op", ","):
                    break
        if self._accept("kw", "LIMIT"):
            self.limit = int(self._next("num")[1])
        if self._i != len(self._t):
            raise QueryError(L.PQ_ERR_UNSUPPORTED, f"unsupported SQL near {self._t[self._i]!r}")

    def _peek(self):
        return self._t[self._i] if self._i < len(self._t) else (None, None)

    def _next(self, kind):
        t = self._peek()
        if t[0] != kind:
            raise QueryError(L.PQ_ERR_INVALID_ARG, f"expected {kind}, found {t!r}")
        self._i += 1
        return t

    def _accept(self, kind, val):
        t = self._peek()
        if t[0] == kind and t[1] == val:
            self._i += 1
            return True
        return False

    def _expect(self, kind, val):
        if not self._accept(kind, val):
            raise QueryError(L.PQ_ERR_INVALID_ARG, f"expected {val}, found {self._peek()!r}")

    def _select_item(self):
        t = self._peek()
        if t == ("op", "*"):
            self._i += 1
            return ("star",)
        if t[0] == "kw" and t[1] in ("COUNT", "SUM", "MIN", "MAX", "AVG"):
            self._i += 1
            self._expect("op", "(")
            if t[1] == "COUNT" and self._accept("op", "*"):
                item = Agg("count_star")
            else:
                item = Agg(t[1].lower(), self._next("id")[1])
            self._expect("op", ")")
            alias = self._next("id")[1] if self._accept("kw", "AS") else None
            return ("agg", item, alias)
        name = self._next("id")[1]
        alias = self._next("id")[1] if self._accept("kw", "AS") else None
        return ("col", name, alias)

    def _or(self):
        e = self._and()
        while self._accept("kw", "OR"):
            e = e | self._and()
        return e

    def _and(self):
        e = self._not()
        while self._accept("kw", "AND"):
            e = e & self._not()
        return e

    def _not(self):
        if self._accept("kw", "NOT"):
            return ~self._not()
        return self._primary()

    def _value(self):
        t = self._peek()
        self._i += 1
        if t[0] == "num":
            return lit(float(t[1]) if any(c in t[1] for c in ".eE") else int(t[1]))
        if t[0] == "str":
            return lit(t[1])
        if t[0] == "id":
            return col(t[1])
        if t == ("kw", "TRUE"):
            return lit(True)
        if t == ("kw", "FALSE"):
            return lit(False)
        if t == ("kw", "NULL"):
            return lit(None)
        raise QueryError(L.PQ_ERR_INVALID_ARG, f"unexpected token {t!r}")

    def _primary(self):
        if self._accept("op", "("):
            e = self._or()
            self._expect("op", ")")
            return e
        a = self._value()
        t = self._peek()
        if t[0] == "op" and t[1] in _CMP:
            self._i += 1
            b = self._value()
            return Expr("cmp", (a, b), _CMP[t[1]])
        if self._accept("kw", "IS"):
            neg = self._accept("kw", "NOT")
            self._expect("kw", "NULL")
            return a.is_not_null() if neg else a.is_null()
        neg = self._accept("kw", "NOT")
        if t := self._peek():
            if t == ("kw", "LIKE") or t == ("kw", "ILIKE"):
                self._i += 1
                p = self._next("str")[1]
                if self._accept("kw", "ESCAPE"):
                    esc = self._next("str")[1]
                    if esc != "\\":     # the matcher's escape character is the backslash (arrow-string's default)
                        raise QueryError(L.PQ_ERR_UNSUPPORTED, f"LIKE ... ESCAPE {esc!r}: only the backslash is supported")
                return a.like(p, negated=neg, case_insensitive=(t[1] == "ILIKE"))
        raise QueryError(L.PQ_ERR_UNSUPPORTED, f"unsupported predicate near {self._peek()!r}")

    # --- src/query/mod.rs:774-856: wrap the scan in the time-range filter unless the user
    #     already filtered on the time column ---
    def final_filters(self) -> list[Expr]:
        filters = [] if self.where is None else [self.where]
        if self.time_range is not None and not _mentions(self.where, DEFAULT_TIMESTAMP_KEY):
            filters.append(col(DEFAULT_TIMESTAMP_KEY) >= Timestamp(self.time_range.start_ms))
            filters.append(col(DEFAULT_TIMESTAMP_KEY) < Timestamp(self.time_range.end_ms))
        return filters


_KEYWORDS = {"SELECT", "FROM", "WHERE", "GROUP", "BY", "AND", "OR", "NOT", "LIKE", "ILIKE", "IS", "NULL", "COUNT",
             "SUM", "MIN", "MAX", "AVG", "AS", "LIMIT", "TRUE", "FALSE", "ESCAPE"}
_CMP = {"=": L.PQ_EQ, "!=": L.PQ_NE, "<>": L.PQ_NE, "<": L.PQ_LT, "<=": L.PQ_LE, ">": L.PQ_GT, ">=": L.PQ_GE}


def _mentions(e: Expr | None, name: str) -> bool:
    if e is None:
        return False
    if e.kind == "col":
        return e.args[0] == name
    return any(_mentions(a, name) for a in e.args if isinstance(a, Expr))


def execute(query: Query, provider: StandardTableProvider, is_streaming: bool = False) -> QueryResult:
    """query::execute (src/query/mod.rs:143-149).  ``is_streaming`` only changes how the
    reference hands batches over (Vec vs stream); the batches are the same."""
    aggs = [it[1] for it in query.select if it[0] == "agg"]
    cols = [it[1] for it in query.select if it[0] == "col"]
    star = any(it[0] == "star" for it in query.select)
    filters = query.final_filters()
    if aggs:
        if star:
            raise QueryError(L.PQ_ERR_INVALID_ARG, "SELECT * next to an aggregate")
        extra = [c for c in cols if c not in query.group_by]
        if extra:
            raise QueryError(L.PQ_ERR_INVALID_ARG, f"column {extra[0]} must appear in GROUP BY")
        res = provider.aggregate(query.group_by, aggs, filters)
        # output columns in SELECT order under their aliases; LIMIT applies to the groups
        if res.batches:
            t = res.table()
            names, picked = [], []
            for it in query.select:
                if it[0] == "col":
                    src = it[1]
                else:
                    a = it[1]
                    src = "count(*)" if a.fn == "count_star" else f"{a.fn}({a.column})"
                picked.append(t.column(src))
                names.append(it[2] or src)
            t = pa.table(picked, names=names)
            if query.limit is not None:
                t = t.slice(0, query.limit)
            res.batches = t.to_batches(max_chunksize=20000) or res.batches[:1]
            res.fields = names
        return res
    if query.group_by:
        raise QueryError(L.PQ_ERR_UNSUPPORTED, "GROUP BY without aggregates")
    if star:
        if not provider.schema:
            raise QueryError(L.PQ_ERR_INVALID_ARG, "SELECT * needs the table schema")
        cols = list(provider.schema.keys())
    res = provider.scan(cols, filters, query.limit)
    aliases = {it[1]: it[2] for it in query.select if it[0] == "col" and it[2]}
    if aliases and res.batches:
        names = [aliases.get(n, n) for n in res.batches[0].schema.names]
        res.batches = [b.rename_columns(names) for b in res.batches]
        res.fields = names
    return res
