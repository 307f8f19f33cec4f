"""Scan-planning mirror (parseable_b200/planning.py) against the reference's own unit-test vectors:
/root/reference/src/query/stream_schema_provider.rs:1104-1286 (manifest overlap, extract_timestamp_bound) and
/root/reference/src/catalog/column.rs:306-445 (TypedStatistics::update), plus the rules of can_be_pruned /
satisfy_constraints (:955-1043), collect_from_snapshot (:478-508), partitioned_files (:351-446) and
supports_filters_pushdown (:665-683, 866-882)."""
import datetime as dt
import math

from parseable_b200 import _lib as L
from parseable_b200.planning import (EXACT, INEXACT, ManifestColumn, ManifestFileEntry, ManifestItem, PartialTimeFilter, TimestampNs,
                                     TypedStatistics, can_be_pruned, collect_from_snapshot, expr_in_boundary, extract_primary_filter,
                                     extract_timestamp_bound, final_time_filters, is_overlapping_query, is_within_staging_window,
                                     partitioned_files, snapshot_manifests, supports_filters_pushdown)
from parseable_b200.query import Timestamp, col


def _items():
    day = lambda d: (dt.datetime(2023, 12, d), dt.datetime(2023, 12, d, 23, 59, 59, 999000))   # noqa: E731
    return [ManifestItem(str(i + 1), *day(15 + i)) for i in range(3)]


def test_is_overlapping_query_reference_vectors():
    low = lambda t: [PartialTimeFilter("low", t, True)]   # noqa: E731
    assert is_overlapping_query(_items(), low(dt.datetime(2023, 12, 14)))                       # bound_min_is_overlapping
    assert is_overlapping_query(_items(), low(dt.datetime(2023, 12, 14, 3)))                    # bound_min_plus_hour_is_overlapping
    assert not is_overlapping_query(_items(), low(dt.datetime(2023, 12, 16)))                   # bound_next_day_min_is_not_overlapping
    assert is_overlapping_query([], low(dt.datetime(2023, 12, 16)))                             # no manifests at all: list


def test_extract_timestamp_bound_reference_vectors():
    t0 = dt.datetime(2023, 1, 1)
    tp = "timestamp_column"
    assert extract_timestamp_bound(col(tp) == Timestamp(1672531200000), tp) == (L.PQ_EQ, t0)                 # timestamp_in_milliseconds
    assert extract_timestamp_bound(col(tp) > TimestampNs(1672531200000000000), tp) == (L.PQ_GT, t0)           # timestamp_in_nanoseconds
    assert extract_timestamp_bound(col(tp) < "2023-01-01T00:00:00", tp) == (L.PQ_LT, t0)                      # string_timestamp
    assert extract_timestamp_bound(col("other_column") == "2023-01-01T00:00:00", tp) is None                  # unexpected_utf8_column
    assert extract_timestamp_bound(col(tp) == 42, tp) is None                                                 # unsupported_literal_type
    assert extract_timestamp_bound(col(tp) == col("other_column"), tp) is None                                # no_literal_on_right
    assert extract_timestamp_bound(col(tp) == Timestamp(1672531200000), None) == (L.PQ_EQ, t0)                # non_time_partition_timestamps
    assert extract_timestamp_bound(col(tp) == TimestampNs(1672531200000000000), None) == (L.PQ_EQ, t0)


def test_supports_filters_pushdown():
    minute = Timestamp(1_700_000_040_000)        # 2023-11-14 22:14:00.000
    off = Timestamp(1_700_000_040_500)
    assert expr_in_boundary(col("p_timestamp") >= minute) and expr_in_boundary(col("p_timestamp") < minute)
    assert not expr_in_boundary(col("p_timestamp") >= off)             # not minute aligned
    assert not expr_in_boundary(col("p_timestamp") == minute)          # only < <= > >=
    assert not expr_in_boundary((col("p_timestamp") >= minute) & (col("level") == "ERROR"))
    assert supports_filters_pushdown([col("p_timestamp") >= minute, col("level") == "ERROR", col("p_timestamp") < off]) == [EXACT, INEXACT, INEXACT]


def test_time_filters_and_manifest_pruning():
    f = [(col("level") == "ERROR") & (col("p_timestamp") >= Timestamp(int(dt.datetime(2023, 12, 16, 5).replace(tzinfo=dt.timezone.utc).timestamp() * 1000))),
         col("p_timestamp") < Timestamp(int(dt.datetime(2023, 12, 17).replace(tzinfo=dt.timezone.utc).timestamp() * 1000))]
    tf = extract_primary_filter(f)
    assert [(t.kind, t.included) for t in tf] == [("low", True), ("high", False)]
    assert [m.manifest_path for m in snapshot_manifests(_items(), tf)] == ["2"]
    assert [m.manifest_path for m in snapshot_manifests(_items(), [PartialTimeFilter("eq", dt.datetime(2023, 12, 17, 12))])] == ["3"]
    assert [m.manifest_path for m in snapshot_manifests(_items(), [PartialTimeFilter("high", dt.datetime(2023, 12, 16), True)])] == ["1", "2"]
    assert [m.manifest_path for m in snapshot_manifests(_items(), [PartialTimeFilter("high", dt.datetime(2023, 12, 16), False)])] == ["1"]
    now = dt.datetime(2024, 1, 1, 12, 0, 30)
    assert is_within_staging_window([PartialTimeFilter("high", dt.datetime(2024, 1, 1, 11, 58), False)], now)
    assert not is_within_staging_window([PartialTimeFilter("high", dt.datetime(2024, 1, 1, 11, 54), False)], now)
    assert is_within_staging_window([PartialTimeFilter("low", dt.datetime(2024, 1, 1, 1), True)], now)          # no upper bound
    ff = final_time_filters([col("level") == "ERROR"], 10, 20)
    assert len(ff) == 3 and final_time_filters([col("p_timestamp") > Timestamp(5)], 10, 20) == [ff2 for ff2 in final_time_filters([col("p_timestamp") > Timestamp(5)], 10, 20)]
    assert len(final_time_filters([col("p_timestamp") > Timestamp(5)], 10, 20)) == 1                            # the user filtered on time already


def _file(path, rows, **stats):
    return ManifestFileEntry(path, rows, 0, [ManifestColumn(k, v) for k, v in stats.items()])


def test_can_be_pruned_satisfy_constraints():
    f = _file("a", 100, status=TypedStatistics("int", 200, 404), cpu=TypedStatistics("float", 0.1, 0.9), host=TypedStatistics("string", "h-10", "h-50"),
              ok=TypedStatistics("bool", False, False), bare=None)
    P = lambda e: can_be_pruned(f, e)   # noqa: E731
    assert not P(col("status") == 200) and not P(col("status") == 404) and P(col("status") == 500) and P(col("status") == 199)
    assert P(col("status") < 200) and not P(col("status") <= 200) and not P(col("status") < 201)
    assert P(col("status") > 404) and not P(col("status") >= 404) and P(col("status") >= 405)
    assert not P(col("status") != 500)                                  # != never prunes
    assert P(col("cpu") > 0.9) and not P(col("cpu") > 0.5) and P(col("cpu") == 1.5)
    assert not P(col("cpu") == 1)                                       # Int literal vs Float statistics: cannot tell, keep the file
    assert P(col("host") == "h-60") and not P(col("host") == "h-20") and P(col("host") < "h-10")
    assert P(col("ok") == True) and not P(col("ok") == False)          # noqa: E712
    assert not P(col("bare") == 1) and not P(col("missing") == 1)       # no statistics / no such column: keep
    assert not P((col("status") == 500) & (col("cpu") > 2.0))            # only a bare comparison prunes
    assert P(col("status") == Timestamp(500))                           # TimestampMillisecond casts to Int


def test_collect_from_snapshot_and_partitioned_files():
    m1 = [_file("old1", 50, v=TypedStatistics("int", 0, 9)), _file("old2", 60, v=TypedStatistics("int", 10, 19))]
    m2 = [_file("new1", 70, v=TypedStatistics("int", 20, 29)), _file("new2", 80, v=TypedStatistics("int", 5, 25))]
    files = collect_from_snapshot([m1, m2], [])
    assert [f.file_path for f in files] == ["new2", "new1", "old2", "old1"]                    # newest first
    assert [f.file_path for f in collect_from_snapshot([m1, m2], [col("v") >= 20])] == ["new2", "new1"]
    assert [f.file_path for f in collect_from_snapshot([m1, m2], [], limit=100)] == ["new2", "new1"]  # 80 + 70 >= 100
    assert [f.file_path for f in collect_from_snapshot([m1, m2], [], limit=10_000)] == ["new2", "new1", "old2", "old1"]
    parts, stats, rows = partitioned_files(files, 3)
    assert [[f.file_path for f in p] for p in parts] == [["new2", "old1"], ["new1"], ["old2"]] and rows == 260
    assert (stats["v"].min, stats["v"].max) == (0, 29)


def test_typed_statistics_update_reference_vectors():
    S = TypedStatistics
    m = S("int", 5, 10).update(S("int", 1, 7))
    assert (m.min, m.max) == (1, 10)                                                          # update_merges_compatible_int_stats
    m = S("string", "b", "y").update(S("string", "a", "z"))
    assert (m.min, m.max) == ("a", "z")                                                       # update_merges_compatible_string_stats
    assert S("string", "2025-01-01", "2025-12-31").update(S("int", 1_700_000_000_000, 1_800_000_000_000)) is None   # type mismatch, both ways
    assert S("int", 1_700_000_000_000, 1_800_000_000_000).update(S("string", "2025-01-01", "2025-12-31")) is None
    b, i, f, s = S("bool", False, True), S("int", 0, 1), S("float", 0.0, 1.0), S("string", "a", "b")
    assert b.update(i) is None and i.update(f) is None and f.update(s) is None and s.update(b) is None
    inv = S("float", 600025.1656670001, 600025.165667)
    assert inv.update(S("float", 600025.1656670001, 600025.165667)) is None                  # both inverted
    assert inv.update(S("float", 100.0, 1_000_000.0)) is None                                 # one inverted, even if bracketed
    nan = S("float", math.nan, 10.0)
    assert nan.update(S("float", 1.0, 20.0)) is None and S("float", 1.0, 20.0).update(nan) is None
    _, stats, _ = partitioned_files([_file("a", 1, c=S("int", 1, 2)), _file("b", 1, c=S("string", "x", "y"))], 2)
    assert stats["c"] is None                                                                 # the planner then skips min/max pushdown


# ---- the same rules on the C side (csrc/planning.cpp, the pq_plan_* entry points of include/parseable_b200.h) ----
import ctypes as C   # noqa: E402

import pytest   # noqa: E402

EPOCH = dt.datetime(1970, 1, 1)


def _ns(t: dt.datetime) -> int:
    d = t - EPOCH
    return (d.days * 86400 + d.seconds) * 1_000_000_000 + d.microseconds * 1000


@pytest.fixture(scope="module")
def lib(built):
    return L.load()


class _Keep(list):
    """keeps the byte strings the C structs point to alive"""


def _c_literal(v, keep):
    lit = L.PqLiteral()
    if isinstance(v, bool):
        lit.type, lit.i64 = L.PQ_T_BOOL, int(v)
    elif isinstance(v, Timestamp):
        lit.type, lit.i64 = L.PQ_T_TS_MS, v.ms
    elif isinstance(v, TimestampNs):
        lit.type, lit.i64 = L.PQ_T_TS_NS, v.ns
    elif isinstance(v, int):
        lit.type, lit.i64 = L.PQ_T_I64, v
    elif isinstance(v, float):
        lit.type, lit.f64 = L.PQ_T_F64, v
    elif isinstance(v, str):
        b = v.encode()
        keep.append(b)
        lit.type, lit.str, lit.str_len = L.PQ_T_UTF8, b, len(b)
    else:
        lit.type = L.PQ_T_NULL
    return lit


def _c_filters(filters, keep):
    arr = (L.PqPlanFilter * max(len(filters), 1))()
    for i, e in enumerate(filters):
        simple = e.kind == "cmp" and e.args[0].kind == "col" and e.args[1].kind == "lit"
        if simple:
            name = e.args[0].args[0].encode()
            keep.append(name)
            arr[i].column, arr[i].cmp, arr[i].lit = name, e.op, _c_literal(e.args[1].args[0], keep)
        else:
            arr[i].column = None
    return arr


def _c_stat(name, st, keep):
    cs = L.PqColumnStat()
    nb = name.encode()
    keep.append(nb)
    cs.column = nb
    if st is None:
        cs.kind = L.PQ_STAT_NONE
    elif st.kind in ("bool", "int"):
        cs.kind, cs.min_i, cs.max_i = (L.PQ_STAT_BOOL if st.kind == "bool" else L.PQ_STAT_INT), int(st.min), int(st.max)
    elif st.kind == "float":
        cs.kind, cs.min_f, cs.max_f = L.PQ_STAT_FLOAT, st.min, st.max
    else:
        lo, hi = st.min.encode(), st.max.encode()
        keep += [lo, hi]
        cs.kind, cs.min_s, cs.min_s_len, cs.max_s, cs.max_s_len = L.PQ_STAT_STRING, lo, len(lo), hi, len(hi)
    return cs


def _c_files(files, keep):
    arr = (L.PqManifestFile * max(len(files), 1))()
    for i, f in enumerate(files):
        stats = (L.PqColumnStat * max(len(f.columns), 1))(*[_c_stat(c.name, c.stats, keep) for c in f.columns])
        keep.append(stats)
        pb = f.file_path.encode()
        keep.append(pb)
        arr[i].path, arr[i].num_rows, arr[i].file_size, arr[i].stats, arr[i].n_stats = pb, f.num_rows, f.file_size, stats, len(f.columns)
    return arr


def _c_bounds(tf):
    arr = (L.PqTimeBound * max(len(tf), 1))()
    for i, t in enumerate(tf):
        arr[i].kind = {"low": L.PQ_BOUND_LOW, "high": L.PQ_BOUND_HIGH, "eq": L.PQ_BOUND_EQ}[t.kind]
        arr[i].included, arr[i].time_ns = int(t.included), _ns(t.time)
    return arr


def _c_items(items):
    arr = (L.PqManifestItem * max(len(items), 1))()
    for i, m in enumerate(items):
        arr[i].time_lower_ns, arr[i].time_upper_ns = _ns(m.time_lower_bound), _ns(m.time_upper_bound)
    return arr


def _c_time_bounds(lib, filters, tp):
    keep = _Keep()
    out = (L.PqTimeBound * max(len(filters), 1))()
    n = lib.pq_plan_time_bounds(_c_filters(filters, keep), len(filters), tp.encode() if tp else None, out)
    assert n >= 0
    kinds = {L.PQ_BOUND_LOW: "low", L.PQ_BOUND_HIGH: "high", L.PQ_BOUND_EQ: "eq"}
    return [(kinds[out[i].kind], bool(out[i].included), out[i].time_ns) for i in range(n)]


def test_c_time_bounds_reference_vectors(lib):
    t0 = _ns(dt.datetime(2023, 1, 1))
    tp = "timestamp_column"
    assert _c_time_bounds(lib, [col(tp) == Timestamp(1672531200000)], tp) == [("eq", True, t0)]
    assert _c_time_bounds(lib, [col(tp) > TimestampNs(1672531200000000000)], tp) == [("low", False, t0)]
    assert _c_time_bounds(lib, [col(tp) < "2023-01-01T00:00:00"], tp) == [("high", False, t0)]
    assert _c_time_bounds(lib, [col(tp) <= "2023-01-01T00:00:00.250"], tp) == [("high", True, t0 + 250_000_000)]
    assert _c_time_bounds(lib, [col("other_column") == "2023-01-01T00:00:00"], tp) == []
    assert _c_time_bounds(lib, [col(tp) == 42], tp) == []
    assert _c_time_bounds(lib, [col(tp) == "not a time"], tp) == []
    assert _c_time_bounds(lib, [col(tp) == col("other_column")], tp) == []
    assert _c_time_bounds(lib, [col(tp) == Timestamp(1672531200000)], None) == [("eq", True, t0)]
    assert _c_time_bounds(lib, [col(tp) != Timestamp(1672531200000)], None) == []
    # the mirror and the C side agree on a mixed list
    fl = [col("p_timestamp") >= Timestamp(1_700_000_040_000), col("level") == "ERROR", col("p_timestamp") < Timestamp(1_700_000_940_500)]
    py = [(t.kind, t.included, _ns(t.time)) for t in extract_primary_filter(fl)]
    assert _c_time_bounds(lib, fl, None) == py


def test_c_manifests_overlap_staging(lib):
    items = _items()
    ci = _c_items(items)

    def kept(tf):
        keep = (C.c_uint8 * len(items))()
        assert lib.pq_plan_manifests(ci, len(items), _c_bounds(tf), len(tf), keep) == 0
        got = [m.manifest_path for m, k in zip(items, keep) if k]
        assert got == [m.manifest_path for m in snapshot_manifests(items, tf)]
        return got
    assert kept([PartialTimeFilter("low", dt.datetime(2023, 12, 16, 5), True), PartialTimeFilter("high", dt.datetime(2023, 12, 17), False)]) == ["2"]
    assert kept([PartialTimeFilter("eq", dt.datetime(2023, 12, 17, 12))]) == ["3"]
    assert kept([PartialTimeFilter("high", dt.datetime(2023, 12, 16), True)]) == ["1", "2"]
    assert kept([PartialTimeFilter("high", dt.datetime(2023, 12, 16), False)]) == ["1"]
    assert kept([]) == ["1", "2", "3"]
    low = lambda t: [PartialTimeFilter("low", t, True)]   # noqa: E731
    ov = lambda it, tf: lib.pq_plan_is_overlapping_query(_c_items(it), len(it), _c_bounds(tf), len(tf))   # noqa: E731
    assert ov(items, low(dt.datetime(2023, 12, 14))) == 1 and ov(items, low(dt.datetime(2023, 12, 14, 3))) == 1   # reference vectors
    assert ov(items, low(dt.datetime(2023, 12, 16))) == 0 and ov([], low(dt.datetime(2023, 12, 16))) == 1
    now = dt.datetime(2024, 1, 1, 12, 0, 30)
    st = lambda tf: lib.pq_plan_within_staging_window(_c_bounds(tf), len(tf), _ns(now))   # noqa: E731
    for tf in ([PartialTimeFilter("high", dt.datetime(2024, 1, 1, 11, 58), False)], [PartialTimeFilter("high", dt.datetime(2024, 1, 1, 11, 54), False)],
               [PartialTimeFilter("high", dt.datetime(2024, 1, 1, 11, 55), False)], [PartialTimeFilter("low", dt.datetime(2024, 1, 1, 1), True)],
               [PartialTimeFilter("eq", dt.datetime(2024, 1, 1, 11, 59))], []):
        assert bool(st(tf)) == is_within_staging_window(tf, now), tf


def test_c_pruning_limit_merge_pushdown(lib):
    f = _file("a", 100, status=TypedStatistics("int", 200, 404), cpu=TypedStatistics("float", 0.1, 0.9), host=TypedStatistics("string", "h-10", "h-50"),
              ok=TypedStatistics("bool", False, False), bare=None)

    def P(e):
        keep = _Keep()
        out = (C.c_uint32 * 1)()
        n = lib.pq_plan_collect_files(_c_files([f], keep), 1, _c_filters([e], keep), 1, -1, out)
        assert n in (0, 1)
        assert (n == 0) == can_be_pruned(f, e), e          # the mirror agrees
        return n == 0
    assert not P(col("status") == 200) and not P(col("status") == 404) and P(col("status") == 500) and P(col("status") == 199)
    assert P(col("status") < 200) and not P(col("status") <= 200) and not P(col("status") < 201)
    assert P(col("status") > 404) and not P(col("status") >= 404) and P(col("status") >= 405)
    assert not P(col("status") != 500)
    assert P(col("cpu") > 0.9) and not P(col("cpu") > 0.5) and P(col("cpu") == 1.5) and not P(col("cpu") == 1)
    assert P(col("cpu") == math.nan)
    assert P(col("host") == "h-60") and not P(col("host") == "h-20") and P(col("host") < "h-10") and not P(col("host") >= "h-5")
    assert P(col("ok") == True) and not P(col("ok") == False)          # noqa: E712
    assert not P(col("bare") == 1) and not P(col("missing") == 1)
    assert not P((col("status") == 500) & (col("cpu") > 2.0))
    assert P(col("status") == Timestamp(500))
    m1 = [_file("old1", 50, v=TypedStatistics("int", 0, 9)), _file("old2", 60, v=TypedStatistics("int", 10, 19))]
    m2 = [_file("new1", 70, v=TypedStatistics("int", 20, 29)), _file("new2", 80, v=TypedStatistics("int", 5, 25))]
    flat = m1 + m2

    def collect(filters, limit):
        keep = _Keep()
        out = (C.c_uint32 * len(flat))()
        n = lib.pq_plan_collect_files(_c_files(flat, keep), len(flat), _c_filters(filters, keep), len(filters), -1 if limit is None else limit, out)
        got = [flat[out[i]].file_path for i in range(n)]
        assert got == [x.file_path for x in collect_from_snapshot([m1, m2], filters, limit)]
        return got
    assert collect([], None) == ["new2", "new1", "old2", "old1"]
    assert collect([col("v") >= 20], None) == ["new2", "new1"]
    assert collect([], 100) == ["new2", "new1"] and collect([], 10_000) == ["new2", "new1", "old2", "old1"]
    assert collect([col("v") >= 20, col("v") < 6], None) == ["new2"] and collect([col("v") > 100], 5) == []
    # TypedStatistics::update (src/catalog/column.rs:306-445)
    S = TypedStatistics

    def merge(a, b):
        keep = _Keep()
        ca, cb, out = _c_stat("c", a, keep), _c_stat("c", b, keep), L.PqColumnStat()
        ok = lib.pq_plan_merge_stat(C.byref(ca), C.byref(cb), C.byref(out))
        py = a.update(b)
        assert bool(ok) == (py is not None)
        if not ok:
            return None
        if a.kind in ("int", "bool"):
            got = (out.min_i, out.max_i)
        elif a.kind == "float":
            got = (out.min_f, out.max_f)
        else:
            got = (C.string_at(out.min_s, out.min_s_len).decode(), C.string_at(out.max_s, out.max_s_len).decode())
        assert got == (py.min, py.max)
        return got
    assert merge(S("int", 5, 10), S("int", 1, 7)) == (1, 10)
    assert merge(S("string", "b", "y"), S("string", "a", "z")) == ("a", "z")
    assert merge(S("string", "2025-01-01", "2025-12-31"), S("int", 1_700_000_000_000, 1_800_000_000_000)) is None
    assert merge(S("float", 0.5, 2.0), S("float", -1.0, 1.0)) == (-1.0, 2.0)
    inv = S("float", 600025.1656670001, 600025.165667)
    assert merge(inv, S("float", 600025.1656670001, 600025.165667)) is None and merge(inv, S("float", 100.0, 1_000_000.0)) is None
    assert merge(S("float", math.nan, 10.0), S("float", 1.0, 20.0)) is None and merge(S("float", 1.0, 20.0), S("float", math.nan, 10.0)) is None
    # supports_filters_pushdown
    minute, off = Timestamp(1_700_000_040_000), Timestamp(1_700_000_040_500)
    fl = [col("p_timestamp") >= minute, col("level") == "ERROR", col("p_timestamp") < off, col("p_timestamp") == minute,
          (col("p_timestamp") >= minute) & (col("level") == "ERROR"), col("p_timestamp") <= TimestampNs(1_700_000_040_000_000_000)]
    keep = _Keep()
    ex = (C.c_uint8 * len(fl))()
    assert lib.pq_plan_pushdown(_c_filters(fl, keep), len(fl), ex) == 0
    assert [EXACT if e else INEXACT for e in ex] == supports_filters_pushdown(fl) == [EXACT, INEXACT, INEXACT, INEXACT, INEXACT, EXACT]
