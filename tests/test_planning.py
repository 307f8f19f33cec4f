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
