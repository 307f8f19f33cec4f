"""The oracle against (1) the reference's own known-answer fixtures (field_stats.rs, re-created in
tests/golden/) and (2) an independent engine, pyarrow/Acero, on everything the reference's tests do
not pin (SUM/MIN/MAX, comparisons, LIKE, Kleene logic)."""
import json
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle.oracle import Oracle
from parseable_b200.query import (avg, col, count, count_star, lit, max_, min_, sum_)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EXP = json.load(open(os.path.join(GOLD, "expected.json")))


def field_stats(ora, field):
    """The shape of get_stats_sql (src/storage/field_stats.rs:298-330): GROUP BY field, COUNT(*)."""
    t = ora.group_by([field], [count_star()])
    vals = t[field].to_pylist()
    cnts = t["count(*)"].to_pylist()
    return dict(zip(vals, cnts))


def test_golden_ten_rows():
    ora = Oracle.from_parquet(os.path.join(GOLD, "field_stats_10rows.parquet"),
                              columns=["id", "name", "score", "active", "created_at", "single_value"])
    e = EXP["ten_rows"]
    name = field_stats(ora, "name")
    assert sum(name.values()) == e["name"]["count"] and len(name) == e["name"]["distinct_count"]
    for k, v in e["name"]["counts"].items():
        assert name[k] == v
    assert name[None] == 1                       # NULL forms its own group
    score = field_stats(ora, "score")
    assert sum(score.values()) == 10 and len(score) == e["score"]["distinct_count"]
    assert score[e["score"]["top"]["value"]] == e["score"]["top"]["count"] == max(score.values())
    active = field_stats(ora, "active")
    assert len(active) == 3 and active[True] == 6 and active[False] == 3 and active[None] == 1
    created = field_stats(ora, "created_at")
    assert len(created) == e["created_at"]["distinct_count"] and max(created.values()) == 2
    single = field_stats(ora, "single_value")
    assert single == {"constant": 10}


def test_golden_thousand_rows_and_empty():
    ora = Oracle.from_parquet(os.path.join(GOLD, "field_stats_1000rows.parquet"))
    cat = field_stats(ora, "category")
    assert len(cat) == 10 and all(v == 100 for v in cat.values()) and sum(cat.values()) == 1000
    empty = Oracle.from_parquet(os.path.join(GOLD, "field_stats_empty.parquet"), columns=["name"])
    assert field_stats(empty, "name") == {}


@pytest.fixture(scope="module")
def ora(small_files):
    return Oracle.from_parquet(small_files["nulls"])


def _acero_mask(tb, expr):
    return pc.fill_null(expr, False)


def test_filters_agree_with_acero(ora):
    tb = ora.table
    lvl = tb["level"].cast(pa.string())
    cases = [
        ([(col("level") == "ERROR") & (col("latency_ms") > 100)],
         pc.and_kleene(pc.equal(lvl, "ERROR"), pc.greater(tb["latency_ms"], 100))),
        ([(col("level") == "FATAL") | (col("bytes") < 1000)],
         pc.or_kleene(pc.equal(lvl, "FATAL"), pc.less(tb["bytes"], 1000))),
        ([~((col("status") == 200) | (col("cpu") >= 0.25))],
         pc.invert(pc.or_kleene(pc.equal(tb["status"], 200), pc.greater_equal(tb["cpu"], 0.25)))),
        ([col("host").is_null()], pc.is_null(tb["host"])),
        ([col("message").like("%timeout-xyzzy%")], pc.match_like(tb["message"].cast(pa.string()), "%timeout-xyzzy%")),
        ([col("path").like("/api/v1/resource/00_1")], pc.match_like(tb["path"].cast(pa.string()), "/api/v1/resource/00_1")),
        ([col("host") >= "host-05000"], pc.greater_equal(tb["host"].cast(pa.string()), "host-05000")),
    ]
    for flt, expr in cases:
        want = pc.sum(_acero_mask(tb, expr)).as_py() or 0
        assert ora.count(flt) == want


def test_like_and_ilike_patterns_agree_with_acero():
    """LIKE is pinned by no reference test: cross-check the oracle's matcher against
    Acero's match_like over wildcards, escapes ('\\' as in the reference's `ESCAPE '\\'`), empty strings,
    NULLs and case folding (ASCII: both sides fold ASCII only)."""
    vals = ["", "a", "ab", "abc", "a%c", "a_c", "A_C", "abcabc", "xxabcxx", "ABC", "aXc", "%", "_", "a\\c", "timeout", "Timeout after 30s",
            "GET /api/v1/users/42", "get /api/v1/users/42", None, "ééé", "aéc"]
    pats = ["%", "", "a", "a%", "%c", "%b%", "a_c", "a\\_c", "a\\%c", "%\\%%", "_", "__", "%abc%abc%", "abc%abc", "%a%b%c%", "A_C", "%timeout%",
            "Timeout%30s", "GET /api/%/users/__", "a%c", "%é%", "a_c%", "%_"]
    t = pa.table({"s": pa.array(vals, pa.string())})
    o = Oracle(t)
    for p in pats:
        for ci in (False, True):
            if ci and any(ord(ch) > 127 for ch in p):
                continue    # ILIKE folds ASCII only in this implementation; Acero folds Unicode
            want = pc.sum(pc.fill_null(pc.match_like(t["s"], p, ignore_case=ci), False)).as_py() or 0
            got = o.count([col("s").like(p, case_insensitive=ci)])
            assert got == want, (p, ci, got, want)
            want_not = pc.sum(pc.fill_null(pc.invert(pc.match_like(t["s"], p, ignore_case=ci)), False)).as_py() or 0
            assert o.count([col("s").like(p, negated=True, case_insensitive=ci)]) == want_not, (p, ci, "NOT")


def test_group_by_agrees_with_acero(ora):
    tb = ora.table
    tb = tb.set_column(tb.column_names.index("host"), "host", tb["host"].cast(pa.string()))
    got = ora.group_by(["host", "status"], [count_star(), sum_("bytes"), min_("latency_ms"), max_("latency_ms"),
                                            count("cpu"), max_("cpu"), avg("bytes")]).sort_by([("host", "ascending"), ("status", "ascending")])
    ref = tb.group_by(["host", "status"]).aggregate([([], "count_all"), ("bytes", "sum"), ("latency_ms", "min"),
                                                     ("latency_ms", "max"), ("cpu", "count"), ("cpu", "max"), ("bytes", "mean")])
    ref = ref.sort_by([("host", "ascending"), ("status", "ascending")])
    assert got["count(*)"].to_pylist() == ref["count_all"].to_pylist()
    assert got["sum(bytes)"].to_pylist() == ref["bytes_sum"].to_pylist()
    assert got["min(latency_ms)"].to_pylist() == ref["latency_ms_min"].to_pylist()
    assert got["max(latency_ms)"].to_pylist() == ref["latency_ms_max"].to_pylist()
    assert got["count(cpu)"].to_pylist() == ref["cpu_count"].to_pylist()
    assert got["max(cpu)"].to_pylist() == ref["cpu_max"].to_pylist()
    a = np.array(got["avg(bytes)"].to_pylist(), dtype=float)
    b = np.array(ref["bytes_mean"].to_pylist(), dtype=float)
    assert np.allclose(a, b, rtol=1e-12, equal_nan=True)


def test_documented_differences_from_acero():
    """Where DataFusion's rules differ from Acero's defaults the oracle follows DataFusion
    (SURVEY.md §8c): float compare is totalOrder, SUM(Int64) wraps."""
    nan = float("nan")
    t = pa.table({"x": pa.array([nan, 1.0, -0.0, 0.0, None]), "k": pa.array([1, 1, 1, 1, 1]),
                  "big": pa.array([2**62, 2**62, 2**62, 0, 0])})
    o = Oracle(t)
    assert o.count([col("x") == nan]) == 1                 # NaN == NaN under totalOrder
    assert o.count([col("x") > 1e308]) == 1                # NaN is the greatest value
    assert o.count([col("x") < 0.0]) == 1                  # -0.0 < +0.0
    assert o.count([col("x") == 0.0]) == 1
    g = o.group_by(["k"], [sum_("big"), max_("x"), min_("x"), count("x")])
    assert g["sum(big)"].to_pylist() == [(3 * 2**62) - 2**64]          # wrapped
    assert np.isnan(g["max(x)"].to_pylist()[0])
    assert str(g["min(x)"].to_pylist()[0]) == "-0.0"
    assert g["count(x)"].to_pylist() == [4]


def test_kleene_truth_table():
    t = pa.table({"a": pa.array([True, True, True, False, False, False, None, None, None]),
                  "b": pa.array([True, False, None, True, False, None, True, False, None])})
    o = Oracle(t)
    A, B = col("a") == True, col("b") == True  # noqa: E712
    assert o.select([A & B]).tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 0]
    assert o.select([A | B]).tolist() == [1, 1, 1, 1, 0, 0, 1, 0, 0]
    assert o.select([~(A & B)]).tolist() == [0, 1, 0, 1, 1, 1, 0, 1, 0]
    assert o.select([~(A | B)]).tolist() == [0, 0, 0, 0, 1, 0, 0, 0, 0]
    assert o.select([lit(None) | A]).tolist() == [1, 1, 1, 0, 0, 0, 0, 0, 0]
