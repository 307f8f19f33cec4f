"""Host-side logic of the reference-facing mirror (parseable_b200/query.py), no GPU: the SQL subset,
the postfix predicate program handed to pq_query_open, literal typing, and the time-range filter the
reference injects on every scan (/root/reference/src/query/mod.rs:774-856)."""
import pyarrow as pa
import pytest

from parseable_b200 import _lib as L
from parseable_b200.query import (DEFAULT_TIMESTAMP_KEY, Agg, Query, QueryError, TimeRange, Timestamp, _Desc, _pq_type, col, lit)


def program(expr):
    d, ops = _Desc(), []
    d.compile_pred(expr, ops)
    return d, ops


def kinds(ops):
    return [o.kind for o in ops]


def test_predicate_compiles_to_postfix_with_columns_by_name():
    e = ((col("level") == "ERROR") | (col("level") == "FATAL")) & ~(col("status") >= 500)
    d, ops = program(e)
    assert kinds(ops) == [L.PQ_OP_CMP, L.PQ_OP_CMP, L.PQ_OP_OR, L.PQ_OP_CMP, L.PQ_OP_NOT, L.PQ_OP_AND]
    assert d.columns == ["level", "status"]                    # first-use order, referenced BY NAME in the descriptor
    assert [ops[i].col for i in (0, 1, 3)] == [0, 0, 1]
    assert ops[0].cmp == L.PQ_EQ and ops[3].cmp == L.PQ_GE
    assert ops[0].lit.type == L.PQ_T_UTF8 and ops[0].lit.str_len == 5 and ops[0].lit.str[:5] == b"ERROR"
    assert ops[3].lit.type == L.PQ_T_I64 and ops[3].lit.i64 == 500


def test_literal_on_the_left_is_flipped_and_typed():
    _, ops = program(Query("SELECT COUNT(*) FROM t WHERE 100 < latency_ms").where)
    assert kinds(ops) == [L.PQ_OP_CMP] and ops[0].cmp == L.PQ_GT and ops[0].lit.i64 == 100
    d = _Desc()
    assert d.literal(True).type == L.PQ_T_BOOL and d.literal(True).i64 == 1          # bool before int
    assert d.literal(1.5).type == L.PQ_T_F64 and d.literal(1.5).f64 == 1.5
    assert d.literal(None).type == L.PQ_T_NULL
    ts = d.literal(Timestamp(1_700_000_000_000))
    assert ts.type == L.PQ_T_TS_MS and ts.i64 == 1_700_000_000_000
    with pytest.raises(TypeError):
        d.literal(object())


def test_like_is_null_and_constants():
    _, ops = program(col("message").like("%timeout%", negated=True) & col("host").is_not_null() & lit(True))
    assert kinds(ops) == [L.PQ_OP_LIKE, L.PQ_OP_IS_NOT_NULL, L.PQ_OP_AND, L.PQ_OP_CONST, L.PQ_OP_AND]
    assert ops[0].flags == L.PQ_LIKE_NEGATED and ops[0].lit.str[:9] == b"%timeout%"
    _, ops = program(col("message").ilike("err%"))
    assert ops[0].flags == L.PQ_LIKE_CASE_INSENSITIVE


def test_column_to_column_comparison_is_refused_not_guessed():
    with pytest.raises(QueryError) as ei:
        program(Query("SELECT COUNT(*) FROM t WHERE a < b").where)
    assert ei.value.code == L.PQ_ERR_UNSUPPORTED


def test_sql_subset_parses_like_the_reference_queries():
    q = Query("SELECT host, COUNT(*), SUM(bytes) AS b, MIN(latency_ms) FROM demo "
              "WHERE level = 'ERROR' AND (status >= 500 OR message LIKE '%can''t%') AND cpu IS NOT NULL GROUP BY host LIMIT 10;")
    assert q.stream == "demo" and q.group_by == ["host"] and q.limit == 10
    assert [it[0] for it in q.select] == ["col", "agg", "agg", "agg"]
    assert q.select[1][1] == Agg("count_star") and q.select[2][1] == Agg("sum", "bytes") and q.select[3][1] == Agg("min", "latency_ms")
    d, ops = program(q.where)
    assert kinds(ops) == [L.PQ_OP_CMP, L.PQ_OP_CMP, L.PQ_OP_LIKE, L.PQ_OP_OR, L.PQ_OP_AND, L.PQ_OP_IS_NOT_NULL, L.PQ_OP_AND]
    assert ops[2].lit.str[:7] == b"%can't%"                      # '' unescapes to '
    assert d.columns == ["level", "status", "message", "cpu"]
    assert Query("select count(*) from t where x <> 1.5e3").where.args[1].args[0] == 1500.0
    assert Query('SELECT "weird col" FROM t').select == [("col", "weird col", None)]
    assert Query('SELECT a AS x, SUM(b) AS s FROM t WHERE c > -3 GROUP BY a').select[1][2] == "s"


@pytest.mark.parametrize("sql", ["SELECT a FROM t ORDER BY a", "SELECT a FROM t WHERE a BETWEEN 1 AND 2", "SELECT FROM t",
                                 "SELECT a FROM t WHERE a IN (1, 2)", "SELECT a t"])
def test_unsupported_sql_is_an_error(sql):
    with pytest.raises(QueryError) as ei:
        Query(sql)
    assert ei.value.code in (L.PQ_ERR_UNSUPPORTED, L.PQ_ERR_INVALID_ARG)


def test_time_range_is_injected_unless_the_user_filters_on_the_time_column():
    tr = TimeRange(1_000, 2_000)
    q = Query("SELECT COUNT(*) FROM demo WHERE status = 200", tr)
    f = q.final_filters()
    assert len(f) == 3
    _, ops = program(f[1])
    assert ops[0].cmp == L.PQ_GE and ops[0].lit.type == L.PQ_T_TS_MS and ops[0].lit.i64 == 1_000   # p_timestamp >= start
    _, ops = program(f[2])
    assert ops[0].cmp == L.PQ_LT and ops[0].lit.i64 == 2_000                                       # p_timestamp <  end
    q = Query(f"SELECT COUNT(*) FROM demo WHERE {DEFAULT_TIMESTAMP_KEY} >= 5 AND status = 200", tr)
    assert len(q.final_filters()) == 1                      # mod.rs:835-856: the user's own time filter wins
    assert Query("SELECT COUNT(*) FROM demo", tr).final_filters()[0].args[0].args[0] == DEFAULT_TIMESTAMP_KEY
    assert Query("SELECT COUNT(*) FROM demo").final_filters() == []


def test_arrow_types_map_to_abi_types():
    assert _pq_type(pa.timestamp("ms")) == L.PQ_T_TS_MS      # Parseable only writes Timestamp(ms) (src/utils/arrow/mod.rs:120-132)
    assert _pq_type(pa.dictionary(pa.int32(), pa.string())) == L.PQ_T_UTF8
    assert _pq_type(pa.float64()) == L.PQ_T_F64 and _pq_type(pa.bool_()) == L.PQ_T_BOOL and _pq_type(pa.int64()) == L.PQ_T_I64
    assert _pq_type(pa.list_(pa.int64())) == L.PQ_T_NULL and _pq_type(None) == L.PQ_T_NULL


def test_staging_batches_become_one_reversed_parquet_image():
    """reversed_mem_table (stream_schema_provider.rs:686-695): batches newest first, rows inside a batch reversed; the
    image is written the way the stream writes Parquet (DELTA_BINARY_PACKED time column, dictionaries elsewhere)."""
    import io

    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq

    from parseable_b200.query import _staging_image
    batches = [pa.table({"p_timestamp": pa.array(np.arange(5) + 10 * i, pa.timestamp("ms")),
                         "k": pa.array(["a", "b", None, "a", "c"])}).to_batches()[0] for i in range(3)]
    hf = _staging_image(batches)
    pf = pq.ParquetFile(io.BytesIO(hf._buf.raw[:hf.size]))
    t = pf.read()
    assert t["p_timestamp"].cast(pa.int64()).to_pylist() == [24, 23, 22, 21, 20, 14, 13, 12, 11, 10, 4, 3, 2, 1, 0]
    assert t["k"].to_pylist() == ["c", "a", None, "b", "a"] * 3
    md = pf.metadata.row_group(0)
    encs = {md.column(i).path_in_schema: set(md.column(i).encodings) for i in range(md.num_columns)}
    assert "DELTA_BINARY_PACKED" in encs["p_timestamp"] and "RLE_DICTIONARY" in encs["k"]


def test_flatten_objects_for_count_reference_vectors():
    """The six unit tests of /root/reference/src/query/mod.rs:1006-1090, value for value."""
    from parseable_b200.query import flatten_objects_for_count as f
    assert f([{"COUNT(*)": 1}, {"COUNT(*)": 2}, {"COUNT(*)": 3}]) == [{"COUNT(*)": 6}]                      # test_flat_simple
    assert f([]) == []                                                                                     # test_flat_empty
    assert f([{"COUNT(ALPHA)": 1}, {"COUNT(ALPHA)": 2}]) == [{"COUNT(ALPHA)": 3}]                           # test_flat_same_multi
    v = [{"COUNT(ALPHA)": 1}, {"COUNT(BETA)": 2}]
    assert f(list(v)) == v                                                                                 # test_flat_diff_multi
    v = [{"Num": 1}, {"Num": 2}, {"Num": 3}]
    assert f(list(v)) == v                                                                                 # test_flat_fail
    v = [{"Num": 1, "COUNT(*)": 1}, {"Num": 2, "COUNT(*)": 2}, {"Num": 3, "COUNT(*)": 3}]
    assert f(list(v)) == v                                                                                 # test_flat_multi_key
