"""Re-creates, value for value, the fixtures of the reference's only known-answer tests at the
DataFusion boundary: /root/reference/src/storage/field_stats.rs:775-924 (10-row table),
:1258-1296 (1000 rows / 10 categories) and :1195-1212 (empty table), and writes the counts those
tests assert (:927-1066, :1069-1095, :1312-1327) to expected.json.

Run once here (pyarrow is the writer; the reference uses parquet-rs' ArrowWriter with default
properties, streams produce the same logical pages):  python tests/golden/make_field_stats_fixture.py
"""
import json
import os

import pyarrow as pa
import pyarrow.parquet as pq

HERE = os.path.dirname(os.path.abspath(__file__))

schema = pa.schema([
    pa.field("id", pa.int64(), False),
    pa.field("name", pa.string(), True),
    pa.field("score", pa.float64(), True),
    pa.field("active", pa.bool_(), True),
    pa.field("created_at", pa.timestamp("ms"), True),
    pa.field("single_value", pa.string(), True),
    pa.field("int_list", pa.list_(pa.field("item", pa.int64(), False)), True),
    pa.field("float_list", pa.list_(pa.field("item", pa.float64(), False)), True),
])

ten = pa.table({
    "id": pa.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], pa.int64()),
    "name": pa.array(["Alice", "Bob", "Alice", "Charlie", "Alice", "Bob", "David", None, "Eve", "Frank"]),
    "score": pa.array([95.5, 87.2, 95.5, 78.9, 92.1, 88.8, 91.0, None, 89.5, 94.2]),
    "active": pa.array([True, False, True, True, True, False, True, None, False, True]),
    "created_at": pa.array([1640995200000, 1640995260000, 1640995200000, 1640995320000, 1640995380000,
                            1640995440000, 1640995500000, None, 1640995560000, 1640995620000], pa.timestamp("ms")),
    "single_value": pa.array(["constant"] * 10),
    "int_list": pa.array([[1, 2, 3], [4, 5], [1, 2, 3], [6, 7, 8, 9], [1], [4, 5], [10, 11], [], [12, 13, 14], [1, 2]],
                         pa.list_(pa.field("item", pa.int64(), False))),
    "float_list": pa.array([[1.1, 2.2], [3.3, 4.4, 5.5], [1.1, 2.2], [6.6], [7.7, 8.8, 9.9], [3.3, 4.4, 5.5], [10.0],
                            [], [11.1, 12.2], [13.3]], pa.list_(pa.field("item", pa.float64(), False))),
}, schema=schema)
pq.write_table(ten, os.path.join(HERE, "field_stats_10rows.parquet"), compression="NONE")

cat_schema = pa.schema([pa.field("id", pa.int64(), False), pa.field("category", pa.string(), True)])
cats = pa.table({"id": pa.array(range(1000), pa.int64()),
                 "category": pa.array([f"cat_{i % 10}" for i in range(1000)])}, schema=cat_schema)
pq.write_table(cats, os.path.join(HERE, "field_stats_1000rows.parquet"), compression="NONE")

pq.write_table(schema.empty_table(), os.path.join(HERE, "field_stats_empty.parquet"), compression="NONE")

# what the reference tests assert (field_stats.rs line numbers in the keys' comments)
expected = {
    "ten_rows": {
        "name": {"count": 10, "distinct_count": 7, "counts": {"Alice": 3, "Bob": 2, "Charlie": 1}},   # :927-975
        "score": {"count": 10, "distinct_count": 9, "top": {"value": 95.5, "count": 2}},                # :978-1006
        "active": {"count": 10, "distinct_count": 3, "counts": {"true": 6, "false": 3}},                # :1009-1037
        "created_at": {"count": 10, "distinct_count": 9, "top_count": 2},                               # :1040-1066
        "single_value": {"count": 10, "distinct_count": 1, "counts": {"constant": 10}},                 # :1069-1095
    },
    "thousand_rows": {"category": {"count": 1000, "distinct_count": 10, "each": 100}},                  # :1312-1327
    "empty": {"name": {"groups": 0}},                                                                   # :1195-1212
}
with open(os.path.join(HERE, "expected.json"), "w") as f:
    json.dump(expected, f, indent=1, sort_keys=True)
print("written")
