import os, sys
sys.path.insert(0, os.getcwd())
from parseable_b200 import synth
from parseable_b200.query import *
import numpy as np, pyarrow as pa
p = "/tmp/dbg_nn.parquet"
synth.write_logs16(p, n_row_groups=3, rows_per_group=70_000)
from oracle.oracle import Oracle
ora = Oracle.from_parquet(p)
prov = StandardTableProvider([p], schema=ora.table.schema)
ts = ora.table["p_timestamp"].drop_null().cast(pa.int64()).to_numpy()
lo, hi = int(np.quantile(ts, 0.2)), int(np.quantile(ts, 0.7))
rng_f = [col("p_timestamp") >= Timestamp(lo), col("p_timestamp") < Timestamp(hi)]
qs = [([date_bin("1m")], [count_star(), sum_("bytes"), max_("cpu")], []),
      ([date_bin("5m"), "level"], [count_star(), sum_("bytes"), max_("cpu")], [col("status") == 200]),
      ([date_bin(7000), "status", "region"], [count_star(), sum_("bytes"), max_("cpu")], [col("latency_ms") > 50]),
      ([date_bin("1m")], [count_star()], rng_f)]
for i, (keys, aggs, flt) in enumerate(qs):
    os.environ["PQB_VERBOSE"] = "1" if len(sys.argv) > 1 else "0"
    try:
        r = prov.aggregate(keys, aggs, flt)
        print(i, "ok", r.table().num_rows, flush=True)
    except Exception as e:
        print(i, "ERR", e, flush=True)
