"""One group-by query over the bench table, for ncu captures: python tools/profile_groupby.py [row_groups] [which]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyarrow as pa
import bench
from parseable_b200.query import *


def main():
    nrg = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    which = sys.argv[2] if len(sys.argv) > 2 else "small"
    files = bench.ensure_data(nrg)
    cols = ["host", "bytes", "level", "status"]
    schema = {"host": pa.string(), "bytes": pa.int64(), "level": pa.string(), "status": pa.int64()}
    table = DeviceTable(files, cols)
    prov = StandardTableProvider(table, schema=schema)
    q = {"small": (["level", "status"], [count_star(), sum_("bytes"), min_("bytes"), max_("bytes")], []),
         "c3": (["host"], [count_star(), sum_("bytes")], []),
         "filtered": (["host"], [count_star(), sum_("bytes")], [col("level") == "ERROR"])}[which]
    for _ in range(4):
        r = prov.aggregate(*q)
    print(which, "k_scan ms", r.metrics["scan_kernel_ms"], "groups", r.metrics["groups"])


if __name__ == "__main__":
    main()
