"""Development probe: where the time of an e2e (host buffers -> result) step goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from parseable_b200.query import HostFile, StandardTableProvider

def main():
    nrg = int(sys.argv[1]) if len(sys.argv) > 1 else 480
    files = bench.ensure_data(nrg)
    hfs = [HostFile(path=p, pinned=True) for p in files]
    prov = StandardTableProvider(hfs, schema=bench.schema())
    keys, aggs = bench.c4_query()
    tf, _ = bench.time_filters(nrg)
    for _ in range(2):
        prov.aggregate(keys, aggs, tf)
    os.environ["PQB_VERBOSE"] = "2"
    t = time.perf_counter()
    r = prov.aggregate(keys, aggs, tf)
    print(f"step {1e3 * (time.perf_counter() - t):.2f} ms upload_ms {r.metrics['upload_ms']:.2f} device {r.metrics['device_ms']:.2f} h2d {r.metrics['h2d_bytes'] / 1e6:.0f} MB", flush=True)
    os.environ["PQB_VERBOSE"] = "0"
    for _ in range(3):
        t = time.perf_counter()
        r = prov.aggregate(keys, aggs, tf)
        print(f"step {1e3 * (time.perf_counter() - t):.2f} ms upload_ms {r.metrics['upload_ms']:.2f}", flush=True)

if __name__ == "__main__":
    main()
