"""World-size-2 CPU (gloo) test of the multi-GPU PROTOCOL the library implements with NCCL:
row groups dealt g % n == rank, per-rank partial tables over rank-local key ids, all-gather of
the distinct key values, identical first-occurrence numbering in rank order, remap, dense
slot-aligned tables, one all-reduce per accumulator array (sum / min / max).  The per-rank
partials come from the CPU oracle (this is a test of the exchange, not of the kernels)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, path, q):
    import pyarrow as pa
    import pyarrow.parquet as pq
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle.oracle import Oracle
    from parseable_b200.query import count_star, max_, min_, sum_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pf = pq.ParquetFile(path)
    mine = [g for g in range(pf.metadata.num_row_groups) if g % world == rank]          # partitioned_files round-robin
    cols = ["host", "status", "bytes", "latency_ms"]
    tb = pa.concat_tables([pf.read_row_group(g, columns=cols) for g in mine]) if mine else None
    aggs = [count_star(), sum_("bytes"), min_("latency_ms"), max_("latency_ms")]
    part = Oracle(tb).group_by(["host", "status"], aggs) if tb is not None else None
    # all-gather distinct key values, number them in rank order (what Query::run does on the host)
    key_lists = [None] * world
    local_keys = {k: (part[k].to_pylist() if part is not None else []) for k in ("host", "status")}
    dist.all_gather_object(key_lists, {k: list(dict.fromkeys(v)) for k, v in local_keys.items()})
    ids = {}
    for k in ("host", "status"):
        ids[k] = {}
        for r in range(world):
            for v in key_lists[r][k]:
                ids[k].setdefault(v, len(ids[k]))
    card = {k: len(v) for k, v in ids.items()}
    nslots = (card["host"] + 1) * (card["status"] + 1)
    rows = np.zeros(nslots, np.int64)
    s_bytes = np.zeros(nslots, np.int64)
    mn = np.full(nslots, np.iinfo(np.int64).max)
    mx = np.full(nslots, np.iinfo(np.int64).min)
    if part is not None:
        for h, st, c, sb, a, b in zip(*[part[n].to_pylist() for n in part.column_names]):
            slot = ids["host"][h] + ids["status"][st] * (card["host"] + 1)
            rows[slot], s_bytes[slot], mn[slot], mx[slot] = c, sb, a, b
    t = [torch.from_numpy(x) for x in (rows, s_bytes, mn, mx)]
    dist.all_reduce(t[0], dist.ReduceOp.SUM)
    dist.all_reduce(t[1], dist.ReduceOp.SUM)
    dist.all_reduce(t[2], dist.ReduceOp.MIN)
    dist.all_reduce(t[3], dist.ReduceOp.MAX)
    inv_h = {v: k for k, v in ids["host"].items()}
    inv_s = {v: k for k, v in ids["status"].items()}
    out = {}
    for slot in np.flatnonzero(rows):
        out[(inv_h[slot % (card["host"] + 1)], inv_s[slot // (card["host"] + 1)])] = (int(rows[slot]), int(s_bytes[slot]), int(mn[slot]), int(mx[slot]))
    q.put((rank, out))
    dist.destroy_process_group()


def test_world2_partial_tables_allreduce(small_files):
    import multiprocessing as mp
    import socket
    from oracle.oracle import Oracle
    from parseable_b200.query import count_star, max_, min_, sum_
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path = small_files["nn"]
    ps = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=300) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    full = Oracle.from_parquet(path, columns=["host", "status", "bytes", "latency_ms"]).group_by(
        ["host", "status"], [count_star(), sum_("bytes"), min_("latency_ms"), max_("latency_ms")])
    want = {(h, st): (c, sb, a, b) for h, st, c, sb, a, b in zip(*[full[n].to_pylist() for n in full.column_names])}
    assert res[0] == want and res[1] == want
