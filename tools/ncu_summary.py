"""Summarise an .ncu-rep (read on the CPU box): headline metrics, stall reasons, hottest SASS segments."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']
for k in keys:
    if k in d: print(f"{k:70s} {d[k]} {u[k]}")
print("-- stall reasons (warp-cycles per issued instruction)")
st = [(float(v), h) for h, v in d.items() if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and v not in ('', 'n/a')]
for v, h in sorted(st, reverse=True)[:8]:
    print(f"  {v:8.3f} {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
data = [dict(zip(h2, r)) for r in rows[2:] if len(r) == len(h2)]
def I(x):
    try: return int(x)
    except: return 0
tot = sum(I(x['Instructions Executed']) for x in data); ts = sum(I(x['# Samples']) for x in data)
print(f"-- SASS: {len(data)} instructions, {tot} warp-instructions executed, {ts} samples")
segs = []; start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or not (0.8 <= (I(data[i]['Instructions Executed']) + 1) / (I(data[start]['Instructions Executed']) + 1) <= 1.25):
        segs.append((sum(I(x['Instructions Executed']) for x in data[start:i]), sum(I(x['# Samples']) for x in data[start:i]), start, i)); start = i
print("-- hottest segments by executed instructions")
for s, sm, a, b in sorted(segs, reverse=True)[:10]:
    ops = {}
    for x in data[a:b]:
        t = x['Source'].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op] = ops.get(op, 0) + 1
    print(f"  {100*s/tot:5.1f}% inst {100*sm/max(ts,1):5.1f}% samples  [{a}:{b}] n={b-a} per-inst={I(data[a]['Instructions Executed'])} {sorted(ops.items(), key=lambda x: -x[1])[:5]}")
print("-- hottest segments by stall samples")
for s, sm, a, b in sorted(segs, key=lambda x: -x[1])[:10]:
    ops = {}
    for x in data[a:b]:
        t = x['Source'].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op] = ops.get(op, 0) + 1
    print(f"  {100*sm/max(ts,1):5.1f}% samples {100*s/tot:5.1f}% inst  [{a}:{b}] n={b-a} per-inst={I(data[a]['Instructions Executed'])} {sorted(ops.items(), key=lambda x: -x[1])[:5]}")
