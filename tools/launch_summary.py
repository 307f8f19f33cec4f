"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> markdown table.

    python tools/launch_summary.py gpurun_out/launches.csv "title line" > profiles/launches_summary.md
"""
import collections
import csv
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r:
        if len(row) <= vi:
            continue
        name = row[ki].split("(")[0].replace("void ", "").split("<")[0].strip()
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row[ui].strip(), 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {title}\n")
    print("Cold-cache, serialised launches: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for name, v in tot.most_common():
        print(f"| `{name}` | {cnt[name]} | {v:.1f} | {v / cnt[name]:.1f} | {100 * v / total:.1f} % |")
    print(f"\n{sum(cnt.values())} launches, {total / 1e3:.1f} ms of kernel time.")


if __name__ == "__main__":
    main()
