"""Per-source-line attribution of an .ncu-rep (read on the CPU box): executed warp-instructions and
stall samples per CUDA source line, hottest first.  Needs a capture made with --import-source on and
code compiled with -lineinfo.  Usage: python tools/ncu_lines.py <report.ncu-rep> [units] [top]
`units` (e.g. the number of slabs of the launch) scales the instruction counts to "per unit"."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":     # rows with a line number are per-source-line aggregates
            try:
                n, s = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            if n or s:
                out.append((n, s, cur, r[0], r[1][:110]))
    tot, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print(f"total {tot} warp-instructions, {ts} samples" + (f", {tot / units:.0f} per unit" if units else ""))
    print("-- by executed instructions")
    for n, s, f, l, src in sorted(out, reverse=True)[:top]:
        per = f" {n / units:8.0f}/unit" if units else ""
        print(f"{100 * n / tot:5.1f}% inst {100 * s / ts:5.1f}% samples{per}  {f}:{l}  {src}")
    print("-- by stall samples")
    for n, s, f, l, src in sorted(out, key=lambda o: -o[1])[:top // 2]:
        print(f"{100 * n / tot:5.1f}% inst {100 * s / ts:5.1f}% samples  {f}:{l}  {src}")


if __name__ == "__main__":
    main()
