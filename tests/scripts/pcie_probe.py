"""Development probe: what one big pinned H2D copy achieves on this box (the bound of the e2e step's upload).
torch is used for the pinned buffer and the events only.

    python tests/scripts/pcie_probe.py [MiB=2048]
"""
import sys

import torch


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    n = mib << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(7)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for chunk_mib in (mib, 64, 8, 1):
        c = chunk_mib << 20
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for o in range(0, n, c):
                d[o:o + c].copy_(h[o:o + c], non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"H2D {mib} MiB in {chunk_mib} MiB copies: {best:.2f} ms = {n / best / 1e6:.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
