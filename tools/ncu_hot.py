#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of one kernel in an .ncu-rep: ncu_hot.py file.ncu-rep [kernel-id like :::2] [min share]"""
import csv
import subprocess
import sys


def main(path, kid=None, share=0.012):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv"] + (["--kernel-id", kid] if kid else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    s0, src = hdr.index("# Samples"), hdr.index("Source")
    data = []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            data.append((int(r[s0]), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data)
    print("total samples", tot, "instructions", len(data))
    for k, (n, r) in enumerate(data):
        if n > tot * share:
            print(f"{k:5d} {n:6d} {100.0 * n / tot:5.1f}%  {r[src][:120]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, float(sys.argv[3]) if len(sys.argv) > 3 else 0.012)
