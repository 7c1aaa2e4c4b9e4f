#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep captured with --import-source on: share of stall samples and of executed
instructions per line of the kernel (runs here, no GPU).   python tools/ncu_lines.py rep.ncu-rep [launch_index] [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    # one block per launch, each starting with a "File Path" row
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "File Path":
            cur = []
            blocks.append(cur)
        if cur is not None:
            cur.append(r)
    rows = blocks[idx]
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iL, iS, iI = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    agg, src = defaultdict(lambda: [0, 0]), {}
    for r in rows:
        if len(r) < len(hdr):
            continue
        try:
            ln, s, ins = int(r[iL]), int(r[iS] or 0), int(r[iI] or 0)
        except ValueError:
            continue
        agg[ln][0] += s
        agg[ln][1] += ins
        src[ln] = r[1][:100]
    ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    print(f"{rows[1][1] if len(rows) > 1 else ''}: samples {ts}, warp instructions {ti}")
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{ln:4d} samp {100 * v[0] / max(ts, 1):5.1f}% instr {100 * v[1] / max(ti, 1):5.1f}% | {src[ln]}")


if __name__ == "__main__":
    main()
