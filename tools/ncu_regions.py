#!/usr/bin/env python
"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) by source line and by
code region: share of stall samples, instructions and shared-memory wavefronts.  Usage: ncu_regions.py page.csv frames"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except (ValueError, TypeError):
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    cur, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            data.append((cur, int(r[0]), num(d.get("# Samples")), num(d.get("Instructions Executed")),
                         num(d.get("L1 Wavefronts Shared")), r[1].strip()[:100]))
    tot = sum(d[2] for d in data) or 1
    toti = sum(d[3] for d in data) or 1
    totw = sum(d[4] for d in data) or 1
    print(f"total samples {tot}  inst/frame {toti / frames:.0f}  smem wavefronts/frame {totw / frames:.0f}")
    top = sorted(data, key=lambda d: -d[2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]
    for f, l, s, i, w, src in top:
        print(f"{100 * s / tot:5.1f}% smp {i / frames:7.1f} inst {w / frames:6.1f} wf  {f}:{l}  {src}")


if __name__ == "__main__":
    main()
