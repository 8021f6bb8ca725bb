#!/usr/bin/env python
"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`.
Usage: ncu_line_summary.py cs.csv [top_n]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    out = []
    fname = "?"
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            col = {}
            for i, name in enumerate(r):
                col.setdefault(name, i)
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[col["Address"]] != "-":
            continue  # SASS row; the source-line row above already aggregates it
        try:
            inst = float(r[col["Instructions Executed"]] or 0)
            smp = float(r[col["# Samples"]] or 0)
            thr = float(r[col["Thread Instructions Executed"]] or 0)
        except ValueError:
            continue
        if inst > 0:
            out.append((inst, smp, thr, fname, r[0], r[1].strip()))
    ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
    print(f"total warp-inst {ti:.3e}, samples {ts:.0f}")
    print("   warp-inst  share  samples share lanes  file:line  source")
    for inst, smp, thr, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
        print(f"{inst:12.3e} {inst / ti:6.3f} {smp:8.0f} {smp / max(ts, 1):6.3f} {thr / inst:5.1f}  {f}:{ln}  {src[:100]}")


if __name__ == "__main__":
    main()
