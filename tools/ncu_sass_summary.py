#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: instruction mix by opcode and the
hottest instructions by stall samples.  Usage: ncu_sass_summary.py sass.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {name: i for i, name in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    inst = lambda r: float(r[col["Instructions Executed"]] or 0)
    thr = lambda r: float(r[col["Thread Instructions Executed"]] or 0)
    smp = lambda r: float(r[col["# Samples"]] or 0)
    total_i, total_t, total_s = sum(map(inst, body)), sum(map(thr, body)), sum(map(smp, body))
    print(f"{len(body)} SASS instructions; executed {total_i:.3e} warp-inst, {total_t:.3e} thread-inst "
          f"(lane eff {total_t / max(1, total_i) / 32:.2f}); {total_s:.0f} samples")
    by_op = defaultdict(lambda: [0.0, 0.0, 0])
    for r in body:
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.split(".")[0]
        by_op[op][0] += inst(r)
        by_op[op][1] += smp(r)
        by_op[op][2] += 1
    print("\nopcode                 warp-inst   share   samples share  static")
    for op, (i, s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{op:20s} {i:12.3e}  {i / total_i:6.3f}  {s:8.0f} {s / max(1, total_s):6.3f}  {n:5d}")
    print(f"\ntop {top} instructions by stall samples")
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    for r in sorted(body, key=lambda r: -smp(r))[:top]:
        stalls = sorted(((float(r[col[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        st = ",".join(f"{c[6:]}={v:.0f}" for v, c in stalls if v > 0)
        print(f"{r[col['Address']][-5:]} {smp(r):7.0f} {inst(r):11.3e} {float(r[col['Avg. Threads Executed']] or 0):5.1f}  {r[col['Source']].strip()[:90]:90s} {st}")


if __name__ == "__main__":
    main()
