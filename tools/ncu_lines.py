#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep captured with --import-source on.
usage: tools/ncu_lines.py rep.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    agg = collections.OrderedDict()
    cur = None
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ie:
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:100])
            agg.setdefault(cur, [0, 0])
        elif r[2].startswith("0x") and cur is not None:
            try:
                agg[cur][0] += int(r[ie]); agg[cur][1] += int(r[sm])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% smp | L{k[0]:>4} | {k[1]}")


if __name__ == "__main__":
    main()
