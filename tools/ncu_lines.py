#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep captured with --import-source on.
usage: tools/ncu_lines.py rep.ncu-rep [top] [inst|smp] [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg",
          "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_dispatch",
          "stall_membar", "stall_tex", "stall_sleep", "stall_drain", "stall_misc"]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    key = sys.argv[3] if len(sys.argv) > 3 else "smp"
    only = int(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    agg = collections.OrderedDict()
    cur = None
    kidx = -1
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            kidx += 1
            ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
            si = [hdr.index(s) for s in STALLS]
            continue
        if hdr is None or len(r) <= ie or (only is not None and kidx != only):
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:90])
            agg.setdefault(cur, [0, 0, [0] * len(STALLS)])
        elif r[2].startswith("0x") and cur is not None:
            try:
                agg[cur][0] += int(r[ie]); agg[cur][1] += int(r[sm])
                for q, i in enumerate(si):
                    agg[cur][2][q] += int(r[i] or 0)
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    allst = [sum(v[2][q] for v in agg.values()) for q in range(len(STALLS))]
    print("stall mix: " + ", ".join(f"{s[6:]} {a / max(sum(allst), 1) * 100:.1f}%" for s, a in
                                   sorted(zip(STALLS, allst), key=lambda x: -x[1])[:8]))
    k = 1 if key == "smp" else 0
    for kk, v in sorted(agg.items(), key=lambda kv: -kv[1][k])[:top]:
        best = sorted(zip(STALLS, v[2]), key=lambda x: -x[1])[:2]
        why = " ".join(f"{s[6:]}:{c}" for s, c in best if c)
        print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% smp | L{kk[0]:>4} | {kk[1]:90s} | {why}")


if __name__ == "__main__":
    main()
