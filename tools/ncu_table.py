#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics ... --csv --log-file` capture (one row per launch)."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
by = collections.OrderedDict()
for r in rows:
    key = (r["ID"], r["Kernel Name"])
    by.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
def num(x):
    try:
        return float(x[0].replace(",", ""))
    except Exception:
        return float("nan")
def to(x, unit_map):
    return num(x) * unit_map.get(x[1], 1.0)
BY = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "second": 1e3}
S = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
stalls = ["long_scoreboard", "barrier", "short_scoreboard", "math_pipe_throttle", "wait", "lg_throttle", "mio_throttle",
          "membar", "no_instruction", "not_selected", "dispatch_stall", "branch_resolving", "sleeping"]
print(f"{'kernel':40s} {'grid':>6s} {'reg':>3s} {'ms':>7s} {'rdGB':>6s} {'wrGB':>6s} {'L2GB':>6s} {'dram%':>5s} {'L2hit':>5s} {'L1hit':>5s} {'occ%':>5s} {'iss%':>5s} {'fp64':>5s} {'Minst':>7s} {'lclGB':>6s} | " + " ".join(s[:6] for s in stalls))
tot = collections.Counter()
for (i, k), m in by.items():
    name = k.split("(")[0].replace("void ", "").replace("kamr::", "")
    g = lambda key, d=float("nan"): num(m[key]) if key in m else d
    ms = to(m["gpu__time_duration.sum"], T)
    rd = to(m["dram__bytes_read.sum"], BY); wr = to(m["dram__bytes_write.sum"], BY)
    l2 = to(m["lts__t_bytes.sum"], BY) if "lts__t_bytes.sum" in m else 0
    lcl = (g("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", 0) + g("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", 0)) * 32e-9
    st = " ".join("%6.2f" % g(S % s) for s in stalls)
    print(f"{name[:40]:40s} {g('launch__grid_size'):6.0f} {g('launch__registers_per_thread'):3.0f} {ms:7.3f} {rd:6.2f} {wr:6.2f} {l2:6.1f} "
          f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {g('lts__t_sector_hit_rate.pct'):5.1f} {g('l1tex__t_sector_hit_rate.pct'):5.1f} "
          f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} {g('smsp__inst_executed.sum')/1e6:7.1f} {lcl:6.2f} | {st}")
    tot["ms"] += ms; tot["rd"] += rd; tot["wr"] += wr; tot["l2"] += l2
print(f"{'TOTAL':40s} {'':6s} {'':3s} {tot['ms']:7.3f} {tot['rd']:6.2f} {tot['wr']:6.2f} {tot['l2']:6.1f}")
