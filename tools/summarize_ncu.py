#!/usr/bin/env python
"""Turn gpurun_out/<tag>_launches.csv and <tag>_<kernel>.ncu-rep into the small text summaries committed
under profiles/.  Usage: python tools/summarize_ncu.py <tag> [kernel ...]"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def launches(tag, cmd):
    path = os.path.join(OUT, tag + "_launches.csv")
    lines = [ln for ln in open(path) if ln.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = io.StringIO()
    out.write("# %s: every kernel launch of `%s` under\n" % (tag, cmd))
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    out.write("kernel,launches,total_ms,avg_ms,share_pct,grid,block\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write('"%s",%d,%.4f,%.5f,%.2f,"%s","%s"\n' % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
    open(os.path.join(PROF, tag + "_launches_summary.csv"), "w").write(out.getvalue())
    print(out.getvalue())


def full(tag, kernel):
    rep = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, kernel))
    if not os.path.exists(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in WANT]
    out = io.StringIO()
    out.write("# %s %s: ncu --set full --clock-control none, per launch (units in the second row)\n" % (tag, kernel))
    out.write(",".join(["kernel"] + [hdr[i] for i in idx]) + "\n")
    out.write(",".join(["-"] + [units[i] for i in idx]) + "\n")
    kcol = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.write(",".join(['"%s"' % r[kcol].split("(")[0]] + [r[i].replace(",", "") for i in idx]) + "\n")
    open(os.path.join(PROF, "%s_%s_full.csv" % (tag, kernel)), "w").write(out.getvalue())
    print(out.getvalue())


if __name__ == "__main__":
    tag = sys.argv[1]
    cmd = os.environ.get("NCU_CMD", "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e")
    os.makedirs(PROF, exist_ok=True)
    launches(tag, cmd)
    for k in sys.argv[2:] or ["k_cg_step", "k_cg_update", "k_zu", "k_cg_init"]:
        full(tag, k)
