#!/usr/bin/env python
"""Build profiles/r1_final_* from gpurun_out/f_* and f3_* (launch lists, ncu --set full exports, bench JSON)."""
import collections, csv, io, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum"]

def short(name):
    name = name.replace("void ", "").replace("mvtv::", "")
    return name.split("(")[0]

def launches(tag, cmd, out_name):
    lines = [ln for ln in open(os.path.join(OUT, tag + "_launches.csv")) if ln.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    s = io.StringIO()
    s.write("# every kernel launch of `%s` under\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n" % cmd)
    s.write("kernel,launches,total_ms,avg_ms,share_pct,grid,block\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        s.write('"%s",%d,%.4f,%.5f,%.2f,"%s","%s"\n' % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
    open(os.path.join(PROF, out_name), "w").write(s.getvalue())
    return agg, tot

def full(tag, kernel, out_name):
    path = os.path.join(OUT, "%s_%s_raw.csv" % (tag, kernel))
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in WANT]
    kcol = hdr.index("Kernel Name")
    s = io.StringIO()
    s.write("# %s: ncu --set full --clock-control none --import-source on, per launch (units in the second row)\n" % kernel)
    s.write(",".join(["kernel"] + [hdr[i] for i in idx]) + "\n")
    s.write(",".join(["-"] + [units[i] for i in idx]) + "\n")
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        s.write(",".join(['"%s"' % short(r[kcol])] + [r[i].replace(",", "") for i in idx]) + "\n")
        res.append({hdr[i]: r[i].replace(",", "") for i in idx} | {"kernel": short(r[kcol]), "units": {hdr[i]: units[i] for i in idx}})
    open(os.path.join(PROF, out_name), "w").write(s.getvalue())
    # top stall locations (source page)
    sp = os.path.join(OUT, "%s_%s_source.csv" % (tag, kernel))
    if os.path.exists(sp):
        rd = list(csv.reader(open(sp)))
        for i, r in enumerate(rd):
            if len(r) > 3 and r[0] == "Address":
                h, start = r, i + 1
                break
        si, sc = h.index("# Samples"), h.index("Source")
        rws = [(int(r[si]) if r[si].isdigit() else 0, r[sc].strip()) for r in rd[start:] if len(r) > si]
        tot = sum(x[0] for x in rws) or 1
        agg = collections.Counter()
        for n, t in rws:
            tk = t.split()
            op = tk[1] if tk and tk[0].startswith("@") and len(tk) > 1 else (tk[0] if tk else "?")
            agg[op.split(".")[0]] += n
        with open(os.path.join(PROF, out_name.replace("_full.csv", "_stalls.txt")), "w") as f:
            f.write("# %s: warp-state samples by SASS opcode (ncu source page), then the 12 hottest instructions\n" % kernel)
            f.write(", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in agg.most_common(12)) + "\n")
            for n, t in sorted(rws, reverse=True)[:12]:
                f.write("%6d %5.1f%%  %s\n" % (n, 100 * n / tot, t[:100]))
    return res

if __name__ == "__main__":
    os.makedirs(PROF, exist_ok=True)
    cmd = "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
    launches("f", cmd, "r1_final_cfg2_launches_summary.csv")
    launches("f3", cmd + " --workload cfg3", "r1_final_cfg3_launches_summary.csv")
    for k in ("k_cg_step", "k_cg_update", "k_zu_march"):
        full("f", k, "r1_final_cfg2_%s_full.csv" % k)
    for k in ("k_cg_step", "k_zu_march"):
        full("f3", k, "r1_final_cfg3_%s_full.csv" % k)
    for src, dst in [("f_bench_default.json", "r1_final_bench_default_cfg2.json"), ("f_bench_reference.json", "r1_final_bench_reference.json"),
                     ("f_bench_cfg3.json", "r1_final_bench_cfg3_512cubed.json"), ("f_bench_cfg4.json", "r1_final_bench_cfg4_96pow4.json"),
                     ("f_bench_cfg2_f32.json", "r1_final_bench_cfg2_f32.json"), ("f_bench_cfg2_jacobi.json", "r1_final_bench_cfg2_jacobi.json"),
                     ("f_plain.json", "r1_final_bench_plain_steps2_cfg2.json"), ("f3_plain.json", "r1_final_bench_plain_steps2_cfg3.json")]:
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    print(open(os.path.join(PROF, "r1_final_cfg2_launches_summary.csv")).read())
    print(open(os.path.join(PROF, "r1_final_cfg3_launches_summary.csv")).read())
