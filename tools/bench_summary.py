import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([ln for ln in open(f) if ln.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    ps = (d.get("roofline") or {}).get("pass", {})
    print("%s: N=%d %.2f ms/pass  %.3g vtx-upd/s  J=%.1f deg=%s  pass 8(d) frac: %.3f of measured, %.3f of 8 TB/s (kernel bytes %.3f)  launches=%d parity=%s" % (
        f.split('/')[-1], d["n_gpus"], d["ms_per_step"], d["value"], d["inner_cg_iters_per_pass"], d["config"].get("poly_degree"),
        ps.get("frac_of_measured_peak", 0), ps.get("frac_of_8TBs", 0), ps.get("kernel_frac_of_measured_peak", 0), d["gpu_launches"],
        (d.get("parity") or {}).get("ok")))
    for k, v in d["stages"].items():
        print("   %-10s avg %.4f ms  %5.0f GB/s  frac %.3f  total %.1f ms" % (k, v["avg_ms"], v["gbs"], v["frac"], v["total_ms"]))
    for blk in ("classical_cg", "cfg2", "cfg4"):
        if blk in d:
            b = d[blk]
            print("   [%s] %.2f ms/pass J=%.1f deg=%s pass frac %.3f / %.3f of 8TB/s" % (blk, b["ms_per_step"], b["inner_cg_iters_per_pass"],
                  b["config"]["poly_degree"], b.get("pass", {}).get("frac_of_measured_peak", 0), b.get("pass", {}).get("frac_of_8TBs", 0)))
    if d.get("e2e"): print("   e2e %.3g" % d["e2e"]["value"])
