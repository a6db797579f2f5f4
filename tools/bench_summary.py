import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print("%s: %.2f ms/pass  %.3g vtx-upd/s  J=%.1f  pass_frac=%.3f  launches=%d" % (f.split('/')[-1], d["ms_per_step"], d["value"], d["inner_cg_iters_per_pass"], d.get("pass_kernel_frac_of_peak", d["pass_frac_of_peak"]), d["gpu_launches"]))
    for k, v in d["stages"].items():
        print("   %-10s avg %.4f ms  %5.0f GB/s  frac %.3f  total %.1f ms" % (k, v["avg_ms"], v["gbs"], v["frac"], v["total_ms"]))
    if d.get("e2e"): print("   e2e %.3g" % d["e2e"]["value"])
