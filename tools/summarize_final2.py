#!/usr/bin/env python
"""Build profiles/r1_final2_* from gpurun_out/i_* (second evidence pass of round 1: k_cg_step2d, pinned e2e path)."""
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_final as sf  # noqa: E402

OUT, PROF = sf.OUT, sf.PROF

if __name__ == "__main__":
    cmd = "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
    sf.launches("i", cmd, "r1_final2_cfg2_launches_summary.csv")
    res = sf.full("i", "k_cg_step2d", "r1_final2_cfg2_k_cg_step2d_full.csv")
    for src, dst in [("i_bench_default.json", "r1_final2_bench_default_cfg2.json"), ("i_plain.json", "r1_final2_bench_plain_steps2_cfg2.json"),
                     ("i_pytest.log", "r1_final2_pytest_gpu.log"), ("i_multi2d.log", "r1_final2_multigpu2_2d_cases.log"),
                     ("h_step2d_probe.log", "r1_final2_step2d_probe1.log"), ("h_step2d_probe2.log", "r1_final2_step2d_probe2.log"),
                     ("h_step2d_probe3.log", "r1_final2_step2d_probe3.log"), ("h_h2d_probe.log", "r1_final2_h2d_probe_before.log"),
                     ("h_h2d_probe2.log", "r1_final2_h2d_probe_after.log")]:
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    # DRAM traffic per launch of the kernels that run now: new capture for the 2-D CG kernels, the earlier capture for
    # the kernels that did not change
    old = json.load(open(os.path.join(PROF, "r1_final_ncu_traffic.json")))
    new = {"cfg2": {k: v for k, v in old["cfg2"].items() if k in ("cg_update", "zu")}}
    for r in res:
        u = r["units"]
        def val(key):
            v = float(r[key])
            unit = u[key]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
        dur = float(r["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}[u["gpu__time_duration.sum"]]
        cls = "cg_prec" if r["kernel"].rstrip().endswith(", 2>") else "cg_step"
        new["cfg2"][cls] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "ncu_duration_s": dur,
                            "kernel": r["kernel"], "launches_captured": 1}
    json.dump(new, open(os.path.join(PROF, "r1_final2_ncu_traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, "r1_final2_cfg2_launches_summary.csv")).read())
    print(open(os.path.join(PROF, "r1_final2_cfg2_k_cg_step2d_full.csv")).read())
    print(open(os.path.join(PROF, "r1_final2_cfg2_k_cg_step2d_stalls.txt")).read())
    print(json.dumps(new, indent=1))
