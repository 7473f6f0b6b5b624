"""One-screen summary of a bench.py JSON line: python tools/bench_summary.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    d = json.load(open(f))
    print("==", f)
    print("value %.2f Mdof/s  %.1f ms/step  N=%d  iters %d  err %.2e  fail %s  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["n_gpus"], d["iterations"], d["rel_l2_vs_exact"], d["parity_failures"], d["gpu_launches"], d["clocks"]))
    r = d["roofline"]
    print("roofline: %.0f GB/s = %.3f of peak, %.4f ms/launch, share %.2f; CG iteration %.4f ms = %.0f GB/s" % (r["achieved"], r["frac"], r["avg_launch_ms"], r["share_of_step"], r["cg_iteration"]["ms"], r["cg_iteration"]["GBps"]))
    e = d.get("e2e")
    if e:
        print("e2e %.2f Mdof/s  %.1f ms  %s" % (e["value"], e["ms_per_step"], e["breakdown_last_step"]))
        print("e2e_cold %.1f ms %s" % (e["e2e_cold"]["ms"], e["e2e_cold"]["breakdown"]))
    if d.get("cpu_baseline"):
        print("cpu_baseline %.4f Mdof/s on %d cores" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
    k = d.get("keep_zeros")
    if k:
        print("keep_zeros %.2f Mdof/s %.1f ms  spmv frac %.3f" % (k["value"], k["ms_per_step"], k["spmv_frac_of_peak"]))
    g = d.get("gmg")
    if g and g.get("value"):
        print("gmg %.1f Mdof/s %.1f ms/step %d iterations; cpu %s" % (g["value"], g["ms_per_step"], g["iterations"], (g.get("cpu_baseline") or {}).get("value")))
    c = d.get("c3")
    if c and c.get("value"):
        j, a = c["jacobi"], c["default_solve_amg"]
        j.setdefault("roofline", None)
        if not j["roofline"]:
            j["roofline"] = {"frac": float("nan")}
        print("   norms: %s %s" % (c.get("solution_l2_norm"), c.get("max_abs_deflection_z")))
        print("c3 jacobi %.2f Mdof/s %.0f ms %d its spmv frac %.3f | solve_amg %.1f Mdof/s %.1f ms %d its | cross %.1e | cpu %s" % (j["value"], j["ms_per_step"], j["iterations"], j["roofline"]["frac"], a["value"], a["ms_per_step"], a["iterations"], c["rel_l2_default_vs_jacobi"], (c.get("cpu_baseline") or {}).get("value")))
    c = d.get("c4")
    if c and c.get("value"):
        print("c4 %.1f Mdof*steps/s %.3f ms/time step, its %s, oracle diff %s, cpu %s, roofline %s" % (c["value"], c["ms_per_time_step"], c["iterations_per_step"], c.get("rel_l2_vs_cpu_oracle_after_20_steps"), (c.get("cpu_baseline") or {}).get("value"), (c.get("roofline") or {}).get("frac")))
    c = d.get("p2")
    if c and c.get("value"):
        print("p2 %.2f Mdof/s %.1f ms %d its err %.1e spmv frac %.3f cpu %s" % (c["value"], c["ms_per_step"], c["iterations"], c["rel_l2_vs_exact"], c["roofline"]["frac"], (c.get("cpu_baseline") or {}).get("value")))
    c = d.get("c5")
    if c and c.get("value"):
        print("c5 %.2f Mdof/s %.1f ms %d its err %.1e spmv frac %.3f iteration %.4f ms" % (c["value"], c["ms_per_step"], c["iterations"], c["rel_l2_vs_exact"], c["roofline"]["frac"], c["roofline"]["cg_iteration_ms"]))
        if c.get("gmg"):
            print("   c5.gmg %s" % ({k: v for k, v in c["gmg"].items() if k != "what"},))
