"""prints the essentials of one bench.py JSON line"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "parity_ok", d.get("parity_ok"), "launches", d["gpu_launches"], "scaling", d["scaling"], "n", d["n_gpus"])
if d.get("fused"):
    f = d["fused"]
    print("fused ms/step", round(f["ms_per_step"], 4), "value", round(f["value"]), "step_frac", round(f["step_frac"], 3), "parity", f["parity"]["ok"], f["parity"]["bit_exact"])
for k, v in d["phases"].items():
    print("  %-24s %.4f ms  %8.1f %s" % (k, v["ms"], v.get("tflops", v.get("gbs")), "TF/s" if "tflops" in v else "GB/s"))
r = d["roofline"]
print("roofline", r["kernel"], "frac", round(r["frac"], 3), "vs_sustained", round(r.get("frac_vs_sustained", 0), 3), "step_frac", round(r["step_frac"], 3), "step TF/s", round(r["step_conv_tflops"], 1))
print("per kernel", {k: round(v, 3) for k, v in r["per_kernel_frac"].items()})
if d.get("c2_f64"):
    c = d["c2_f64"]
    print("f64", round(c["value"]), "samples/s", round(c["ms_per_step"], 3), "ms; parity", c["parity"]["ok"], "roofline", c["roofline"]["kernel"], round(c["roofline"]["frac"], 3), "step_frac", round(c["roofline"]["step_frac"], 3))
    print("   per kernel", {k: round(v, 3) for k, v in c["roofline"]["per_kernel_frac"].items()})
if d.get("e2e"):
    e = d["e2e"]
    print("e2e", round(e["value"]), "ms", round(e["ms_per_step"], 3), "frac_of_pcie", round(e.get("frac_of_pcie", 0), 3), e.get("pcie"))
print("clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
