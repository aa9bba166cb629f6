"""Print the key numbers of a bench.py JSON line:  python scripts/bench_brief.py gpurun_out/bench_x.json"""
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print("predictive", f"{d['value']:.4g}", "ms", round(d["ms_per_step"], 4), {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["kernels"].items()})
wp = d.get("with_probs")
if wp: print("with_probs ms", round(wp["ms_per_step"], 4), {k: round(v["avg_ms"], 4) for k, v in wp["kernels"].items()})
for key in ("kfac", "kfac_cfg2", "kfac_h14", "kfac_siglip"):
    k = d.get(key)
    if k: print(key, f"{k['value']:.4g}", "ms", round(k["ms_per_step"], 3), "frac", round(k.get("frac_of_bf16_sustained", 0), 3),
                {a: round(b["avg_ms"], 4) for a, b in k["kernels"].items()})
e = d.get("epig")
if e:
    print("epig", f"{e['value']:.4g}", "ms", round(e["ms"], 2), "joint", round(e["joint_kernel_ms"], 2), "mufu", round(e["joint_frac_of_mufu_roof"], 3),
          "red ms", round(e["reductions"]["ms"], 3), "gbs", round(e["reductions"]["gbs"]), "parity", e.get("parity", {}).get("exact_match_rate"))
    c = e.get("cl65")
    if c: print("epig65", f"{c['value']:.4g}", "ms", round(c["ms"], 2), "mufu", round(c["joint_frac_of_mufu_roof"], 3), "red gbs", round(c["reductions"]["gbs"]))
print("e2e", f"{d['e2e']['value']:.4g}", round(d["e2e"]["ms_per_step"], 3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("parity_check", d.get("parity_check", {}).get("max_excess"))
ss = d.get("kfac", {}).get("syrk_solo")
if ss: print("syrk_solo", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in ss.items() if k not in ("flops", "unit")})
