import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d["roofline"]
        tot = sum(v["ms_per_step"] for v in r["families"].values())
        print(f, "step", round(d["ms_per_step"], 2), "sum_fam", round(tot, 2), "attn_bwd", r["families"]["s3d_attn_bwd"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], "mem", d.get("hbm_peak_allocated_gb"))
