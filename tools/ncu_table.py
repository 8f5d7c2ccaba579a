"""Joins an ncu capture of tools/ncu_shapes.py with the launch order it wrote and emits profiles/ncu_table.json, the table
bench.py's `roofline.traffic` / `roofline.ncu` are read from (run here, no GPU needed):

    python tools/ncu_table.py gpurun_out/r02_shapes.ncu-rep gpurun_out/ncu_shapes_order.json [profiles/ncu_table.json]
    (or the raw page exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > gpurun_out/r02_shapes_raw.csv)

Per (kernel, shape) key: traffic = dram__bytes_read.sum + dram__bytes_write.sum (bytes, summed over the kernels of the
call), ms = gpu__time_duration (summed), tensor_pipe_pct = sm__pipe_tensor_cycles_active (duration-weighted mean), plus
the per-kernel rows. Also writes the selected raw metrics next to it as CSV."""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = {"ms": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
           "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "tc_smem_pct": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "regs": "launch__registers_per_thread", "inst": "smsp__inst_executed.sum"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3,
        "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def main():
    rep, order_path = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                  "profiles", "ncu_table.json")
    if rep.endswith(".csv"):  # `ncu -i X.ncu-rep --page raw --csv` run on the GPU box (full reports exceed the copy-back limit)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}

    def val(r, metric):
        i = col[metric]
        try:
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        except ValueError:
            return 0.0

    kernels = [dict(name=r[col["Kernel Name"]], **{k: val(r, m) for k, m in METRICS.items() if m in col}) for r in data]
    order = json.load(open(order_path))
    entries, pos, sel = {}, 0, []
    for key, pats in order:
        got = []
        for pat in pats:
            while pos < len(kernels) and not re.search(pat, kernels[pos]["name"]):
                pos += 1
            if pos >= len(kernels):
                raise SystemExit(f"capture ended before {key} / {pat}")
            got.append(kernels[pos])
            pos += 1
        ms = sum(k["ms"] for k in got)
        entries[key] = {"traffic": sum(k["dram_read"] + k["dram_write"] for k in got), "ms": round(ms, 4),
                        "tensor_pipe_pct": round(sum(k["tensor"] * k["ms"] for k in got) / ms, 2),
                        "kernels": [{"name": k["name"].split("(")[0][-60:], "ms": round(k["ms"], 4), "tensor_pipe_pct": round(k["tensor"], 2),
                                     "dram_pct": round(k.get("dram_pct", 0.0), 2), "tc_smem_pct": round(k.get("tc_smem_pct", 0.0), 2),
                                     "issue_active_pct": round(k.get("issue_active_pct", 0.0), 2), "regs": int(k.get("regs", 0))}
                                    for k in got]}
        for k in got:
            sel.append([key] + [k["name"].split("(")[0]] + [k.get(m, "") for m in METRICS])
    with open(out_path, "w") as f:
        json.dump({"source": os.path.basename(rep) + " (ncu --set full --clock-control none, tools/ncu_shapes.py)",
                   "entries": entries}, f, indent=1)
    with open(os.path.splitext(out_path)[0] + "_raw_selected.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["shape_key", "kernel"] + list(METRICS))
        w.writerows(sel)
    print(f"{len(entries)} entries -> {out_path}")


if __name__ == "__main__":
    main()
