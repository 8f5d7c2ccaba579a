"""Print kernel name / duration (ms) pairs from an `ncu --metrics gpu__time_duration.sum --csv` log, own kernels only."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if "s3d::" in d["Kernel Name"] and d["Metric Name"] == "gpu__time_duration.sum":
            print(f'{d["Kernel Name"][:70]:72s} {float(d["Metric Value"].replace(",", "")) / 1e6:9.3f} ms')
