import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
tot=sum(int(r[ix["# Samples"]]) for r in data)
print("total samples",tot, "n instr", len(data), rows[0][1][:80])
top=sorted(data,key=lambda r:-int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in top:
    st=sorted(((int(r[ix[c]]),c) for c in stall_cols),reverse=True)[:2]
    print(f'{int(r[ix["# Samples"]]):7d} {100*int(r[ix["# Samples"]])/tot:5.1f}% exec={r[ix["Instructions Executed"]]:>10s} {r[ix["Source"]].strip()[:64]:64s} {st}')
# totals by stall
agg={c:sum(int(r[ix[c]]) for r in data) for c in stall_cols}
print(sorted(((v,k) for k,v in agg.items()),reverse=True)[:8])
