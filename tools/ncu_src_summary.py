"""Summarise `ncu --page source --csv` output: instructions executed / stall samples per SASS region.
usage: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_src_summary.py f.csv [bucket]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iinst, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[ia], r[isrc], int(r[iinst] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > isamp]
tot_i = sum(d[2] for d in data); tot_s = sum(d[3] for d in data)
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}  sass lines {len(data)}")
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k in range(0, len(data), bucket):
    seg = data[k:k + bucket]
    si, ss = sum(d[2] for d in seg), sum(d[3] for d in seg)
    ops = {}
    for d in seg:
        op = d[1].split()[0] if not d[1].startswith("@") else d[1].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + d[2]
    top = ", ".join(f"{o}:{100*c/max(si,1):.0f}%" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print(f"[{k:4d}-{k+len(seg)-1:4d}] inst {100*si/tot_i:5.1f}%  samples {100*ss/max(tot_s,1):5.1f}%  {top}")
