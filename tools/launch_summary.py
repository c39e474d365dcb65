"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if r[0] == 'ID':
        hdr = r; rows = rows[i + 1:]; break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows:
    agg[r[ki].split('(')[0][-60:]].append(float(r[vi].replace(',', '')))
tot = sum(sum(v) / len(v) for k, v in agg.items() if 'ssdx' in k)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
    if 'ssdx' in k:
        m = sum(v) / len(v) / 1000
        print(f"{k:60s} n={len(v):4d} mean={m:9.1f} us  share={100 * m * 1000 / tot:5.1f}%")
print(f"sum of per-launch means (one step): {tot / 1000:.1f} us")
