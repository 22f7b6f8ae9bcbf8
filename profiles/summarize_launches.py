"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share.
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_xxx.md"""
import csv, sys, collections, re
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    key = (name, r["Grid Size"], r["Block Size"]) if "--by-grid" in sys.argv else (name,)
    ns = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"): ns *= 1e3
    if r["Metric Unit"] in ("ms", "msecond"): ns *= 1e6
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"launches: {sum(v[0] for v in agg.values())}, total device time {tot/1e6:.3f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {' '.join(k)} | {n} | {ns/1e6:.3f} | {100*ns/tot:.1f}% | {ns/n/1e3:.1f} |")
