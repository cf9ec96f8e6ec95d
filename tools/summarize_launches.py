"""ncu launch list (csv, gpu__time_duration.sum) -> per-kernel table (markdown)."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {us:.1f} | {us/tot:.1%} |")
