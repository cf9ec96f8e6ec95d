"""ncu --page source --csv --print-source cuda,sass  ->  warp instructions executed per CUDA source line (top N).
usage: python tools/ncu_source_lines.py src.csv [file-substring] [topN]"""
import csv, sys, collections
path, want, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "nn_culled.cu"), int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur, hdr = None, None
agg = collections.OrderedDict()
samples = collections.Counter()
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur is None or want not in cur:
        continue
    try:
        ln = int(r[0])
        n = float(r[hdr.index("Instructions Executed")] or 0)
        s = float(r[hdr.index("# Samples")] or 0)
    except (ValueError, IndexError):
        continue
    a = agg.setdefault(ln, [0.0, r[1]])
    a[0] += n
    samples[ln] += s
tot = sum(a[0] for a in agg.values())
stot = sum(samples.values()) or 1
print(f"{want}: {tot/1e9:.3f} G warp instructions attributed, {int(stot)} stall samples")
for ln, (n, src) in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    print(f"{ln:5d} inst {n/tot:6.2%} {n/1e6:9.1f}M  samples {samples[ln]/stot:6.2%}  {src.strip()[:100]}")
