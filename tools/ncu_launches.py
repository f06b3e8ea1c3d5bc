"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + the launches in order.
usage: python tools/ncu_launches.py gpurun_out/launches.csv [min_us]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(lines))
hdr = r[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
seq = []
for row in r[1:]:
    name = re.sub(r"\(.*", "", row[ki]).replace("<unnamed>::", "").replace("void ", "")
    v = float(row[vi].replace(",", ""))
    v = v / 1000 if row[ui] == "ns" else v * 1000 if row[ui] == "ms" else v
    seq.append((name, v, row[gi] if gi is not None else ""))
tot = sum(v for _, v, _ in seq)
agg = collections.OrderedDict()
for n, v, _ in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
print(f"# {len(seq)} launches, {tot:.1f} us (ncu gpu__time_duration: serialised, cold caches)")
print("kernel,launches,us,share")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:80]},{n},{v:.1f},{v/tot:.3f}")
mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
print("\nidx,kernel,grid,us")
for i, (n, v, g) in enumerate(seq):
    if v >= mn:
        print(f"{i},{n[:60]},{g},{v:.1f}")
