"""Join the ordered op list of tools/profile_step.py (#ORDER lines) with an ncu launch list of the same step.
usage: python tools/join_ncu.py plain.log launches.csv  -> per-op: ncu time, roofline floor, ratio"""
import collections, csv, re, sys
ops = [l[7:].strip().split(",") for l in open(sys.argv[1]) if l.startswith("#ORDER ") and not l.startswith("#ORDER idx")]
lines = [l for l in open(sys.argv[2]) if l.startswith('"')]
r = list(csv.reader(lines)); hdr = r[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for row in r[1:]:
    name = re.sub(r"\(.*", "", row[ki]).replace("<unnamed>::", "").replace("void ", "")
    v = float(row[vi].replace(",", "")); v = v / 1000 if row[ui] == "ns" else v * 1000 if row[ui] == "ms" else v
    seq.append((name, v))
NL = {"cbam_spatial": 2}     # ops that launch two kernels (the tail is one launch since the class-branch conv writes the keys)
if len(seq) - len(ops) == 2:
    NL["v10_decode_topk"] = 2      # standalone key pass (fp32 mode / unsupported shapes)
j = 0; out = []
for idx, kind, *rest in ops:
    tag = ",".join(rest[:-4]); us_e, gf, mb, floor = map(float, rest[-4:])
    n = NL.get(kind, 1); t = sum(v for _, v in seq[j:j + n]); kn = seq[j][0] if j < len(seq) else "?"; j += n
    out.append((kind, tag, kn, t, floor, gf, mb))
tot = sum(o[3] for o in out); ftot = sum(o[4] for o in out)
print(f"# {len(out)} ops / {j} of {len(seq)} launches; ncu total {tot:.1f} us; sum of roofline floors {ftot:.1f} us ({ftot/tot:.2f})")
print("kind,tag,kernel,ncu_us,floor_us,ratio,TFLOP/s,GB/s")
for kind, tag, kn, t, floor, gf, mb in sorted(out, key=lambda o: -(o[3] - o[4])):
    print(f"{kind},{tag},{kn[:34]},{t:.1f},{floor:.1f},{t/max(floor,1e-9):.1f},{gf/t/1e3*1e3:.0f},{mb/t*1e3/1e3:.0f}")
agg = collections.OrderedDict()
for kind, tag, kn, t, floor, gf, mb in out:
    a = agg.setdefault(kn, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += floor
print("\nkernel,launches,ncu_us,floor_us,frac_of_roofline")
for k, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]},{n},{t:.1f},{f:.1f},{f/t:.2f}")
