"""Break an `ncu --metrics gpu__time_duration.sum --csv` launch list of scripts/prof_backbone.py down by kernel
(last forward only)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = [(r[ki], r[gi], float(r[vi].replace(",", ""))) for r in rows[start:] if len(r) > vi]
idx = [i for i, (n, g, v) in enumerate(data) if "conv_bn_plif_kernel<64, 1, 16" in n]
fw = data[idx[-1] - 8:]
agg = {}
for n, g, v in fw:
    k = n.split("(")[0][-60:]
    agg.setdefault(k, [0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:12]:
    print("%-62s %10.1f us %4d launches %5.1f%%" % (k, v[0] / 1e3, v[1], 100 * v[0] / tot))
print("total us (last forward)", tot / 1e3)
conv = [(g, v / 1e3, n) for n, g, v in fw if "conv_bn_plif" in n]
print(len(conv), "conv launches", sum(v for _, v, _ in conv), "us")
print(" | ".join("%s %s %.0f" % (g.replace(", 1, 1", ""), re.search(r"kernel<([^>]*)>", n).group(1), v) for g, v, n in conv))
