"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share.
usage: launch_summary.py launches.csv "description of the command" > profiles/rN_launches_*_summary.txt"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= mv or not r[mv].strip():
        continue
    us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu].strip(), 1e-3)
    name = re.sub(r"\(.*", "", r[kn])
    m = re.search(r"(conv_bn_plif_kernel<[^>]*>|sampler_tc2_step_kernel<[^>]*>)", r[kn])
    key = m.group(1) if m else name[:92]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("ncu launch list of `%s`, cold-cache serialised durations: compare SHARES, not absolutes." % sys.argv[2])
print("%d launches, total %.1f ms\n" % (sum(a[0] for a in agg.values()), tot / 1e3))
print("%-94s %6s %10s %6s" % ("kernel", "count", "total us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-94s %6d %10.1f %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
