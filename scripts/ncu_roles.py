"""Per-instruction stall samples of one kernel from an .ncu-rep, grouped into code regions.
usage: ncu_roles.py report.ncu-rep [kernel-id-index] [boundary,boundary,...]"""
import csv, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, src = hdr.index("# Samples"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("instructions", len(data), "samples", tot)
marks = {}
for i, r in enumerate(data):
    for k in ("UTCHMMA", "LDTM", "MUFU.EX2", "STS.128", "UTCBAR", "SYNCS", "EXIT", "LDG", "STG", "ST.E"):
        if k in r[src]:
            marks.setdefault(k, []).append(i)
for k, v in marks.items():
    print("  %-8s n=%3d first %s last %s" % (k, len(v), v[:4], v[-2:]))
if len(sys.argv) > 3:
    b = [int(x) for x in sys.argv[3].split(",")]
    b = [0] + b + [len(data)]
    for a0, a1 in zip(b[:-1], b[1:]):
        agg, n = {}, 0
        for r in data[a0:a1]:
            n += int(r[si])
            for c in stall:
                if r[c].isdigit():
                    agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
        top = sorted(agg.items(), key=lambda x: -x[1])[:5]
        print("region [%d,%d): %d samples (%.1f%%) %s" % (a0, a1, n, 100.0 * n / tot, top))
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:25]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall if r[c].isdigit()), reverse=True)[:2]
    print("%5d %6d %5.1f%% | %-60s | %s" % (i, int(r[si]), 100.0 * int(r[si]) / tot, r[src].strip()[:60],
                                          ", ".join("%s=%d" % (h, v) for v, h in st)))
