"""Executed warp-instructions and stall samples per code region of one kernel in an .ncu-rep; regions are
found as runs of SASS instructions with similar execution counts (the warp roles of a specialised kernel).
usage: ncu_regions.py report.ncu-rep [kernel-id-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
marks = ("UTCHMMA", "LDTM", "MUFU.EX2", "STS.128", "UTCBAR", "SYNCS", "EXIT", "LDG", "STG", "BAR.SYNC")
tot_s = sum(int(r[si]) for r in data)
tot_i = sum(int(r[ie]) for r in data)
print("instructions", len(data), "samples", tot_s, "warp-instructions executed", tot_i)
# split at marker EXIT / BRA boundaries where the execution count changes by > 4x
bounds = [0]
win = 24
for i in range(win, len(data) - win):
    a = sorted(int(r[ie]) for r in data[i - win:i])[win // 2]
    b = sorted(int(r[ie]) for r in data[i:i + win])[win // 2]
    if (a == 0) != (b == 0) or (a and b and (a / b > 3 or b / a > 3)):
        if i - bounds[-1] > 2 * win:
            bounds.append(i)
bounds.append(len(data))
for a0, a1 in zip(bounds[:-1], bounds[1:]):
    n = sum(int(r[si]) for r in data[a0:a1])
    ni = sum(int(r[ie]) for r in data[a0:a1])
    agg = {}
    for r in data[a0:a1]:
        for c in stall:
            if r[c].isdigit():
                agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
    top = sorted(agg.items(), key=lambda x: -x[1])[:4]
    mk = {}
    for r in data[a0:a1]:
        for k in marks:
            if k in r[src]:
                mk[k] = mk.get(k, 0) + 1
    print("[%5d,%5d) exec %9d (%4.1f%%) samples %6d (%4.1f%%) %s | %s" % (a0, a1, ni, 100.0 * ni / max(tot_i, 1), n, 100.0 * n / max(tot_s, 1),
          " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top), " ".join("%s:%d" % kv for kv in mk.items())))
