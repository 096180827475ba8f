"""Stall samples of one kernel of an .ncu-rep aggregated per CUDA source line (the ncu CLI prints no
per-line metrics): SASS order is matched with `nvdisasm -g` of the object's cubin.
usage: ncu_lines.py report.ncu-rep object.o kernel-substring [kernel-id-index] [min-percent]"""
import csv, os, re, subprocess, sys, tempfile
rep, obj, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
kid = sys.argv[4] if len(sys.argv) > 4 else "1"
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, cur, infn, line, fname = [], None, False, None, None
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        infn = ksub in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        fname, line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        lines.append((fname, line, m.group(2).strip()))
n = min(len(lines), len(data))
if len(lines) != len(data):
    print("warning: %d SASS instructions in the cubin vs %d in the report" % (len(lines), len(data)))
tot = sum(int(r[si]) for r in data)
agg = {}
for (f, ln, txt), r in zip(lines[:n], data[:n]):
    a = agg.setdefault((f, ln), {"n": 0, "ex": 0, "cnt": 0, "st": {}})
    a["n"] += int(r[si]); a["ex"] += int(r[ie]); a["cnt"] += 1
    for c in stall:
        if r[c].isdigit() and int(r[c]):
            a["st"][hdr[c][6:]] = a["st"].get(hdr[c][6:], 0) + int(r[c])
srcs = {}
print("samples", tot, "instructions", n)
for (f, ln), a in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1] or 0)):
    if a["n"] < tot * minpct / 100:
        continue
    text = ""
    if f and f.endswith(".cu"):
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        if os.path.exists(p):
            srcs.setdefault(f, open(p).read().splitlines())
            text = srcs[f][ln - 1].strip()[:70] if ln - 1 < len(srcs[f]) else ""
    top = " ".join("%s=%d" % kv for kv in sorted(a["st"].items(), key=lambda x: -x[1])[:3])
    print("%-22s %5s %6d %5.1f%% sass=%3d exec=%9d | %-70s | %s" % (f, ln, a["n"], 100.0 * a["n"] / tot, a["cnt"], a["ex"], text, top))
