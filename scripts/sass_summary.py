"""Counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, per compiled object:
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync,
FFMA2 = packed fp32 FMA, ATOMS/RED = shared / global atomics.  Runs on the CPU box (cuobjdump).
usage: python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "eas_snn_b200", "build")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FFMA", "ATOMS", "RED", "ATOMG",
       "LDGSTS", "MUFU"]
print("%-22s %8s " % ("object", "instrs") + " ".join("%8s" % p for p in PAT))
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass, re.M)
    cnt = {p: 0 for p in PAT}
    for o in ops:
        for p in PAT:
            if o == p or (p in ("RED", "ATOMS", "ATOMG", "MUFU", "SYNCS") and o.startswith(p)):
                cnt[p] += 1
    print("%-22s %8d " % (f, len(ops)) + " ".join("%8d" % cnt[p] for p in PAT))
arch = subprocess.run(["cuobjdump", "-lelf", os.path.join(OBJ, "sampler_tc2.o")], capture_output=True, text=True).stdout
print("\nelf images of sampler_tc2.o:", ", ".join(l.split(":")[-1].strip() for l in arch.splitlines() if "sm_" in l))
