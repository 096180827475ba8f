"""DRAM traffic of the launches of one kernel in an .ncu-rep -> JSON that bench.py reads by key (roofline.traffic).
usage: ncu_traffic.py report.ncu-rep kernel-substring out.json [launch index used for dram_bytes_per_launch]"""
import csv, json, subprocess, sys
rep, ksub, outp = sys.argv[1], sys.argv[2], sys.argv[3]
pick = int(sys.argv[4]) if len(sys.argv) > 4 else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit = rows[0], rows[1]
def col(name):
    return hdr.index(name)
def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
launches = []
for r in rows[2:]:
    if ksub not in r[col("Kernel Name")]:
        continue
    rd = to_bytes(r[col("dram__bytes_read.sum")], unit[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], unit[col("dram__bytes_write.sum")])
    launches.append({"kernel": r[col("Kernel Name")][:120], "duration_us": float(r[col("gpu__time_duration.sum")].replace(",", "")),
                     "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "tensor_active_pct": float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "issue_active_pct": float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")])})
# one forward = from a first-step instantiation (<.., 0> / <.., 3>) up to the next one
import re
first = [i for i, l in enumerate(launches) if re.search(r", \(?(?:int\))?[03]\)?>", l["kernel"])]
if first:
    end = first[1] if len(first) > 1 else len(launches)
    launches = launches[first[0]:end]
d = {"report": rep.split("/")[-1], "launches": launches,
     "dram_bytes_per_launch": launches[pick]["dram_read_bytes"] + launches[pick]["dram_write_bytes"],
     "dram_bytes_per_forward": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches),
     "what": "dram__bytes_read.sum + dram__bytes_write.sum of launch %d (a middle step) of %d captured with ncu --set full "
             "--cache-control all (cold L2: an upper bound for back-to-back steps)" % (pick, len(launches))}
json.dump(d, open(outp, "w"), indent=1)
print(json.dumps(d, indent=1))
