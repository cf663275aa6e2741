"""Turns the ncu captures of a round into the tracked summaries under profiles/:

    python scripts/export_profiles.py r02 gpurun_out/r02_prof_step.ncu-rep gpurun_out/r02_ncu_launches.csv ROWS

  profiles/<tag>_ncu_full_<kernel>.json   every raw metric of the first captured launch of each kernel (ncu --set full)
  profiles/<tag>_ncu_launches.csv         the launch list (gpu__time_duration.sum per launch) of the same bench command
  profiles/<tag>_ncu_launch_shares.json   per kernel: launches, total ns, share of the captured window
  profiles/traffic.json                   dram bytes read + written per launch (what bench.py reports as roofline.traffic)
ROWS = rows of the captured launches (bench.py's chunk_rows)."""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG, REP, LAUNCHES = sys.argv[1], sys.argv[2], sys.argv[3]
ROWS = int(sys.argv[4]) if len(sys.argv) > 4 else 4194304
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name


out = subprocess.run(["ncu", "-i", REP, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
traffic = {}
for r in rows[2:]:
    m = {f"{h} [{u}]" if u else h: v for h, u, v in zip(hdr, units, r)}
    k = short(m["Kernel Name"])
    if k in traffic:
        continue
    json.dump(m, open(os.path.join(ROOT, "profiles", f"{TAG}_ncu_full_{k}.json"), "w"), indent=1)

    def get(name):
        for key, v in m.items():
            if key.startswith(name + " ["):
                return float(v.replace(",", "")) * UNIT[key[len(name) + 2:-1]]
        return 0.0
    traffic[k] = {"dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"), "rows": ROWS}
    print(k, traffic[k])
traffic["_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum of ONE launch over a {ROWS}-row chunk of the 256^3 grid, from the "
                    f"ncu --set full capture summarised in profiles/{TAG}_ncu_full_*.json; bench.py scales it linearly to the rows of "
                    "its own launches")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

dst = os.path.join(ROOT, "profiles", f"{TAG}_ncu_launches.csv")
with open(LAUNCHES) as f, open(dst, "w") as g:
    for line in f:
        if line.startswith('"'):
            g.write(line)
share = {}
for r in csv.DictReader(open(dst)):
    k = short(r["Kernel Name"])
    e = share.setdefault(k, {"launches": 0, "ns": 0.0})
    e["launches"] += 1
    e["ns"] += float(r["Metric Value"].replace(",", ""))
total = sum(e["ns"] for e in share.values())
for e in share.values():
    e["share"] = e["ns"] / total
json.dump(dict(sorted(share.items(), key=lambda kv: -kv[1]["ns"])), open(os.path.join(ROOT, "profiles", f"{TAG}_ncu_launch_shares.json"), "w"), indent=1)
print(json.dumps({k: round(v["share"], 3) for k, v in sorted(share.items(), key=lambda kv: -kv[1]["ns"])[:8]}))
