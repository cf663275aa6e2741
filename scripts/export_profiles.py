"""Reads gpurun_out/prof_<kernel>.ncu-rep (ncu --set full captures made by scripts/gpu_ncu_*.sh) and writes
profiles/r01_ncu_full_<kernel>.json (every raw metric of the captured launch) plus profiles/traffic.json
(dram bytes read + written per launch, what bench.py reports as roofline.traffic)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 262144      # rows of the captured launch (scripts/gpu_ncu_hoist.sh: --chunk 262144)
KERNELS = {"hoist_addend": "hoist_addend_kernel", "hoist_rest": "hoist_rest_kernel", "mlp_tc": "mlp_tc_kernel"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{f"{h} [{u}]" if u else h: v for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


traffic = {}
for short, kernel in KERNELS.items():
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{short}.ncu-rep")
    if not os.path.exists(rep):
        continue
    launches = raw(rep)
    dst = os.path.join(ROOT, "profiles", f"{TAG}_ncu_full_{kernel}.json")
    json.dump(launches, open(dst, "w"), indent=1)
    m = launches[0]

    def get(name):
        for k, v in m.items():
            if k.startswith(name + " ["):
                return float(v.replace(",", "")) * UNIT[k[len(name) + 2:-1]]
        return None
    traffic[kernel] = {"dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"), "rows": ROWS}
    print(kernel, "dram bytes/launch", traffic[kernel], "->", dst)
traffic["_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum of ONE launch over a {ROWS}-row chunk of the 256^3 grid, from the "
                    f"ncu --set full captures in profiles/{TAG}_ncu_full_*.json (bench.py --chunk {ROWS}); bench.py scales it "
                    "linearly to the rows of its own launches")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
