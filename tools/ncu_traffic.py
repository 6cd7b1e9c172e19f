"""profiles/r1_ncu_traffic.json from an `ncu --set full` report: per kernel name, the launches of the LAST iteration
of tools/profile_step.py (duration, DRAM bytes, issue-slot utilisation, warp instructions).
Usage: python tools/ncu_traffic.py report.ncu-rep out.json [launches_per_iteration_divisor=3]"""
import csv
import json
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum")}


def scaled(r, key):
    v, u = float(r[col[key]].replace(",", "")), units[col[key]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
            "msecond": 1e3, "%": 1, "inst": 1, "": 1}
    return v * mult.get(u, 1)


body = rows[2:]
n = len(body) // iters
kernels = {}
for r in body[-n:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("lg::", "")
    kernels.setdefault(name, []).append({
        "dram_read_bytes": scaled(r, "dram__bytes_read.sum"), "dram_write_bytes": scaled(r, "dram__bytes_write.sum"),
        "duration_us": round(scaled(r, "gpu__time_duration.sum"), 3),
        "issue_active_pct": scaled(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": scaled(r, "smsp__inst_executed.sum")})
json.dump({"source": "ncu --set full --clock-control none --import-source on, tools/profile_step.py 1000000 ours %d (metric "
                     "scene: 1M Gaussians, 800x800, SH3; last iteration, one launch each; cold-cache, serialised)" % iters,
           "kernels": kernels}, open(out_path, "w"), indent=1)
tot = sum(k["duration_us"] for v in kernels.values() for k in v)
print("kernels", len(kernels), "sum of durations %.1f us" % tot)
