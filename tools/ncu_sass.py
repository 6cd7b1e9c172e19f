"""Per-SASS-instruction view of one kernel from `ncu --page source --csv` output.
Usage: python tools/ncu_sass.py src.csv <kernel substring> [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s in enumerate(secs):
    if want not in rows[s][1]:
        continue
    e = secs[n + 1] if n + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]
    ia, ie, isamp, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    body = [r for r in rows[s + 2:e] if len(r) > ie]
    tot = sum(int(r[ie]) for r in body) or 1
    totS = sum(int(r[isamp]) for r in body) or 1
    print("#", rows[s][1][:100], "inst", tot, "samples", totS)
    for k, r in enumerate(body):
        pi, ps = 100 * int(r[ie]) / tot, 100 * int(r[isamp]) / totS
        if pi >= min_pct or ps >= min_pct:
            print("%4d %-64s %10d %5.2f%% samp %5.2f%% thr %s" % (k, r[ia].strip()[:64], int(r[ie]), pi, ps, r[it]))
    break
