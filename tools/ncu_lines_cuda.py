"""Per-CUDA-source-line view of one kernel from `ncu -i rep --page source --print-source cuda --csv`.
Usage: python tools/ncu_lines_cuda.py src_cuda.csv <kernel substring> [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s in enumerate(secs):
    if want not in rows[s][1]:
        continue
    e = secs[n + 1] if n + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]
    try:
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    except ValueError:
        print("columns:", hdr)
        break
    il = hdr.index("#") if "#" in hdr else None
    body = [r for r in rows[s + 2:e] if len(r) > max(ie, isamp)]
    num = lambda x: int(float(x.replace(",", ""))) if x.strip() not in ("", "-") else 0
    tot = sum(num(r[ie]) for r in body) or 1
    totS = sum(num(r[isamp]) for r in body) or 1
    print("#", rows[s][1][:100], "inst", tot, "samples", totS)
    for k, r in enumerate(body):
        pi, ps = 100 * num(r[ie]) / tot, 100 * num(r[isamp]) / totS
        if pi >= min_pct or ps >= min_pct:
            print("%5s %-110s inst %5.2f%% samp %5.2f%%" % (r[il] if il is not None else k, r[ia].strip()[:110], pi, ps))
    break
