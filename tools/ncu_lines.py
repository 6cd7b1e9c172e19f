"""Hot CUDA source lines of a kernel from `ncu -i rep --page source --csv --print-source cuda,sass`:
samples, executed instructions and the dominant stall reasons per source line.
Usage: python tools/ncu_lines.py src.csv [kernel substring] [top N]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sec = [k for k, r in enumerate(rows) if r and r[0] == "Function Name"]
for si, k in enumerate(sec):
    name = rows[k][1]
    if want not in name:
        continue
    hdr = rows[k + 1]
    end = sec[si + 1] - 1 if si + 1 < len(sec) else len(rows)
    body = [r for r in rows[k + 2:end] if len(r) == len(hdr)]
    iline, isamp, iex, iaddr = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
    stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    cuda = [r for r in body if r[iaddr] in ("", "-")]
    src = cuda if cuda else body
    tot_s = sum(num(r[isamp]) for r in src) or 1
    tot_i = sum(num(r[iex]) for r in src) or 1
    print("==", name[:70], "| lines", len(src), "samples", tot_s, "warp instructions", tot_i)
    for r in sorted(src, key=lambda r: -num(r[isamp]))[:topn]:
        st = sorted(((num(r[j]), h) for j, h in stall_cols), reverse=True)[:3]
        print("%5s %-78s samp %5.1f%% inst %5.1f%%  %s" % (r[iline], r[1].strip()[:78], 100 * num(r[isamp]) / tot_s,
                                                          100 * num(r[iex]) / tot_i,
                                                          " ".join("%s:%d" % (h[6:], v) for v, h in st if v)))
    tot = collections.Counter()
    for r in src:
        for j, h in stall_cols:
            tot[h] += num(r[j])
    print("   stall totals:", ", ".join("%s %d" % (h[6:], v) for h, v in tot.most_common(8)))
