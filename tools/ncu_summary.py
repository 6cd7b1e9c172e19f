"""Summarise an .ncu-rep (raw page) into one line per launch: duration, DRAM traffic/throughput, pipe utilisation,
occupancy.  Usage: python tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "inst"),
    ("lts__t_sectors_op_red.sum", "red_sectors"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    table = []
    for r in rows[2:]:
        rec = {"kernel": r[name_i].split("(")[0].replace("void ", "").replace("lg::", "")[:40]}
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[short] = r[i] + ("" if short not in ("us", "rd", "wr") else " " + units[i])
        table.append(rec)
    cols = ["kernel"] + [s for _, s in KEYS]
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    w.writerow(cols)
    for rec in table:
        w.writerow([rec.get(c, "") for c in cols])


if __name__ == "__main__":
    main()
