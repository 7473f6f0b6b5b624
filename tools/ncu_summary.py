"""Summarise ncu output for profiles/:  ncu_summary.py launches <csv>   |   ncu_summary.py full <ncu-rep>"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total us", "avg us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %10.2f %6.1f%%" % (k[:60], n, t, t / n, 100 * t / tot))
    print("%-60s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[-1]
        print("== %s   (%d launches captured, last shown)" % (name, len(rs)))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("   %-82s %16s %s" % (w, r[i], units[i]))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        def b(i):
            v = float(r[i]); u = units[i]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        print("   %-82s %16.0f byte" % ("dram traffic per launch (read+write)", b(rd) + b(wr)))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
