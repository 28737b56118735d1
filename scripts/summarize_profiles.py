"""Turn the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
usage: summarize_profiles.py launches <csv> <out.txt> <title...>   |   full <raw csv> <out.json>"""
import csv
import json
import sys
from collections import OrderedDict

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(path)) if r and r[0] == "ID")
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        v = float(r[ival].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iunit], 1e-6)
        name = r[iname].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(title + "\n")
        f.write("total %.3f ms over %d launches\n" % (total, sum(a[0] for a in agg.values())))
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-50s n=%6d total %8.3f ms  share %5.1f%%  avg %.4f ms\n" % (name[:50], n, ms, 100 * ms / total, ms / n))


def full(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    res = OrderedDict()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0]
        key = short
        i = 2
        while key in res:
            key = "%s #%d" % (short, i)
            i += 1
        d = OrderedDict([("Kernel Name", name)])
        for k in KEYS:
            if k in hdr:
                j = hdr.index(k)
                d[k] = ("%s %s" % (r[j], units[j])).strip()
        res[key] = d
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], " ".join(sys.argv[4:]))
    else:
        full(sys.argv[2], sys.argv[3])
