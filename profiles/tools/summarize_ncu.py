"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports: one table row per captured kernel launch with the metrics the roofline
discussion uses.  usage: summarize_ncu.py out.md file1.csv file2.csv ..."""
import csv, sys, os, re

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def col(hdr, key):
    for i, h in enumerate(hdr):
        if h == key or h.endswith("." + key) or h.endswith(key):
            return i
    return None


out = ["| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---:|" * len(WANT)]
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", r[kn].replace("void ", "").replace("<unnamed>::", ""))
        cells = []
        for key, _ in WANT:
            i = col(hdr, key)
            if i is None or r[i] == "":
                cells.append("-")
                continue
            v = r[i]
            try:
                f = float(v)
                v = ("%.3f" % f if f < 100 else "%.0f" % f)
            except ValueError:
                pass
            cells.append("%s %s" % (v, units[i]) if units[i] not in ("", "%") else v)
        out.append("| `%s` | " % name[:70] + " | ".join(cells) + " |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
