#!/usr/bin/env python
"""Key per-kernel metrics of an `ncu --set full` report as a markdown table.

    ncu -i gpurun_out/xxx.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv "title" > profiles/rNN_ncu_xxx.md
"""
import csv
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "SM cycles (slowest)"),
    ("smsp__cycles_active.avg", "SMSP cycles active (mean)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active, % of active cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  of which bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")[:40] for r in body]
    print("# %s\n" % title)
    print("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in METRICS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in body:
            v = r[i].replace(",", "")
            try:
                f = float(v)
                v = ("%.3f" % f).rstrip("0").rstrip(".") if abs(f) < 1e4 else "%.3e" % f
            except ValueError:
                pass
            vals.append("%s %s" % (v, units[i]))
        print("| %s | " % label + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
