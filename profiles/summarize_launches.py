#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python profiles/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_xxx.md

ncu times are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < 10:
            continue
        name = r[4].split("(")[0].replace("void ", "").replace("sparf::<unnamed>::", "sparf::")[:70]
        t = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("source: `%s` (ncu launch list; per-launch times are cold-cache and serialised: read the shares)\n" % path)
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    print("\ntotal %.1f us over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
