#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name and grid, launches, median and
total time, share of the list (cold-cache, serialised launches: the SHARES are comparable with the live step, not the absolutes).
python tools/launch_summary.py gpurun_out/<list>.csv [first_id last_id]"""
import csv
import statistics
import sys
from collections import OrderedDict


def main(path, lo=None, hi=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    iname, igrid, ival, iid = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value"), h.index("ID")
    groups = OrderedDict()
    for r in rows[1:]:
        k = int(r[iid])
        if (lo is not None and k < lo) or (hi is not None and k > hi):
            continue
        name = r[iname].split("(")[0].replace("void ", "").replace("dwdf::<unnamed>::", "").replace("<unnamed>::", "")
        groups.setdefault((name[:70], r[igrid]), []).append(float(r[ival].replace(",", "")) / 1e3)
    total = sum(sum(v) for v in groups.values())
    print(f"{'kernel':70s} {'grid':>16s} {'n':>5s} {'median us':>10s} {'total us':>10s} {'share':>6s}")
    for (name, grid), v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name:70s} {grid:>16s} {len(v):5d} {statistics.median(v):10.1f} {sum(v):10.1f} {100 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], *(int(a) for a in sys.argv[2:4]))
