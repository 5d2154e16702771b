#!/usr/bin/env python
"""Turn the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <bench command>`
into the launch list kept under profiles/ (kernel, grid, block, ms) with the shares of the last step on top.
usage: ncu_launches.py launches.csv > profiles/rNN_launches_c5_full.txt"""
import csv
import re
import sys

CMD = ("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_(umma|score|dp|select|pack|minmax|permute|pair|"
       "epilogue|pairs) -c 60   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-sa")


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ki, gi, bi, vi = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Value"))
    out = []
    for r in rows[1:]:
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ki])).replace("void ", "")
        out.append((name, r[gi], r[bi], float(r[vi].replace(",", "")) / 1e6))
    starts = [i for i, o in enumerate(out) if o[0] == "k_permute"]
    last = out[starts[-1]:]                       # one step = from the last k_permute on
    tot = sum(o[3] for o in last)
    agg = {}
    for o in last:
        agg[o[0]] = agg.get(o[0], 0.0) + o[3]
    print("# " + CMD)
    print("# C5 at full size on one B200: the pack of the sample matrix and the pair tables on the tensor cores (once per data "
          "set), then 4 steps (3 warm-up + 1 timed).")
    print("# Per-launch times are cold-cache and serialised; compare SHARES.  One step = %.1f ms of kernels: %s." % (
        tot, ", ".join("%s %.1f %%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]))))
    print("%-28s %-16s %-14s %10s" % ("kernel", "grid", "block", "ms"))
    for o in out:
        print("%-28s %-16s %-14s %10.3f" % o)


if __name__ == "__main__":
    main()
