#!/usr/bin/env python
"""tools/launch_breakdown.py <ncu launch list .csv> [marker kernel] -- per-kernel times of ONE step from an
`ncu --metrics gpu__time_duration.sum --csv` launch list: the launches between two consecutive occurrences of the marker."""
import csv
import re
import sys


def main():
    path, marker = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "k_probe")
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -2
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in rows]
    idx = [i for i, (n, _) in enumerate(names) if marker in n]
    i0, i1 = idx[which], idx[which + 1] if which + 1 < 0 or which + 1 < len(idx) else len(names)
    tot = 0.0
    for n, v in names[i0:i1]:
        n = re.sub(r"void |<unnamed>::|cub::CUB_[0-9a-z_]*::", "", n)
        n = re.sub(r"\(.*", "", n)[:70]
        print(f"{v / 1e3:9.1f} us  {n}")
        tot += v
    print(f"{tot / 1e6:9.3f} ms total")


if __name__ == "__main__":
    main()
