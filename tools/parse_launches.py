"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share."""
import csv
import sys
from collections import OrderedDict


def main(path, verbose=False):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    tot = 0.0
    agg = OrderedDict()
    for r in rows:
        name = r[4].replace("void unnamed>::", "").split("(")[0]
        val = float(r[-1].replace(",", ""))
        if r[-2] == "us":
            val *= 1e3
        elif r[-2] == "ms":
            val *= 1e6
        tot += val
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
        if verbose:
            print("%3s %-60s %10.1f us" % (r[0], name[:60], val / 1e3))
    print("%-62s %8s %12s %7s" % ("kernel", "launches", "total us", "share"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %8d %12.1f %6.1f%%" % (k[:62], n, v / 1e3, 100 * v / tot))
    print("%-62s %8d %12.1f" % ("TOTAL", len(rows), tot / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
