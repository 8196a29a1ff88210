"""Top stall-sample instructions of an ncu report's SASS source page (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    si, so = idx["# Samples"], idx["Source"]
    stall_cols = [(h, i) for h, i in idx.items() if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    for ln, r in enumerate(rows[2:]):
        try:
            v = int(r[si])
        except (ValueError, IndexError):
            continue
        tot += v
        data.append((v, ln, r))
    print("total samples", tot)
    for v, ln, r in sorted(data, key=lambda t: -t[0])[:top]:
        reasons = sorted(((int(r[i]), h) for h, i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
        print("%6d %5.1f%% line %5d %-70s %s" % (v, 100.0 * v / max(tot, 1), ln, r[so].strip()[:70],
                                                 " ".join("%s=%d" % (h, c) for c, h in reasons)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
