"""Per-launch table (duration, tensor-pipe %, DRAM bytes, ...) of an ncu --set full report."""
import csv
import re
import subprocess
import sys

COLS = [("us", "gpu__time_duration.sum"), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smemTC%", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("L2hit%", "lts__t_sector_hit_rate.pct"), ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-3s %-46s " % ("#", "kernel") + " ".join("%9s" % c for c, _ in COLS))
    for n, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]]
        name = re.sub(r"\(.*$", "", name.replace("void ", "").replace("te::<unnamed>::", "").replace("te::", ""))
        vals = []
        for c, k in COLS:
            if k not in idx:
                vals.append("%9s" % "-")
                continue
            v = float(r[idx[k]].replace(",", "") or 0) * SCALE.get(units[idx[k]], 1.0)
            vals.append("%9.1f" % v)
        print("%-3d %-46s " % (n, name[:46]) + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
