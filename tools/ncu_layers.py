"""Per-layer table of one chunk of the U-net path from an `ncu --set full` report of tools/prof_chunk.py, and the DRAM
traffic file bench.py reads (profiles/traffic_r02.json).

    python tools/ncu_layers.py gpurun_out/r02_layers.ncu-rep gpurun_out/r02_layer_names.json \
        profiles/ncu_r02_layers.txt profiles/traffic_r02.json
"""
import csv
import json
import re
import subprocess
import sys

COLS = [("us", "gpu__time_duration.sum"), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smemTC%", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("issue%", "sm__inst_executed.avg.pct_of_peak_sustained_elapsed"),
        ("sm_MHz", "sm__cycles_elapsed.avg.per_second"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3,
         "hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}


def main(rep, names_json, out_txt, out_json):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = json.load(open(names_json))
    body = rows[2:]
    lines = ["one chunk (32 x 64^3 tiles of a 512^3 float16 volume) of the BASELINE configs[1] path, ncu --set full --clock-control none",
             "%-18s %-34s %8s " % ("layer", "kernel", "TFLOP/s") + " ".join("%8s" % c for c, _ in COLS)]
    traffic = {}
    for n, r in enumerate(body):
        meta = names[n] if n < len(names) else "?"
        layer = meta["layer"] if isinstance(meta, dict) else meta
        kname = r[idx["Kernel Name"]]
        kname = re.sub(r"\(.*$", "", kname.replace("(int)", "").replace("void ", "").replace("te::<unnamed>::", "").replace("te::", ""))
        vals = {}
        for c, k in COLS:
            if k in idx and r[idx[k]] != "":
                vals[c] = float(r[idx[k]].replace(",", "")) * SCALE.get(units[idx[k]], 1.0)
        tf = (meta["gflop_per_chunk"] / 1e3 / (vals["us"] / 1e6)) if isinstance(meta, dict) and vals.get("us") else None
        lines.append("%-18s %-34s %8s " % (layer, kname[:34], "%.0f" % tf if tf else "-") +
                     " ".join("%8.1f" % vals[c] if c in vals else "%8s" % "-" for c, _ in COLS))
        traffic[layer] = {"kernel": (meta["kernel"].replace("+head", "") if isinstance(meta, dict) else kname),
                          "us": vals.get("us"), "dram_read_bytes": vals.get("rdMB", 0.0) * 1e6,
                          "dram_write_bytes": vals.get("wrMB", 0.0) * 1e6,
                          "tensor_pipe_pct": vals.get("tensor%")}
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"source": rep, "what": "ncu --set full, one launch per layer at 32 x 64^3 patches", "layers": traffic},
              open(out_json, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:5])
