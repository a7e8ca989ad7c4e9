#!/usr/bin/env python
"""Print the roofline-relevant counters of every launch in an .ncu-rep (needs `ncu` on PATH).
  ncu_metrics.py report.ncu-rep [--traffic-json out.json]
--traffic-json merges {kernel name (template arguments kept, parameters dropped): dram bytes read + written per launch,
duration} into out.json -- the file bench.py reads for `roofline.traffic`."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("-" * 100)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:70s} {r[i][:70]} {units[i]}")

if "--traffic-json" in sys.argv:
    import json, os
    path = sys.argv[sys.argv.index("--traffic-json") + 1]
    db = json.load(open(path)) if os.path.exists(path) else {}
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i]) * to_bytes[units[i]]
        i = hdr.index("gpu__time_duration.sum")
        us = float(r[i]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[i], 1.0)
        db[name] = {"dram_bytes": tot, "duration_us": us, "report": os.path.basename(sys.argv[1])}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
