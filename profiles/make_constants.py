#!/usr/bin/env python
"""Per-pair constants of the dominant kernel from a committed `ncu --set full` capture -> profiles/k2_constants.json.

    python profiles/make_constants.py <report.ncu-rep> <pairs per launch> <cells per launch> <source note>

bench.py reads the JSON (roofline.traffic, roofline.binding): DRAM bytes and executed lane-instructions per pair /
per cell of fast_emit_kernel, with the git hash of the build the capture was taken from.  The capture command is the
one in profiles/README.md: tools/k2_time.py with a known pair count and a single chunk, so one launch = that many
pairs."""
import csv
import json
import os
import subprocess
import sys

rep, pairs, cells, note = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(h)}


def val(r, k):
    v = float(r[idx[k]].replace(",", ""))
    u = units[idx[k]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


out = {"git": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
       "source": note, "report": os.path.relpath(rep), "pairs_per_launch": pairs, "cells_per_launch": cells}
for r in data:
    name = r[idx["Kernel Name"]]
    import re
    mh = re.search(r"fast_hist_kernel<[^,>]*4[^,>]*,\s*[^,>]*([01])\)?>", name) or re.search(r"tc_hist_kernel<[^>]*([01])\)?>", name)
    key = "emit" if ("fast_emit_kernel" in name or "tc_emit_kernel" in name) else ("hist_col" if mh.group(1) == "0" else "hist_row") if mh else None
    if key is None or key in out:
        continue
    winst = val(r, "smsp__inst_executed.sum")
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    out[key] = {"duration_ms_under_ncu": val(r, "gpu__time_duration.sum") / 1e6 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else val(r, "gpu__time_duration.sum"),
                "warp_inst_per_launch": winst, "lane_inst_per_cell": winst * 32 / cells,
                "dram_bytes_per_launch": dram, "dram_bytes_per_pair": dram / pairs,
                "dram_read_bytes_per_pair": val(r, "dram__bytes_read.sum") / pairs,
                "dram_write_bytes_per_pair": val(r, "dram__bytes_write.sum") / pairs,
                "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                "kernel": re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", ""),
                "registers": int(float(r[idx["launch__registers_per_thread"]]))}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "k2_constants.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
