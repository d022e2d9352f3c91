"""Write profiles/bm_traffic.json from an `ncu --set full` capture of the BM kernel (developer tool).
usage: make_traffic.py report.ncu-rep workload frames_per_launch
The file records a SHA-256 over the CUDA sources the capture was taken with; bench.py reports `roofline.traffic` only when that
equals the sources of the library it is running (a capture of other code is stale by construction)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, unit, v = rows[0], rows[1], rows[2]


def val(name):
    i = h.index(name)
    x = float(v[i].replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[i]]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
sys.path.insert(0, ROOT)
from u96_slam_b200 import build as _build  # noqa: E402
d = {"workload": workload, "kernel": v[h.index("Kernel Name")], "frames_per_launch": frames, "dram_bytes_read": int(rd),
     "dram_bytes_write": int(wr), "dram_bytes_per_frame": int((rd + wr) / frames),
     "src_sha16": _build.source_sha16(),
     "source": f"{os.path.basename(rep)}: ncu --set full --clock-control none, {frames} frames per launch"}
json.dump(d, open(os.path.join(ROOT, "profiles", "bm_traffic.json"), "w"), indent=1)
print(d)
