"""Aggregate an `ncu --page source --csv --print-source sass` export by code region (developer tool).
usage: ncu_regions.py file.csv [name:hexstart:hexend ...]   (addresses relative to the kernel start; no regions = 16 equal chunks)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(body[0][0], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[ix["# Samples"]]) for r in body)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
regs = [a.split(":") for a in sys.argv[2:]]
if not regs:
    n = len(body)
    regs = [(f"c{k}", hex(16 * (k * n // 16)), hex(16 * ((k + 1) * n // 16))) for k in range(16)]
print(f"total samples {tot_s}, warp instructions {tot_i}")
for name, a, b in regs:
    a, b = int(a, 16), int(b, 16)
    sel = [r for r in body if a <= int(r[0], 16) - base < b]
    s = sum(int(r[ix["# Samples"]]) for r in sel)
    i = sum(int(r[ix["Instructions Executed"]]) for r in sel)
    wf = sum(int(r[ix["L1 Wavefronts Shared"]]) for r in sel)
    st = {n: sum(int(r[ix[n]]) for r in sel) for n in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print(f"{name:10s} samples {100.0 * s / tot_s:5.1f}%  inst {100.0 * i / tot_i:5.1f}% ({i})  smem wavefronts {wf}  " +
          " ".join(f"{k[6:]}={100.0 * v / max(s, 1):.0f}%" for k, v in top))
