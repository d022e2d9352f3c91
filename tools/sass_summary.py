"""Opcode histogram per kernel of libu96stereo.so (developer tool): `cuobjdump -sass` -> profiles/sass_summary.txt.
Shows which machine instructions the claims in DESIGN.md rest on (UTMALDG = TMA tensor load, SYNCS = mbarrier,
STAS = st.async to a peer CTA's shared memory, VABSDIFF4 / VIADDMNMX / VIMNMX = packed integer SIMD, IDP = dp2a/dp4a, HADD2 / HFMA2 = the fp16 forms of the saturating subtract and the abs-diff of the fused BM kernel, REDUX)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "u96_slam_b200", "lib", "libu96stereo.so")
KEY = ["UTMALDG", "SYNCS", "STAS", "VABSDIFF4", "VIADDMNMX", "VIMNMX3", "VIMNMX", "HADD2", "HFMA2", "IDP", "PRMT", "SHF", "REDUX", "LDS", "STS", "LDG", "STG",
       "ATOMS", "ATOMG", "RED", "BAR", "SHFL", "MUFU", "FFMA", "IMAD", "DFMA", "DMUL", "DADD", "HMMA", "UTCHMMA", "MEMBAR"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
        hist[kern]["_full_" + m.group(1)] += 1
lines = [f"libu96stereo.so sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]}  cubins: {', '.join(sorted(arch))}  kernels: {len(hist)}",
         "opcode counts per kernel (static SASS; base mnemonic, modifiers stripped).  HMMA / UTCHMMA = tensor-core MMA: none, by design.", ""]
tot = collections.Counter()
for k, c in hist.items():
    n = sum(v for kk, v in c.items() if not kk.startswith("_full_"))
    keys = " ".join(f"{kk}={c[kk]}" for kk in KEY if c.get(kk))
    lines.append(f"{k[:110]:110s} {n:6d} instr | {keys}")
    for kk in KEY:
        tot[kk] += c.get(kk, 0)
    for kk, v in c.items():
        if kk.startswith("_full_") and any(s in kk for s in ("U16x2", "U8", "2A", "4A", "TRANS64", "3D", "E.128", "E.64")):
            tot[kk[6:]] += v
lines += ["", "whole library: " + " ".join(f"{k}={v}" for k, v in tot.items() if v and not any(ch.islower() for ch in k) and "." not in k),
          "selected full mnemonics: " + " ".join(f"{k}={v}" for k, v in sorted(tot.items()) if "." in k)]
path = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "sass_summary.txt")
open(path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:3] + lines[-2:]))
