"""Regenerates tests/golden/*.npz from the reference's own assets.

Run in the authoring container only (needs /root/reference and cv2):
    python tests/golden/make_golden.py

Outputs
  ref_rect_xsbl.npz : the reference's golden vectors data/ref_rect_{l,r}.zip and
                      data/ref_xsbl_{l,r}.zip (text hex, 480 lines x 640 "%02X "
                      tokens, written by src/dvp/sim/sim_dvp.v:846-927) as uint8.
  cv2_bm_golden.npz : cv2.StereoBM (4.13.0, the library the reference's CPU mode
                      calls at src/slam/src/core/main.cpp:197-217) outputs on
                      ref_rect for several parameter sets, post filters off
                      (speckleWindowSize=0, disp12MaxDiff=-1) and, for the
                      main.cpp parameter set, also with them on.
  dat_sha256.json   : SHA-256 and length of the four reference .dat text files (pins write_dat).
  rect_remap_ref.npz: output of the reference's own rect_remap() (fpga.c:303-366,
                      compiled from where it lies into oracle/_ref) for the
                      shipped parameter set (fpga.c:190-226).
"""
import ctypes
import io
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, ".."))


def read_dat(name):
    with zipfile.ZipFile(os.path.join(REF, "data", name + ".zip")) as z:
        txt = z.read(z.namelist()[0]).decode("ascii")
    rows = [[int(t, 16) for t in line.split()] for line in txt.splitlines() if line.strip()]
    a = np.array(rows, dtype=np.uint8)
    assert a.shape == (480, 640), a.shape
    return a


def dat_digests():
    """SHA-256 of the reference's .dat text files: pins the byte-exact writer u96_slam_b200.formats.write_dat."""
    import hashlib
    import json
    d = {}
    for name in ("ref_rect_l", "ref_rect_r", "ref_xsbl_l", "ref_xsbl_r"):
        with zipfile.ZipFile(os.path.join(REF, "data", name + ".zip")) as z:
            b = z.read(z.namelist()[0])
        d[name] = {"sha256": hashlib.sha256(b).hexdigest(), "bytes": len(b)}
    json.dump(d, open(os.path.join(HERE, "dat_sha256.json"), "w"), indent=1)


def main():
    import cv2
    dat_digests()

    rl, rr = read_dat("ref_rect_l"), read_dat("ref_rect_r")
    xl, xr = read_dat("ref_xsbl_l"), read_dat("ref_xsbl_r")
    np.savez_compressed(os.path.join(HERE, "ref_rect_xsbl.npz"), rect_l=rl, rect_r=rr, xsbl_l=xl, xsbl_r=xr)

    out = {}
    for (D, B, tex, uniq) in [(64, 21, 10, 10), (64, 15, 10, 10), (128, 9, 10, 15), (64, 21, 0, 0), (32, 5, 0, 0)]:
        bm = cv2.StereoBM_create(D, B)
        bm.setPreFilterCap(31); bm.setMinDisparity(0)
        bm.setTextureThreshold(tex); bm.setUniquenessRatio(uniq)
        bm.setSpeckleWindowSize(0); bm.setSpeckleRange(0); bm.setDisp12MaxDiff(-1)
        out[f"D{D}_B{B}_T{tex}_U{uniq}"] = bm.compute(rl, rr)
    # main.cpp:198-212 exactly (post filters on)
    bm = cv2.StereoBM_create(16, 9)
    bm.setPreFilterCap(31); bm.setBlockSize(21); bm.setMinDisparity(0); bm.setNumDisparities(64)
    bm.setTextureThreshold(10); bm.setUniquenessRatio(10)
    bm.setSpeckleWindowSize(50); bm.setSpeckleRange(32); bm.setDisp12MaxDiff(1)
    out["maincpp_postfilter"] = bm.compute(rl, rr)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "cv2_bm_golden.npz"), **out)

    # reference rect_remap() from oracle/_ref
    from oracle_py import RefFpga, SHIPPED_RECT
    ref = RefFpga()
    m = ref.rect_remap(SHIPPED_RECT, 640, 480)
    np.savez_compressed(os.path.join(HERE, "rect_remap_ref.npz"),
                        xs_l=m[0][0], ys_l=m[0][1], xs_r=m[1][0], ys_r=m[1][1])
    # the reference's shipped rectifier command dump (src/dvp/sim/cmd.dat, read by the testbench sim_dvp.v:174)
    words = np.array([int(t, 16) for t in open(os.path.join(REF, "src", "dvp", "sim", "cmd.dat")).read().split()], np.uint32)
    np.savez_compressed(os.path.join(HERE, "rect_cmd_dat.npz"), words=words)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
