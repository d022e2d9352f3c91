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
  reproject_ref.npz : outputs of the reference's own 3-D code (Stereo.cpp,
                      StereoCameraModel.cpp, Transform.cpp compiled unmodified into
                      oracle/_ref/libstereo_ref.so): the KITTI calib loader with and
                      without the 640x480 rescale, generateKeypoints3DStereo on 2000
                      keypoints with several depth gates, and the dense consumer
                      (main.cpp:522-551) with localTransform and a pose.
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


def reproject_inputs():
    """Deterministic inputs of the reprojection fixture (shared with the tests)."""
    rng = np.random.default_rng(96)
    W, H = 160, 120
    disp = rng.integers(-40, 64 * 16, (H, W)).astype(np.int16)
    disp[rng.random((H, W)) < 0.25] = -1                     # RTL invalid
    disp[rng.random((H, W)) < 0.05] = -16                    # OpenCV invalid
    disp[rng.random((H, W)) < 0.03] = 0
    disp[0, :6] = [1, 32767, -32768, 16, 15, 0]
    uv = np.stack([rng.random(2000) * W, rng.random(2000) * H], 1).astype(np.float32)
    uv[:6] = [[0, 0], [W - 0.01, H - 0.01], [0.5, 0.5], [-0.5, -0.5], [5.0, 0.0], [1.75, 0.25]]
    P_l = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)
    P_r = P_l.copy(); P_r[0, 3] = -386.1448
    pose = np.array([0.9, 0.1, -0.2, 1.5, -0.1, 0.95, 0.05, -2.25, 0.2, -0.04, 0.97, 0.3], np.float32)
    gates = [(0.0, 0.0), (-1.0, 0.0), (2.0, 30.0), (0.0, 5.0)]
    return W, H, disp, uv, P_l, P_r, pose, gates


def reproject_golden():
    import tempfile
    from oracle_py import RefStereo
    r = RefStereo()
    W, H, disp, uv, P_l, P_r, pose, gates = reproject_inputs()
    calib = os.path.join(tempfile.mkdtemp(), "calib.txt")
    r.write_kitti_calib(calib, P_l, P_r)
    out = {"local_transform": r.local_transform()[0]}
    for rs in (0, 1):
        Pl, Pr = r.model_load(calib, rs)
        out[f"P_l_resize{rs}"] = Pl; out[f"P_r_resize{rs}"] = Pr
        for gi, (mn, mx) in enumerate(gates):
            out[f"kp_resize{rs}_gate{gi}"] = r.keypoints3d(calib, rs, uv, disp, mn, mx)
        dd = np.ascontiguousarray(disp[::4, ::4][:H // 4, :W // 4])
        out[f"dense_resize{rs}_local"] = r.dense_cloud(calib, rs, dd, 4, True, None)
        out[f"dense_resize{rs}_local_pose"] = r.dense_cloud(calib, rs, dd, 4, True, pose)
        if rs == 0:
            out["dense_resize0_plain"] = r.dense_cloud(calib, rs, np.ascontiguousarray(disp), 1, False, None)
    np.savez_compressed(os.path.join(HERE, "reproject_ref.npz"), **out)


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
    reproject_golden()
    print("golden fixtures written")


if __name__ == "__main__":
    main()
