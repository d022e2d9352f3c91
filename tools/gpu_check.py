"""Verbose GPU parity run (developer tool, run under gpurun): every stage through the C ABI vs the
oracle / golden fixtures, with mismatch diagnostics written to gpurun_out/gpu_check.log."""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import u96_slam_b200 as u  # noqa: E402
from u96_slam_b200.stereo import microbench  # noqa: E402
from oracle_py import Oracle  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "w")
FAILS = 0


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n"); LOG.flush()


def compare(name, got, want):
    global FAILS
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        log("FAIL", name, "shape", got.shape, want.shape); FAILS += 1; return False
    if got.dtype.kind == "f":
        bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    else:
        bad = got != want
    nb = int(bad.sum())
    if nb == 0:
        log("ok  ", name, got.shape)
        return True
    FAILS += 1
    idx = np.argwhere(bad)
    log("FAIL", name, "mismatches", nb, "of", bad.size, "first:", [tuple(int(v) for v in i) for i in idx[:6]])
    for i in idx[:6]:
        log("     at", tuple(int(v) for v in i), "got", got[tuple(i)], "want", want[tuple(i)])
    if got.ndim >= 2:
        rows = np.unique(idx[:, -2]); cols = np.unique(idx[:, -1])
        log("     rows", rows[:5], "..", rows[-5:], "cols", cols[:5], "..", cols[-5:])
    return False


def main():
    o = Oracle()
    g = np.load(os.path.join(ROOT, "tests/golden/ref_rect_xsbl.npz"))
    rl, rr, xl, xr = g["rect_l"], g["rect_r"], g["xsbl_l"], g["xsbl_r"]
    cvg = np.load(os.path.join(ROOT, "tests/golden/cv2_bm_golden.npz"))

    names = ["IADD3", "VABSDIFF4", "VIADDMNMX.U16x2", "VIMNMX3.U32", "PRMT", "IMAD", "LDS.128(GB/s)", "SHFL", "VIMNMX.U16x2", "LOP3"]
    for w, nme in enumerate(names):
        try:
            log("microbench", nme, "%.1f G/s" % microbench(w))
        except Exception as e:  # noqa: BLE001
            log("microbench", nme, "ERR", e)

    fe = u.StereoFrontEnd(0, 640, 480, 4)
    # ---- C1: bundled pair, RTL profile ----
    fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe.set_bm_params(x_store_offset=1)
    fe.submit_rect(0, rl, rr); b = fe.wait()
    sl, sr = fe.receive_xsbl(b)
    compare("xsbl_rtl L vs ref_xsbl", sl[0], xl); compare("xsbl_rtl R vs ref_xsbl", sr[0], xr)
    d = fe.receive_disp(b)[0]
    want = o.bm_rtl(xl, xr, wsz=21, ndisp=64)
    compare("bm_rtl wsz21 D64 (from rect)", d, want)
    log("crc %08x (survey appendix B: 3c312d26)" % (zlib.crc32(d.tobytes()) & 0xFFFFFFFF))
    for wsz, uni, thr, mode in ((21, 1, 921, 0), (15, 0, 0, 0), (9, 1, 800, 1), (5, 0, 0, 0), (31, 0, 0, 0), (17, 0, 0, 0)):
        fe.set_bm_params(block_size=wsz, uni_enable=uni, uni_thr=thr, uni_mode=mode)
        fe.submit_xsbl(1, xl, xr); b = fe.wait()
        compare(f"bm_rtl wsz{wsz} uni{uni}/{thr}/{mode}", fe.receive_disp(b)[0],
                o.bm_rtl(xl, xr, wsz=wsz, ndisp=64, uni_enb=uni, uni_thr=thr, uni_mode=mode))
    for D in (32, 128, 256):
        fe.set_bm_params(block_size=21, num_disparities=D, uni_enable=0, rtl_extended=int(D > 128))
        fe.submit_xsbl(0, xl, xr); b = fe.wait()
        compare(f"bm_rtl wsz21 D{D}", fe.receive_disp(b)[0], o.bm_rtl(xl, xr, wsz=21, ndisp=D, rtl_extended=int(D > 128)))
    # ---- OPENCV profile vs cv2 golden ----
    for k in cvg.files:
        if not k.startswith("D"):
            continue
        D, B, T, U = [int(s[1:]) for s in k.split("_")]
        fe.set_bm_params(profile=u.PROFILE_OPENCV, num_disparities=D, block_size=B, texture_threshold=T,
                         uniqueness_ratio=U, prefilter_cap=31)
        fe.submit_rect(0, rl, rr); b = fe.wait()
        pl, pr = fe.receive_xsbl(b)
        compare(f"xsbl_cv L {k}", pl[0], o.xsobel_cv(rl, 31))
        compare(f"bm_cv {k} vs cv2 golden", fe.receive_disp(b)[0], cvg[k])
    # ---- C2: raw pipeline on synthetic frames, batch of 3, both banks ----
    L, R = u.synth_batch(1, 0, 3, 640, 480, 64)
    fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
    fe.submit_raw(1, L, R); b = fe.wait()
    gl, gr = fe.receive_rect(b); sl, sr = fe.receive_xsbl(b); d = fe.receive_disp(b)
    for i in range(3):
        wl, wr = o.rectify(L[i], u.SHIPPED_RECT_PARAMS, 0), o.rectify(R[i], u.SHIPPED_RECT_PARAMS, 1)
        compare(f"rect L f{i}", gl[i], wl); compare(f"rect R f{i}", gr[i], wr)
        wxl, wxr = o.xsobel_rtl(wl), o.xsobel_rtl(wr)
        compare(f"xsbl L f{i}", sl[i], wxl)
        compare(f"disp raw f{i}", d[i], o.bm_rtl(wxl, wxr, wsz=21, ndisp=64))
    P_l = np.array([[500.0, 0, 320, 0], [0, 500, 240, 0], [0, 0, 1, 0]]); P_r = P_l.copy(); P_r[0, 3] = -60.0
    for decim in (1, 4):
        xyz = fe.reproject(b, P_l, P_r, decim, True)
        compare(f"reproject decim{decim}", xyz[0], o.reproject(d[0], P_l, P_r, decim, 1))
    fe.close()
    # ---- C3 shape ----
    fe = u.StereoFrontEnd(0, 1242, 375, 2)
    L, R = u.synth_batch(2, 0, 2, 1242, 375, 128)
    rp = u.identity_rect_params(1242, 375, 700.0)
    fe.set_rect_params(rp)
    for B in (9, 15, 21):
        fe.set_bm_params(width=1242, height=375, profile=u.PROFILE_RTL, block_size=B, num_disparities=128, x_store_offset=1)
        fe.submit_raw(0, L, R); b = fe.wait()
        gl, gr = fe.receive_rect(b); d = fe.receive_disp(b)
        wl, wr = o.rect_interp32(L[1], rp, 0), o.rect_interp32(R[1], rp, 1)
        compare(f"C3 rect L B{B}", gl[1], wl)
        compare(f"C3 disp rtl B{B}", d[1], o.bm_rtl(o.xsobel_rtl(wl), o.xsobel_rtl(wr), wsz=B, ndisp=128))
        fe.set_bm_params(profile=u.PROFILE_OPENCV, texture_threshold=10, uniqueness_ratio=10, prefilter_cap=31)
        fe.submit_rect(1, L, R); b = fe.wait()
        compare(f"C3 disp cv B{B}", fe.receive_disp(b)[1], o.bm_cv(o.xsobel_cv(L[1]), o.xsobel_cv(R[1]), wsz=B, ndisp=128))
    fe.close()
    # ---- C4 shape ----
    fe = u.StereoFrontEnd(0, 1920, 1080, 1)
    L, R = u.synth_pair(3, 0, 1920, 1080, 256)
    fe.set_bm_params(width=1920, height=1080, profile=u.PROFILE_RTL, block_size=21, num_disparities=256, rtl_extended=1, x_store_offset=1)
    fe.submit_rect(0, L, R); b = fe.wait()
    t = time.time()
    compare("C4 disp rtl", fe.receive_disp(b)[0], o.bm_rtl(o.xsobel_rtl(L), o.xsobel_rtl(R), wsz=21, ndisp=256, rtl_extended=1))
    log("oracle 1080p s", time.time() - t)
    fe.close()
    log("FAILS", FAILS)
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
