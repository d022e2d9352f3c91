"""Small end-to-end runs of every kernel for compute-sanitizer (developer tool, gpurun):
  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import u96_slam_b200 as u  # noqa: E402

CONFIGS = ((640, 96, 64, 21, 0, 0), (640, 96, 64, 15, 0, 0), (640, 96, 64, 21, 0, 1), (640, 96, 64, 15, 1, 0), (330, 80, 128, 9, 0, 0),
                                (330, 80, 128, 9, 1, 0), (530, 70, 256, 21, 0, 0), (530, 70, 256, 11, 1, 0), (200, 64, 96, 7, 0, 0), (332, 60, 64, 11, 1, 0))
rot = int(sys.argv[1]) if len(sys.argv) > 1 else 0                     # start with another configuration (first-launch effects)
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else len(CONFIGS)           # only the first cnt configurations (initcheck is slow)
for (W, H, D, B, prof, uni) in (CONFIGS[rot:] + CONFIGS[:rot])[:cnt]:
    n = 3
    L, R = u.synth_batch(5, 0, n, W, H, D)
    with u.StereoFrontEnd(0, W, H, n) as fe:
        if prof == 0:
            fe.set_bm_params(width=W, height=H, profile=0, block_size=B, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128),
                             uni_enable=uni, uni_thr=900)
        else:
            fe.set_bm_params(width=W, height=H, profile=1, block_size=B, num_disparities=D, prefilter_cap=31, texture_threshold=10,
                             uniqueness_ratio=10, disp12_max_diff=1, speckle_window_size=50, speckle_range=32)
        fe.set_rect_params(u.SHIPPED_RECT_PARAMS if W == 640 else u.identity_rect_params(W, H, float(W)))
        fe.set_gftt(True)
        fe.submit_raw(0, L, R)
        fe.wait()
        d = fe.receive_disp(0)
        fe.receive_rect(0); fe.receive_xsbl(0); fe.receive_eigen(0)
        P_l = np.array([[370.0, 0, 313.0, 0], [0, 917.0, 236.0, 0], [0, 0, 1, 0]]); P_r = P_l.copy(); P_r[0, 3] = -199.0
        fe.reproject_ex(0, P_l, P_r, 4, u.LOCAL_TRANSFORM, np.tile(np.eye(3, 4, dtype=np.float32).reshape(1, 12), (n, 1)))
        fe.reproject_points(0, P_l, P_r, np.array([[1.5, 2.5], [W - 1.0, H - 1.0], [-3.0, 0.0]], np.float32), 1)
        fe.receive_uvc(0, u.UVC_BM)
        print(W, H, D, B, prof, uni, "checksum", int(d.astype(np.int64).sum()), flush=True)
