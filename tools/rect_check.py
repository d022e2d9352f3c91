"""Rect/x-Sobel stage parity + timing (developer tool, run under gpurun)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402
from oracle_py import Oracle  # noqa: E402

o = Oracle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
for (W, H, D) in [(640, 480, 64), (1242, 375, 128), (1920, 1080, 256)]:
    nn = n if W == 640 else max(4, n * 640 * 480 // (W * H))
    rp = u.SHIPPED_RECT_PARAMS if W == 640 else u.identity_rect_params(W, H, float(W))
    L, R = u.synth_batch(1, 0, 4, W, H, D)
    reps = (nn + 3) // 4
    hL = np.concatenate([L] * reps)[:nn]; hR = np.concatenate([R] * reps)[:nn]
    fe = u.StereoFrontEnd(0, W, H, nn)
    fe.set_bm_params(width=W, height=H, profile=0, block_size=21, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128))
    fe.set_rect_params(rp)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    fe.set_profiling(True)
    dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
    ms = {"rect": [], "xsbl": [], "bm": []}
    for i in range(6):
        fe.submit_device("raw", i & 1, dL.data_ptr(), dR.data_ptr(), W, nn); b = fe.wait()
        if i >= 2:
            for k in ms:
                ms[k].append(fe.last_stage_ms(b)[k])
    gl, gr = fe.receive_rect(b)
    sl, sr = fe.receive_xsbl(b)
    j = nn - 1
    wl, wr = o.rectify(hL[j], rp, 0), o.rectify(hR[j], rp, 1)
    bad_r = int((gl[j] != wl).sum() + (gr[j] != wr).sum())
    bad_x = int((sl[j] != o.xsobel_rtl(wl)).sum() + (sr[j] != o.xsobel_rtl(wr)).sum())
    r, x, bm = (float(np.median(ms[k])) for k in ("rect", "xsbl", "bm"))
    gb = 4.0 * W * H * nn / 1e9
    print(f"{W}x{H} n={nn}: rect {r:.3f} ms ({gb / r * 1e3:.0f} GB/s)  xsbl {x:.3f} ms ({gb / x * 1e3:.0f} GB/s)  bm {bm:.3f} ms"
          f"  mismatches rect={bad_r} xsbl={bad_x}", flush=True)
    fe.close()
