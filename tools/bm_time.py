"""Quick BM timing + parity spot-check (developer tool, run under gpurun)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402
from oracle_py import Oracle  # noqa: E402

o = Oracle()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cfgs = [(640, 480, 64, 21, 0), (640, 480, 64, 15, 0), (640, 480, 64, 9, 0), (640, 480, 64, 21, 1)]
if len(sys.argv) > 2 and sys.argv[2] == "all":
    cfgs += [(1242, 375, 128, 15, 0), (1242, 375, 128, 15, 1), (1242, 375, 128, 21, 0), (1920, 1080, 256, 21, 0), (1920, 1080, 256, 21, 1)]
for (W, H, D, B, prof) in cfgs:
    n = nb if W == 640 else (nb // 2 if W == 1242 else max(8, nb * 32 // 296))
    L, R = u.synth_batch(1, 0, 4, W, H, D)
    reps = (n + 3) // 4
    hL = np.concatenate([L] * reps)[:n]; hR = np.concatenate([R] * reps)[:n]
    fe = u.StereoFrontEnd(0, W, H, n)
    if prof == 0:
        fe.set_bm_params(width=W, height=H, profile=0, block_size=B, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128),
                         uni_enable=0)
    else:
        fe.set_bm_params(width=W, height=H, profile=1, block_size=B, num_disparities=D, prefilter_cap=31, texture_threshold=10,
                         uniqueness_ratio=10)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    fe.set_profiling(True)
    dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
    for i in range(3):
        fe.submit_device("rect", i & 1, dL.data_ptr(), dR.data_ptr(), W, n); fe.wait()
    ms = []
    for i in range(8):
        fe.submit_device("rect", i & 1, dL.data_ptr(), dR.data_ptr(), W, n); b = fe.wait()
        ms.append(fe.last_stage_ms(b)["bm"])
    d = fe.receive_disp(b)
    if prof == 0:
        want = o.bm_rtl(o.xsobel_rtl(L[1]), o.xsobel_rtl(R[1]), wsz=B, ndisp=D, rtl_extended=int(D > 128))
    else:
        want = o.bm_cv(o.xsobel_cv(L[1]), o.xsobel_cv(R[1]), wsz=B, ndisp=D)
    j = ((n - 1) // 4) * 4 + 1 if ((n - 1) // 4) * 4 + 1 < n else 1
    bad = int((d[1] != want).sum()) + int((d[j] != want).sum())
    m = float(np.median(ms))
    print(f"{W}x{H} D{D} B{B} prof{prof} n={n}: bm {m:.3f} ms -> {n / m * 1e3:9.0f} fps, {n * W * H * D / m / 1e9:7.3f} Tpxd/s, "
          f"roofline(6op/18.4T) {6 * n * W * H * D / m / 1e9 / 18.4:.3f}  mismatches={bad}", flush=True)
    fe.close()
