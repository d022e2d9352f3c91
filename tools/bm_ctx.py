"""Why does the BM stage take longer inside the full pipeline?  Times the BM stage of C2 for (entry point, frame pool) pairs (developer tool)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402

W, H, D, B, n = 640, 480, 64, 21, 296
for kind in ("rect", "raw"):
    for pool in (4, 16):
        L, R = u.synth_batch(1, 0, pool, W, H, D)
        reps = (n + pool - 1) // pool
        hL = np.concatenate([L] * reps)[:n]; hR = np.concatenate([R] * reps)[:n]
        fe = u.StereoFrontEnd(0, W, H, n)
        fe.set_bm_params(width=W, height=H, profile=0, block_size=B, num_disparities=D, x_store_offset=1)
        fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
        fe.set_stream(torch.cuda.current_stream().cuda_stream)
        fe.set_profiling(True)
        dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
        ms = []
        for i in range(8):
            fe.submit_device(kind, i & 1, dL.data_ptr(), dR.data_ptr(), W, n); b = fe.wait()
            ms.append(fe.last_stage_ms(b)["bm"])
        d = fe.receive_disp(b)
        print(f"{kind} pool {pool}: bm {np.median(ms[3:]):.3f} ms   valid px/frame {int((d[0] >= 0).sum())}  d==0-ish {int((d[0] == -1).sum())}", flush=True)
        fe.close()
