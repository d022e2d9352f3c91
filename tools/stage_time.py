"""Stage timings of the HBM-side kernels (rect remap, x-Sobel, GFTT) on device-resident batches (developer tool, gpurun).
Prints ms per launch and the fraction of the measured HBM copy peak for the algorithmic bytes of each stage."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402

peak = 6500.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
for (W, H, D) in ((640, 480, 64), (1242, 375, 128), (1920, 1080, 256)):
    nn = n if W == 640 else max(8, n * 640 * 480 // (W * H))
    L, R = u.synth_batch(1, 0, 4, W, H, D)
    reps = (nn + 3) // 4
    hL = np.concatenate([L] * reps)[:nn]; hR = np.concatenate([R] * reps)[:nn]
    fe = u.StereoFrontEnd(0, W, H, nn)
    fe.set_bm_params(width=W, height=H, profile=0, block_size=9, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128))
    fe.set_rect_params(u.SHIPPED_RECT_PARAMS if W == 640 else u.identity_rect_params(W, H))
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    fe.set_profiling(True)
    dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
    res = {}
    for gftt in (0, 1):
        fe.set_gftt(bool(gftt))
        for kind in ("raw", "rect"):
            ms = {"rect": [], "xsbl": []}
            for i in range(10):
                fe.submit_device(kind, i & 1, dL.data_ptr(), dR.data_ptr(), W, nn); b = fe.wait()
                if i >= 3:
                    t = fe.last_stage_ms(b)
                    ms["rect"].append(t["rect"]); ms["xsbl"].append(t["xsbl"])
            res[(gftt, kind)] = {k: float(np.median(v)) for k, v in ms.items()}
    rect = res[(0, "raw")]["rect"]; xs = res[(0, "rect")]["xsbl"]; gf = res[(1, "rect")]["rect"] - res[(0, "rect")]["rect"]
    px = nn * W * H
    print(f"{W}x{H} n={nn}: rect {rect:.4f} ms ({4 * px / rect / 1e6 / peak:.3f} of HBM {peak:.0f}), xsobel {xs:.4f} ms ({4 * px / xs / 1e6 / peak:.3f}), "
          f"gftt {gf:.4f} ms ({3 * px / gf / 1e6 / peak:.3f})", flush=True)
    fe.close()
