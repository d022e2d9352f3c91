"""Run the BM stage of one configuration a few times (developer tool: the target of ncu captures).
usage: bm_one.py W H D B profile n [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402

W, H, D, B, prof, n = (int(v) for v in sys.argv[1:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 4
L, R = u.synth_batch(1, 0, 4, W, H, D)
k = (n + 3) // 4
hL = np.concatenate([L] * k)[:n]; hR = np.concatenate([R] * k)[:n]
fe = u.StereoFrontEnd(0, W, H, n)
if prof == 0:
    fe.set_bm_params(width=W, height=H, profile=0, block_size=B, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128), uni_enable=0)
else:
    fe.set_bm_params(width=W, height=H, profile=1, block_size=B, num_disparities=D, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10)
fe.set_stream(torch.cuda.current_stream().cuda_stream)
fe.set_profiling(True)
dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
ms = []
for i in range(reps):
    fe.submit_device("rect", i & 1, dL.data_ptr(), dR.data_ptr(), W, n); b = fe.wait()
    ms.append(fe.last_stage_ms(b)["bm"])
print(f"{W}x{H} D{D} B{B} prof{prof} n={n}: bm {min(ms):.3f} ms")
fe.close()
