"""Post-filter timing (developer tool, gpurun): OPENCV profile + validateDisparity + filterSpeckles on 256 synthetic pairs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H, D = 640, 480, 64
L, R = u.synth_batch(1, 0, 16, W, H, D)
hL = np.concatenate([L] * ((n + 15) // 16))[:n]; hR = np.concatenate([R] * ((n + 15) // 16))[:n]
fe = u.StereoFrontEnd(0, W, H, n)
fe.set_bm_params(width=W, height=H, profile=1, block_size=21, num_disparities=D, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10,
                 disp12_max_diff=1, speckle_window_size=50, speckle_range=32)
fe.set_stream(torch.cuda.current_stream().cuda_stream)
fe.set_profiling(True)
dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
for i in range(6):
    fe.submit_device("rect", i & 1, dL.data_ptr(), dR.data_ptr(), W, n); b = fe.wait()
    st = fe.last_stage_ms_ex(b)
print({k: round(v, 4) for k, v in st.items()})
d = fe.receive_disp(b)
if len(sys.argv) > 2:
    from oracle_py import Oracle
    o = Oracle()
    bad = 0
    for i in (0, 5, n - 1):
        want = o.bm_cv_post(o.xsobel_cv(hL[i]), o.xsobel_cv(hR[i]), wsz=21, ndisp=D)
        bad += int((d[i] != want).sum())
    print("mismatches vs oracle:", bad)
fe.close()
