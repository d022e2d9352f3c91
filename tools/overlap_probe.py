"""Does overlapping rect/x-Sobel of one half-batch with the BM kernel of the other pay?  (developer tool)
One 296-frame submit per step on one stream  vs  two 148-frame submits on the two bank streams."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import u96_slam_b200 as u  # noqa: E402

W, H, D, B, n = 640, 480, 64, 21, 296
L, R = u.synth_batch(1, 0, 16, W, H, D)
hL = np.concatenate([L] * 19)[:n]; hR = np.concatenate([R] * 19)[:n]
dL, dR = torch.from_numpy(hL).cuda(), torch.from_numpy(hR).cuda()
fe = u.StereoFrontEnd(0, W, H, n)
fe.set_bm_params(width=W, height=H, profile=0, block_size=B, num_disparities=D, x_store_offset=1)
fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
half = n // 2
off = half * W * H


def serial(k):
    for i in range(k):
        fe.submit_device("raw", 0, dL.data_ptr(), dR.data_ptr(), W, n); fe.wait()


def split(k, parts):
    sz = n // parts
    for i in range(k):
        for p in range(parts):
            # two banks only: alternate, waiting for the bank's previous part
            b = p & 1
            if p >= 2:
                fe.wait()
            fe.submit_device("raw", b, dL.data_ptr() + p * sz * W * H, dR.data_ptr() + p * sz * W * H, W, sz)
        for _ in range(min(parts, 2)):
            fe.wait()


for name, fn in (("serial 296", lambda k: serial(k)), ("2 x 148 on two bank streams", lambda k: split(k, 2)),
                 ("4 x 74 alternating banks", lambda k: split(k, 4))):
    fn(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(20); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"{name}: {dt * 1e3:.3f} ms per 296 frames -> {n / dt:.0f} frames/s", flush=True)
fe.close()
