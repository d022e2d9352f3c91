"""Data formats either side of the hot path: the RTL testbench's hex dumps and the
calibration -> rectification-register conversion.  Host-side only; nothing here computes on images."""
import numpy as np


def write_dat(img):
    """Bytes of a `.dat` hex dump as the RTL testbench writes it (src/dvp/sim/sim_dvp.v:846-927: one "%02X " token
    per pixel, one line per image row).  The bundled data/ref_*.dat files show XSim's rendering: lower-case digits,
    CR LF line ends -- reproduced byte for byte (pinned by tests against the SHA-256 of the reference's files)."""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim != 2:
        raise ValueError("write_dat expects one 2-D u8 image")
    lut = np.array([f"{v:02x} ".encode("ascii") for v in range(256)], dtype="S3")
    rows = lut[a]                                     # (H, W) of 3-byte tokens
    return b"".join(r.tobytes() + b"\r\n" for r in rows)


def read_dat(data, width=None):
    """Inverse of write_dat (also the loader of sim_dvp.v:435-490 stimulus files): text -> (H, W) u8."""
    if isinstance(data, (bytes, bytearray)):
        data = data.decode("ascii")
    rows = [[int(t, 16) for t in line.split()] for line in data.splitlines() if line.strip()]
    a = np.array(rows, dtype=np.uint8)
    if width is not None and a.shape[1] != width:
        raise ValueError(f"expected {width} tokens per line, found {a.shape[1]}")
    return a


def rect_params_from_calibration(K_src, R_rect, K_new):
    """The 27 fixed-point rectification registers (struct RECT_PARAM, StereoBM/src/fpga.h:250-260; consumed by
    rect_remap(), fpga.c:303-366) from a pinhole calibration without lens distortion.

      K_src  : per camera (fx, fy, cx, cy) of the RAW images               [left, right]
      R_rect : per camera 3x3 rectifying rotation (cv::stereoRectify R1/R2) [left, right]
      K_new  : (fx', fy', cx', cy') of the common rectified camera

    The reference generates these numbers with a tool that is not in its repository (fpga.c:188); the formats are
    read off the consumer: f = source focal length u10.16; rot = R_rect in s0.24, applied transposed (= inverse
    rotation) by rect_remap; c = source principal point, integer pixels, and -- like f2inv = 2^32/f' and
    c2_f2 = c'/f' in u0.24 -- shared by both cameras (only ch[0]'s copies are used, fpga.c:295-300)."""
    fxn, fyn, cxn, cyn = (float(v) for v in K_new)
    p = dict(f=[], rot=[], c=[int(round(K_src[0][2])), int(round(K_src[0][3]))],
             f2inv=[int(round(2.0**32 / fxn)), int(round(2.0**32 / fyn))],
             c2_f2=[int(round(cxn / fxn * 2.0**24)), int(round(cyn / fyn * 2.0**24))])
    for cam in range(2):
        fx, fy = float(K_src[cam][0]), float(K_src[cam][1])
        p["f"].append([int(round(fx * 65536.0)), int(round(fy * 65536.0))])
        R = np.asarray(R_rect[cam], dtype=np.float64).reshape(3, 3)
        p["rot"].append([[int(round(R[i, j] * 2.0**24)) for j in range(3)] for i in range(3)])
    return p


# ---- projection-matrix files of the reference's camera model (slam/src/core/StereoCameraModel.cpp:19-122) --------------------
def _resize_projection(P, size, do_resize):
    """StereoCameraModel.cpp:108-119: scale fx, cx, Tx by 640/width and fy, cy, Ty by 480/height."""
    P = [np.array(p, np.float64).reshape(3, 4) for p in P]
    if do_resize:
        sx, sy = 640.0 / size[0], 480.0 / size[1]
        for p in P:
            p[0, 0] *= sx; p[0, 2] *= sx; p[0, 3] *= sx
            p[1, 1] *= sy; p[1, 2] *= sy; p[1, 3] *= sy
    return P[0], P[1]


def load_projection_kitti(path, do_resize=True):
    """KITTI odometry `calib.txt` (left/right combined, StereoCameraModel.cpp:71-104): lines `P0: 12 doubles`, `P1: 12 doubles`;
    the image size is not in the file and is assumed 1241x376 like the reference does.  -> (P_l, P_r) 3x4 float64, ready for
    u96_reproject."""
    P = {}
    for ln in open(path):
        k, _, rest = ln.partition(":")
        if k.strip() in ("P0", "P1"):
            v = [float(t) for t in rest.split()]
            if len(v) != 12:
                raise ValueError(f"{path}: {k.strip()} needs 12 values")
            P[k.strip()] = v
    if "P0" not in P or "P1" not in P:
        raise ValueError(f"{path}: P0/P1 not found")
    return _resize_projection([P["P0"], P["P1"]], (1241, 376), do_resize)


def load_projection_opencv_yml(path_left, path_right, do_resize=True):
    """OpenCV FileStorage YAML per camera (StereoCameraModel.cpp:32-68): `image_width`, `image_height` (left file) and
    `projection_matrix: !!opencv-matrix {rows: 3, cols: 4, dt: d, data: [...]}`.  -> (P_l, P_r)."""
    import re
    size, P = [0, 0], []
    for lr, path in enumerate((path_left, path_right)):
        txt = open(path).read()
        if lr == 0:
            for i, key in enumerate(("image_width", "image_height")):
                m = re.search(rf"^{key}\s*:\s*(\d+)", txt, re.M)
                if m:
                    size[i] = int(m.group(1))
        m = re.search(r"projection_matrix\s*:.*?rows\s*:\s*(\d+).*?cols\s*:\s*(\d+).*?data\s*:\s*\[(.*?)\]", txt, re.S)
        if not m:
            raise ValueError(f"{path}: projection_matrix not found")
        if (int(m.group(1)), int(m.group(2))) != (3, 4):
            raise ValueError(f"{path}: illegal projection matrix size ({m.group(1)},{m.group(2)})")
        P.append([float(t) for t in m.group(3).replace("\n", " ").split(",")])
    if do_resize and (size[0] <= 0 or size[1] <= 0):
        raise ValueError("image_width / image_height missing: cannot resize")
    return _resize_projection(P, size, do_resize)
