"""ctypes bindings for the CPU oracle (oracle/_build/liboracle.so) and for the
reference's own rect_remap() (oracle/_ref/libfpga_ref.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libfpga_ref.so")
REF_STEREO_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libstereo_ref.so")
# StereoCameraModel.cpp:9-14
LOCAL_TRANSFORM = np.array([0, 0, 1, 0, -1, 0, 0, 0, 0, -1, 0, 0], np.float32)

# shipped rectification parameter set, src/StereoBM/src/fpga.c:190-226
SHIPPED_RECT = dict(
    f=[[40419817, 40382910], [39609530, 39627967]],
    c=[320, 240],
    f2inv=[6338213, 6338213],
    c2_f2=[4984405, 5932596],
    rot=[
        [[16598538, -120818, 2439034], [137992, 16776300, -108069], [-2438123, 126979, 16598626]],
        [[16569087, -69780, 2633522], [51223, 16776692, 122251], [-2633948, -112694, 16568783]],
    ],
)


class RectParams(ctypes.Structure):
    _fields_ = [("f", (ctypes.c_int32 * 2) * 2), ("c", ctypes.c_int32 * 2),
                ("f2inv", ctypes.c_int32 * 2), ("c2_f2", ctypes.c_int32 * 2),
                ("rot", ((ctypes.c_int32 * 3) * 3) * 2)]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for cam in range(2):
            for k in range(2):
                p.f[cam][k] = d["f"][cam][k]
            for i in range(3):
                for j in range(3):
                    p.rot[cam][i][j] = d["rot"][cam][i][j]
        for k in range(2):
            p.c[k] = d["c"][k]; p.f2inv[k] = d["f2inv"][k]; p.c2_f2[k] = d["c2_f2"][k]
        return p


class BmRtlParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("wsz", "ndisp", "uni_enb", "uni_mode", "uni_thr", "x_store_offset", "rtl_extended", "bitserial_div")]


class BmCvParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("wsz", "ndisp", "prefilter_cap", "texture_threshold", "uniqueness_ratio")]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            build_oracle()
        L = ctypes.CDLL(LIB_PATH)
        self.L = L
        u8p, i16p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int16)
        L.orc_diven.restype = ctypes.c_uint64
        L.orc_diven.argtypes = [ctypes.c_int] * 4 + [ctypes.c_uint64] * 2
        L.orc_rect_remap.argtypes = [ctypes.POINTER(RectParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, i16p, i16p]
        L.orc_rect_interp.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i16p, i16p, u8p]
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.orc_rect_remap32.argtypes = [ctypes.POINTER(RectParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p]
        L.orc_rect_interp32.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, u8p]
        L.orc_xsobel_rtl.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p]
        L.orc_xsobel_cv.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
        L.orc_bm_rtl.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(BmRtlParams), i16p]
        L.orc_bm_rtl_last_sat_events.restype = ctypes.c_int64
        L.orc_bm_cv.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(BmCvParams), i16p]
        L.orc_bm_cv_cost.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(BmCvParams), i16p, i16p]
        L.orc_validate_disparity.argtypes = [i16p, i16p] + [ctypes.c_int] * 5
        L.orc_filter_speckles.argtypes = [i16p] + [ctypes.c_int] * 5
        L.orc_reproject.argtypes = [i16p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_float)]

    def pack_uvc(self, mode, L=None, R=None, disp=None):
        """UVC payload (xusb_main.c:293-376): mode 1 rect / 2 xsbl (planar u8 L, R), 3 disparity (s16)"""
        u8p = ctypes.POINTER(ctypes.c_uint8); i16p = ctypes.POINTER(ctypes.c_int16)
        if mode == 3:
            disp = np.ascontiguousarray(disp, np.int16); H, W = disp.shape
            a = (None, None, disp.ctypes.data_as(i16p))
        else:
            L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8); H, W = L.shape
            a = (L.ctypes.data_as(u8p), R.ctypes.data_as(u8p), None)
        out = np.empty((H, 2 * W, 2), np.uint8)
        self.L.orc_pack_uvc.argtypes = [ctypes.c_int, u8p, u8p, i16p, ctypes.c_int, ctypes.c_int, u8p]
        self.L.orc_pack_uvc.restype = None
        self.L.orc_pack_uvc(mode, a[0], a[1], a[2], W, H, out.ctypes.data_as(u8p))
        return out

    def gftt_eig(self, src):
        """dvp/rtl/gftt*.v -> (H, W) u16 min-eigenvalue map, per-frame maximum (gftt.Max)"""
        a, p = _u8(src); H, W = a.shape
        out = np.empty((H, W), np.uint16); mx = ctypes.c_uint16(0)
        u16p = ctypes.POINTER(ctypes.c_uint16)
        self.L.orc_gftt_eig.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int, u16p, u16p]
        self.L.orc_gftt_eig.restype = None
        self.L.orc_gftt_eig(p, W, H, W, out.ctypes.data_as(u16p), ctypes.byref(mx))
        return out, int(mx.value)

    def diven(self, DW, VW, QW, MSB_INV, dividend, divisor):
        return int(self.L.orc_diven(DW, VW, QW, MSB_INV, dividend & (2**64 - 1), divisor & (2**64 - 1)))

    def rect_remap(self, params, lr, W, H):
        p = params if isinstance(params, RectParams) else RectParams.from_dict(params)
        xs = np.empty((H, W), np.int16); ys = np.empty((H, W), np.int16)
        i16p = ctypes.POINTER(ctypes.c_int16)
        self.L.orc_rect_remap(ctypes.byref(p), lr, W, H, xs.ctypes.data_as(i16p), ys.ctypes.data_as(i16p))
        return xs, ys

    def rect_interp(self, src, xs, ys):
        src, sp = _u8(src)
        H, W = src.shape
        xs = np.ascontiguousarray(xs, np.int16); ys = np.ascontiguousarray(ys, np.int16)
        dst = np.empty((H, W), np.uint8)
        i16p = ctypes.POINTER(ctypes.c_int16)
        self.L.orc_rect_interp(sp, W, H, W, xs.ctypes.data_as(i16p), ys.ctypes.data_as(i16p),
                               dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return dst

    def rect_interp32(self, src, params, lr):
        """RTL-extended rectification (32-bit coordinates, no short wrap)."""
        src, sp = _u8(src)
        H, W = src.shape
        p = params if isinstance(params, RectParams) else RectParams.from_dict(params)
        xs = np.empty((H, W), np.int32); ys = np.empty((H, W), np.int32)
        i32p = ctypes.POINTER(ctypes.c_int32)
        self.L.orc_rect_remap32(ctypes.byref(p), lr, W, H, xs.ctypes.data_as(i32p), ys.ctypes.data_as(i32p))
        dst = np.empty((H, W), np.uint8)
        self.L.orc_rect_interp32(sp, W, H, W, xs.ctypes.data_as(i32p), ys.ctypes.data_as(i32p),
                                 dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return dst

    def rectify(self, src, params, lr):
        """Reference semantics (short map) inside the RTL counter widths, RTL-extended beyond."""
        H, W = src.shape
        if W > 1023 or H > 511:
            return self.rect_interp32(src, params, lr)
        xs, ys = self.rect_remap(params, lr, W, H)
        return self.rect_interp(src, xs, ys)

    def xsobel_rtl(self, src):
        src, sp = _u8(src)
        H, W = src.shape
        dst = np.empty((H, W), np.uint8)
        self.L.orc_xsobel_rtl(sp, W, H, dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return dst

    def xsobel_cv(self, src, cap=31):
        src, sp = _u8(src)
        H, W = src.shape
        dst = np.empty((H, W), np.uint8)
        self.L.orc_xsobel_cv(sp, W, H, cap, dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return dst

    def bm_rtl(self, xl, xr, wsz=21, ndisp=64, uni_enb=0, uni_mode=0, uni_thr=0,
               x_store_offset=1, rtl_extended=0, bitserial_div=1):
        xl, lp = _u8(xl); xr, rp = _u8(xr)
        H, W = xl.shape
        p = BmRtlParams(wsz, ndisp, uni_enb, uni_mode, uni_thr, x_store_offset, rtl_extended, bitserial_div)
        disp = np.empty((H, W), np.int16)
        rc = self.L.orc_bm_rtl(lp, rp, W, H, ctypes.byref(p), disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
        if rc != 0:
            raise ValueError(f"orc_bm_rtl rc={rc}")
        return disp

    def sat_events(self):
        return int(self.L.orc_bm_rtl_last_sat_events())

    def bm_cv(self, pl, pr, wsz=21, ndisp=64, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10):
        pl, lp = _u8(pl); pr, rp = _u8(pr)
        H, W = pl.shape
        p = BmCvParams(wsz, ndisp, prefilter_cap, texture_threshold, uniqueness_ratio)
        disp = np.empty((H, W), np.int16)
        rc = self.L.orc_bm_cv(lp, rp, W, H, ctypes.byref(p), disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
        if rc != 0:
            raise ValueError(f"orc_bm_cv rc={rc}")
        return disp

    def bm_cv_post(self, pl, pr, wsz=21, ndisp=64, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10,
                   disp12_max_diff=1, speckle_window=50, speckle_range=32):
        """cv::StereoBM as configured at main.cpp:198-212: BM, then validateDisparity, then filterSpeckles."""
        pl, lp = _u8(pl); pr, rp = _u8(pr)
        H, W = pl.shape
        p = BmCvParams(wsz, ndisp, prefilter_cap, texture_threshold, uniqueness_ratio)
        disp = np.empty((H, W), np.int16); cost = np.zeros((H, W), np.int16)
        i16p = ctypes.POINTER(ctypes.c_int16)
        rc = self.L.orc_bm_cv_cost(lp, rp, W, H, ctypes.byref(p), disp.ctypes.data_as(i16p), cost.ctypes.data_as(i16p))
        if rc != 0:
            raise ValueError(f"orc_bm_cv_cost rc={rc}")
        if disp12_max_diff >= 0:
            self.L.orc_validate_disparity(disp.ctypes.data_as(i16p), cost.ctypes.data_as(i16p), W, H, 0, ndisp, disp12_max_diff)
        if speckle_window > 0 and speckle_range >= 0:
            self.L.orc_filter_speckles(disp.ctypes.data_as(i16p), W, H, -16, speckle_window, speckle_range)
        return disp

    def reproject(self, disp, P_l, P_r, decim=1, apply_local=0):
        disp = np.ascontiguousarray(disp, np.int16)
        H, W = disp.shape
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        out = np.empty((H // decim, W // decim, 3), np.float32)
        dp = ctypes.POINTER(ctypes.c_double)
        self.L.orc_reproject(disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), W, H, Pl.ctypes.data_as(dp),
                             Pr.ctypes.data_as(dp), decim, apply_local, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return out

    def reproject_ex(self, disp, P_l, P_r, decim=1, local_T=None, pose=None):
        """dense consumer of main.cpp:522-551 with explicit 3x4 float transforms (None = skipped)"""
        disp = np.ascontiguousarray(disp, np.int16)
        H, W = disp.shape
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        out = np.empty((H // decim, W // decim, 3), np.float32)
        dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
        lt = None if local_T is None else np.ascontiguousarray(local_T, np.float32).reshape(12)
        ps = None if pose is None else np.ascontiguousarray(pose, np.float32).reshape(12)
        self.L.orc_reproject_ex.argtypes = [ctypes.POINTER(ctypes.c_int16), ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int, fp, fp, fp]
        self.L.orc_reproject_ex.restype = None
        self.L.orc_reproject_ex(disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), W, H, Pl.ctypes.data_as(dp), Pr.ctypes.data_as(dp),
                                decim, None if lt is None else lt.ctypes.data_as(fp), None if ps is None else ps.ctypes.data_as(fp),
                                out.ctypes.data_as(fp))
        return out

    def reproject_points(self, disp, P_l, P_r, uv, min_depth=0.0, max_depth=0.0, local_T=LOCAL_TRANSFORM, mask=None):
        """generateKeypoints3DStereo (Stereo.cpp:53-117) on a dense 16x disparity map; uv = (n, 2) float32 (x, y)"""
        disp = np.ascontiguousarray(disp, np.int16)
        H, W = disp.shape
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        n = uv.shape[0]
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        out = np.empty((n, 3), np.float32)
        dp, fp, u8p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
        lt = None if local_T is None else np.ascontiguousarray(local_T, np.float32).reshape(12)
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.L.orc_reproject_points.argtypes = [ctypes.POINTER(ctypes.c_int16), ctypes.c_int, ctypes.c_int, dp, dp, fp, ctypes.c_int, u8p,
                                                ctypes.c_float, ctypes.c_float, fp, fp]
        self.L.orc_reproject_points.restype = None
        self.L.orc_reproject_points(disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), W, H, Pl.ctypes.data_as(dp), Pr.ctypes.data_as(dp),
                                    uv.ctypes.data_as(fp), n, None if mk is None else mk.ctypes.data_as(u8p), min_depth, max_depth,
                                    None if lt is None else lt.ctypes.data_as(fp), out.ctypes.data_as(fp))
        return out


# ---- the reference's own 3-D code (Stereo.cpp, StereoCameraModel.cpp, Transform.cpp), compiled from /root/reference ----
DEPTH_METHOD_CV_BM, DEPTH_METHOD_FPGA_BM = 2, 4        # enum DEPTH_METHOD, slam/include/core/Parameters.h:25-31


class RefStereo:
    """oracle/_ref/libstereo_ref.so (oracle/Makefile target `ref`): the reference's sources unmodified, OpenCV containers stubbed.
    The camera model is loaded by the reference's own KITTI calib.txt reader, so every call takes the path of such a file."""

    def __init__(self):
        self.L = ctypes.CDLL(REF_STEREO_LIB_PATH)

    @staticmethod
    def write_kitti_calib(path, P_l, P_r):
        with open(path, "w") as f:
            for k, P in (("P0", P_l), ("P1", P_r)):
                f.write(f"{k}: " + " ".join(repr(float(v)) for v in np.asarray(P, np.float64).reshape(12)) + "\n")

    def model_load(self, calib, do_resize):
        out = (ctypes.c_double * 10)()
        rc = self.L.ref_model_load(calib.encode(), int(do_resize), out)
        if rc != 0:
            raise ValueError("StereoCameraModel::load failed")
        v = list(out)
        P_l = np.array([[v[0], 0, v[2], v[4]], [0, v[1], v[3], 0], [0, 0, 1, 0]], np.float64)
        P_r = np.array([[v[5], 0, v[7], v[9]], [0, v[6], v[8], 0], [0, 0, 1, 0]], np.float64)
        return P_l, P_r

    def local_transform(self):
        out = (ctypes.c_float * 12)()
        is_null = self.L.ref_local_transform(out)
        return np.array(list(out), np.float32), bool(is_null)

    def keypoints3d(self, calib, do_resize, uv, disp, min_depth=0.0, max_depth=0.0, depth_method=DEPTH_METHOD_FPGA_BM):
        disp = np.ascontiguousarray(disp, np.int16)
        H, W = disp.shape
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.empty((uv.shape[0], 3), np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        self.L.ref_keypoints3d.argtypes = [ctypes.c_char_p, ctypes.c_int, fp, ctypes.c_int, ctypes.POINTER(ctypes.c_int16), ctypes.c_int,
                                           ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
        rc = self.L.ref_keypoints3d(calib.encode(), int(do_resize), uv.ctypes.data_as(fp), uv.shape[0],
                                    disp.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), W, H, min_depth, max_depth, depth_method,
                                    out.ctypes.data_as(fp))
        if rc != 0:
            raise ValueError("StereoCameraModel::load failed")
        return out

    def dense_cloud(self, calib, do_resize, depth, scale, apply_local=True, pose=None):
        """main.cpp:522-551 over an already decimated map (SensorData.cpp:50-58) with dispScale = scale"""
        depth = np.ascontiguousarray(depth, np.int16)
        rows, cols = depth.shape
        out = np.empty((rows, cols, 3), np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        ps = None if pose is None else np.ascontiguousarray(pose, np.float32).reshape(12)
        self.L.ref_dense_cloud.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int16), ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, fp, fp]
        rc = self.L.ref_dense_cloud(calib.encode(), int(do_resize), depth.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), rows, cols,
                                    scale, int(apply_local), None if ps is None else ps.ctypes.data_as(fp), out.ctypes.data_as(fp))
        if rc != 0:
            raise ValueError("StereoCameraModel::load failed")
        return out


# ---- the reference's own rect_remap(), compiled from /root/reference into oracle/_ref ----
class _RefRectParamCh(ctypes.Structure):      # struct RECT_PARAM_CH, fpga.h:250-256 (LP64 host layout)
    _fields_ = [("f", ctypes.c_long * 2), ("c", ctypes.c_short * 2), ("f2inv", ctypes.c_long * 2),
                ("c2_f2", ctypes.c_long * 2), ("rot", (ctypes.c_long * 3) * 3)]


class _RefRectParam(ctypes.Structure):        # struct RECT_PARAM, fpga.h:258-260
    _fields_ = [("ch", _RefRectParamCh * 2)]


class _RefMat2S(ctypes.Structure):            # struct MAT2S, fpga.h:262-266
    _fields_ = [("rows", ctypes.c_int), ("cols", ctypes.c_int), ("data", ctypes.POINTER(ctypes.c_short) * 2)]


class RefFpga:
    """The reference's fpga.c compiled unmodified (oracle/Makefile target `ref`)."""

    def __init__(self):
        self.L = ctypes.CDLL(REF_LIB_PATH)
        self.L.rect_remap.argtypes = [ctypes.POINTER(_RefRectParam), ctypes.POINTER(_RefMat2S), ctypes.POINTER(_RefMat2S)]
        self.L.rect_remap.restype = None

    def rect_remap(self, d, W, H):
        p = _RefRectParam()
        for cam in range(2):
            ch = p.ch[cam]
            for k in range(2):
                ch.f[k] = d["f"][cam][k]; ch.c[k] = d["c"][k]
                ch.f2inv[k] = d["f2inv"][k]; ch.c2_f2[k] = d["c2_f2"][k]
            for i in range(3):
                for j in range(3):
                    ch.rot[i][j] = d["rot"][cam][i][j]
        maps, bufs = [], []
        for _ in range(2):
            m = _RefMat2S(); m.rows = H; m.cols = W
            a = [np.zeros((H, W), np.int16) for _ in range(2)]
            for k in range(2):
                m.data[k] = a[k].ctypes.data_as(ctypes.POINTER(ctypes.c_short))
            maps.append(m); bufs.append(a)
        self.L.rect_remap(ctypes.byref(p), ctypes.byref(maps[0]), ctypes.byref(maps[1]))
        self._last = (p, maps, bufs)
        return bufs

    def rect_cmd_stream(self, d, W, H):
        """rect_remap() + rect_cmd_gen() of the reference (fpga.c:303-605): the 18-bit words issue_cmd() writes to the
        rectifier's command port, captured by oracle/ref_stubs/capture_cmd.h."""
        self.rect_remap(d, W, H)
        p, maps, _ = self._last
        L = self.L
        L.set_rect_param.argtypes = [ctypes.POINTER(_RefRectParam)]; L.set_rect_param.restype = None
        L.rect_cmd_gen.argtypes = [_RefMat2S, _RefMat2S]; L.rect_cmd_gen.restype = None
        L.u96_ref_cmd_data.restype = ctypes.POINTER(ctypes.c_uint)
        L.set_rect_param(ctypes.byref(p))
        L.u96_ref_cmd_reset()
        L.rect_cmd_gen(maps[0], maps[1])
        n = L.u96_ref_cmd_count()
        return np.ctypeslib.as_array(L.u96_ref_cmd_data(), shape=(n,)).astype(np.uint32)
