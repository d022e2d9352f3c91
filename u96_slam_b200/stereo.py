"""Host-side mirror of the reference interface over the C ABI (include/u96_stereo.h).

  StereoFrontEnd : thin object wrapper of the u96_* entry points (one per GPU).
  Fpga           : same data-plane method names as the reference's `class Fpga`
                   (slam/include/core/FPGA.h:347-397): setRectImage, receiveRectImages,
                   receiveDepthMap, receiveData -- so slam-style callers read the same.
  StereoBM       : same surface as the cv::StereoBM object the reference's CPU mode builds at
                   slam/src/core/main.cpp:198-215 (create / set* / compute).

ctypes only; numpy for host buffers.  No torch types cross the boundary.  If libu96stereo.so is
missing or no CUDA device is present this module raises -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
PROFILE_RTL, PROFILE_OPENCV = 0, 1
UVC_RECT, UVC_XSBL, UVC_BM = 1, 2, 3
BUF_RAW_L, BUF_RAW_R, BUF_RECT_L, BUF_RECT_R, BUF_XSBL_L, BUF_XSBL_R, BUF_DISP = range(7)

# StereoCameraModel's localTransform (slam/src/core/StereoCameraModel.cpp:9-14): z-forward camera -> x-forward body
LOCAL_TRANSFORM = np.array([0, 0, 1, 0, -1, 0, 0, 0, 0, -1, 0, 0], np.float32)

# shipped rectification parameter set (StereoBM/src/fpga.c:190-226)
SHIPPED_RECT_PARAMS = dict(
    f=[[40419817, 40382910], [39609530, 39627967]], c=[320, 240],
    f2inv=[6338213, 6338213], c2_f2=[4984405, 5932596],
    rot=[[[16598538, -120818, 2439034], [137992, 16776300, -108069], [-2438123, 126979, 16598626]],
         [[16569087, -69780, 2633522], [51223, 16776692, 122251], [-2633948, -112694, 16568783]]])


class U96Error(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__(f"{what}: {code} ({detail})")


class BmParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "width", "height", "block_size", "num_disparities", "min_disparity", "prefilter_cap",
        "uniqueness_ratio", "texture_threshold", "profile", "uni_enable", "uni_mode", "uni_thr",
        "x_store_offset", "rtl_extended", "disp12_max_diff", "speckle_window_size", "speckle_range")]


class RectParams(ctypes.Structure):
    _fields_ = [("f", (ctypes.c_int32 * 2) * 2), ("c", ctypes.c_int32 * 2), ("f2inv", ctypes.c_int32 * 2),
                ("c2_f2", ctypes.c_int32 * 2), ("rot", ((ctypes.c_int32 * 3) * 3) * 2)]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for cam in range(2):
            for k in range(2):
                p.f[cam][k] = int(d["f"][cam][k])
            for i in range(3):
                for j in range(3):
                    p.rot[cam][i][j] = int(d["rot"][cam][i][j])
        for k in range(2):
            p.c[k] = int(d["c"][k]); p.f2inv[k] = int(d["f2inv"][k]); p.c2_f2[k] = int(d["c2_f2"][k])
        return p


def lib_path():
    return os.environ.get("U96_LIB") or os.path.join(PKG, "lib", "libu96stereo.so")      # U96_LIB: developer A/B builds (tools/ab_build.sh)


_LIB = None


def load_library():
    """Loads libu96stereo.so (in-tree).  Raises ImportError when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m u96_slam_b200.build` (CUDA extension is mandatory, "
                          "there is no CPU fallback)")
    L = ctypes.CDLL(path)
    vp, i32, u8p = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
    L.u96_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32]
    L.u96_destroy.argtypes = [vp]; L.u96_destroy.restype = None
    L.u96_set_bm_params.argtypes = [vp, ctypes.POINTER(BmParams)]
    L.u96_get_bm_params.argtypes = [vp, ctypes.POINTER(BmParams)]
    L.u96_set_bm_registers.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    L.u96_set_rect_params.argtypes = [vp, ctypes.POINTER(RectParams)]
    L.u96_set_stream.argtypes = [vp, vp]
    for name in ("u96_submit_raw", "u96_submit_rect", "u96_submit_xsbl",
                 "u96_submit_raw_device", "u96_submit_rect_device", "u96_submit_xsbl_device"):
        getattr(L, name).argtypes = [vp, i32, u8p, u8p, i32, i32]
    L.u96_submit_raw_async.argtypes = [vp, i32, u8p, u8p, i32, i32, vp]
    L.u96_submit_rect_async.argtypes = [vp, i32, u8p, u8p, i32, i32, vp]
    L.u96_wait.argtypes = [vp, ctypes.POINTER(i32)]
    L.u96_receive_rect.argtypes = [vp, i32, u8p, u8p]
    L.u96_receive_xsbl.argtypes = [vp, i32, u8p, u8p]
    L.u96_receive_disp.argtypes = [vp, i32, vp]
    L.u96_receive_uvc.argtypes = [vp, i32, i32, vp]
    L.u96_set_gftt.argtypes = [vp, i32]
    L.u96_receive_eigen.argtypes = [vp, i32, vp, vp]
    L.u96_enqueue_receive_disp.argtypes = [vp, i32, vp]
    L.u96_reproject.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), i32, i32, vp]
    fp = ctypes.POINTER(ctypes.c_float)
    L.u96_reproject_ex.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), i32, fp, fp, vp]
    L.u96_reproject_points.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), vp, i32, vp,
                                       ctypes.c_float, ctypes.c_float, fp, vp]
    L.u96_set_rect_image.argtypes = [vp, i32, u8p, u8p, i32, i32]
    L.u96_start_xsbl.argtypes = [vp, i32]
    L.u96_host_alloc_wc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.u96_last_stage_ms_ex.argtypes = [vp, i32, fp, i32]
    L.u96_last_aux_ms.argtypes = [vp, i32, fp]
    L.u96_bank_device_ptr.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_size_t)]
    L.u96_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.u96_host_free.argtypes = [vp]
    L.u96_set_profiling.argtypes = [vp, i32]
    L.u96_last_stage_ms.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_float)]
    L.u96_kernel_launches.argtypes = [vp]; L.u96_kernel_launches.restype = ctypes.c_int64
    L.u96_microbench.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.u96_strerror.argtypes = [i32]; L.u96_strerror.restype = ctypes.c_char_p
    L.u96_last_cuda_error.restype = ctypes.c_char_p
    L.u96_abi_version.restype = i32
    _LIB = L
    return L


def _check(L, rc, what):
    if rc != 0:
        detail = L.u96_strerror(rc).decode()
        if rc == -2:
            detail += ": " + L.u96_last_cuda_error().decode()
        raise U96Error(rc, what, detail)


def _as_batch(a, dtype=np.uint8):
    a = np.asarray(a)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("expected [H,W] or [n,H,W]")
    return np.ascontiguousarray(a, dtype=dtype)


class StereoFrontEnd:
    """One handle per GPU (u96_create .. u96_destroy)."""

    def __init__(self, device=0, max_w=640, max_h=480, max_batch=1):
        self.L = load_library()
        self.h = ctypes.c_void_p()
        _check(self.L, self.L.u96_create(ctypes.byref(self.h), device, max_w, max_h, max_batch), "u96_create")
        self.device, self.max_batch = device, max_batch
        self._n = [0, 0]
        self._geom = [None, None]                       # (W, H) each bank was filled with
        p = self.get_bm_params()                        # the handle starts with the firmware's defaults (fpga.c:150-160)
        self.W, self.H = p["width"], p["height"]

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.u96_destroy(self.h)
            self.h = ctypes.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration ----
    def get_bm_params(self):
        p = BmParams()
        _check(self.L, self.L.u96_get_bm_params(self.h, ctypes.byref(p)), "u96_get_bm_params")
        return {n: getattr(p, n) for n, _ in BmParams._fields_}

    def set_bm_params(self, **kw):
        cur = self.get_bm_params()
        cur.update(kw)
        p = BmParams(**cur)
        _check(self.L, self.L.u96_set_bm_params(self.h, ctypes.byref(p)), "u96_set_bm_params")
        self.W, self.H = p.width, p.height

    def set_bm_registers(self, image_size, bm_setting, uni_filt_ctrl=0):
        _check(self.L, self.L.u96_set_bm_registers(self.h, image_size, bm_setting, uni_filt_ctrl), "u96_set_bm_registers")
        p = self.get_bm_params()
        self.W, self.H = p["width"], p["height"]

    def set_rect_params(self, d):
        p = d if isinstance(d, RectParams) else RectParams.from_dict(d)
        _check(self.L, self.L.u96_set_rect_params(self.h, ctypes.byref(p)), "u96_set_rect_params")

    def set_stream(self, cuda_stream_ptr):
        _check(self.L, self.L.u96_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)), "u96_set_stream")

    def set_profiling(self, on=True):
        _check(self.L, self.L.u96_set_profiling(self.h, int(on)), "u96_set_profiling")

    # ---- data plane ----
    def _submit(self, fn, bank, L, R):
        L, R = _as_batch(L), _as_batch(R)
        if L.shape != R.shape:
            raise ValueError("L/R shape mismatch")
        n, H, W = L.shape
        if (W, H) != (self.W, self.H):
            raise ValueError(f"image {W}x{H} != configured {self.W}x{self.H}")
        _check(self.L, fn(self.h, bank, L.ctypes.data, R.ctypes.data, W, n), fn.__name__)
        self._n[bank] = n; self._geom[bank] = (self.W, self.H)

    def set_rect_image(self, bank, L, R):
        """Fpga::setRectImage (FPGA.cpp:236-249): write the pair(s) into the RECT bank, run nothing"""
        self._submit(self.L.u96_set_rect_image, bank, L, R)

    def start_xsbl(self, bank):
        """reg->xsbl.Control |= FPGA_XSBL_SW_START (main.cpp:172-174)"""
        _check(self.L, self.L.u96_start_xsbl(self.h, bank), "u96_start_xsbl")

    def submit_raw(self, bank, L, R):
        self._submit(self.L.u96_submit_raw, bank, L, R)

    def submit_rect(self, bank, L, R):
        self._submit(self.L.u96_submit_rect, bank, L, R)

    def submit_xsbl(self, bank, L, R):
        self._submit(self.L.u96_submit_xsbl, bank, L, R)

    def submit_device(self, kind, bank, dptr_l, dptr_r, stride, n):
        fn = {"raw": self.L.u96_submit_raw_device, "rect": self.L.u96_submit_rect_device,
              "xsbl": self.L.u96_submit_xsbl_device}[kind]
        _check(self.L, fn(self.h, bank, ctypes.c_void_p(dptr_l), ctypes.c_void_p(dptr_r), stride, n), fn.__name__)
        self._n[bank] = n; self._geom[bank] = (self.W, self.H)

    def submit_host_ptr(self, kind, bank, ptr_l, ptr_r, stride, n):
        fn = {"raw": self.L.u96_submit_raw, "rect": self.L.u96_submit_rect, "xsbl": self.L.u96_submit_xsbl}[kind]
        _check(self.L, fn(self.h, bank, ctypes.c_void_p(ptr_l), ctypes.c_void_p(ptr_r), stride, n), fn.__name__)
        self._n[bank] = n; self._geom[bank] = (self.W, self.H)

    def submit_host_ptr_async(self, kind, bank, ptr_l, ptr_r, stride, n, disp_out_ptr):
        """pipelined submit: H2D, kernels and the D2H of the disparity overlap chunk by chunk"""
        fn = {"raw": self.L.u96_submit_raw_async, "rect": self.L.u96_submit_rect_async}[kind]
        _check(self.L, fn(self.h, bank, ctypes.c_void_p(ptr_l), ctypes.c_void_p(ptr_r), stride, n, ctypes.c_void_p(disp_out_ptr)), fn.__name__)
        self._n[bank] = n; self._geom[bank] = (self.W, self.H)

    def wait(self):
        b = ctypes.c_int(-1)
        _check(self.L, self.L.u96_wait(self.h, ctypes.byref(b)), "u96_wait")
        return b.value

    def _recv_pair(self, fn, bank):
        n = self._n[bank]; W, H = self._geom[bank]
        L = np.empty((n, H, W), np.uint8); R = np.empty_like(L)
        _check(self.L, fn(self.h, bank, L.ctypes.data, R.ctypes.data), fn.__name__)
        return L, R

    def receive_rect(self, bank):
        return self._recv_pair(self.L.u96_receive_rect, bank)

    def receive_xsbl(self, bank):
        return self._recv_pair(self.L.u96_receive_xsbl, bank)

    def receive_disp(self, bank, out=None):
        n = self._n[bank]; W, H = self._geom[bank]
        d = out if out is not None else np.empty((n, H, W), np.int16)
        _check(self.L, self.L.u96_receive_disp(self.h, bank, d.ctypes.data), "u96_receive_disp")
        return d

    def set_gftt(self, on=True):
        """fpga->gftt.Control = FPGA_GFTT_CTRL_ENABLE (fpga.c:162-172)"""
        _check(self.L, self.L.u96_set_gftt(self.h, 1 if on else 0), "u96_set_gftt")

    def receive_eigen(self, bank):
        """Fpga::receiveEigen (FPGA.cpp:281-296): (n, H, W) u16 min-eigenvalue map, (n,) u16 per-frame maxima (gftt.Max)"""
        n = self._n[bank]; W, H = self._geom[bank]
        e = np.empty((n, H, W), np.uint16); m = np.empty(n, np.uint16)
        _check(self.L, self.L.u96_receive_eigen(self.h, bank, e.ctypes.data, m.ctypes.data), "u96_receive_eigen")
        return e, m

    def receive_uvc(self, bank, which):
        """UVC payload of the firmware (xusb_main.c:293-376): (n, H, 2W, 2) u8 YUYV; which = UVC_RECT / UVC_XSBL / UVC_BM"""
        n = self._n[bank]; W, H = self._geom[bank]
        f = np.empty((n, H, 2 * W, 2), np.uint8)
        _check(self.L, self.L.u96_receive_uvc(self.h, bank, which, f.ctypes.data), "u96_receive_uvc")
        return f

    def receive_disp_ptr(self, bank, host_ptr):
        _check(self.L, self.L.u96_receive_disp(self.h, bank, ctypes.c_void_p(host_ptr)), "u96_receive_disp")

    def enqueue_receive_disp_ptr(self, bank, host_ptr):
        """async D2H of the disparity behind the bank's kernels; the next wait() of this bank covers it"""
        _check(self.L, self.L.u96_enqueue_receive_disp(self.h, bank, ctypes.c_void_p(host_ptr)), "u96_enqueue_receive_disp")

    def reproject(self, bank, P_l, P_r, decim=1, apply_local=False):
        n = self._n[bank]
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        W, H = self._geom[bank]
        out = np.empty((n, H // decim, W // decim, 3), np.float32)
        dp = ctypes.POINTER(ctypes.c_double)
        _check(self.L, self.L.u96_reproject(self.h, bank, Pl.ctypes.data_as(dp), Pr.ctypes.data_as(dp), decim,
                                            1 if apply_local else 0, out.ctypes.data), "u96_reproject")
        return out

    def reproject_ex(self, bank, P_l, P_r, decim=1, local_T=None, poses=None):
        """dense consumer of main.cpp:522-551: local_T 12 floats or None, poses (n, 12) floats or None"""
        n = self._n[bank]
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        W, H = self._geom[bank]
        out = np.empty((n, H // decim, W // decim, 3), np.float32)
        dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
        lt = None if local_T is None else np.ascontiguousarray(local_T, np.float32).reshape(12)
        ps = None if poses is None else np.ascontiguousarray(poses, np.float32).reshape(n, 12)
        _check(self.L, self.L.u96_reproject_ex(self.h, bank, Pl.ctypes.data_as(dp), Pr.ctypes.data_as(dp), decim,
                                               None if lt is None else lt.ctypes.data_as(fp), None if ps is None else ps.ctypes.data_as(fp),
                                               out.ctypes.data), "u96_reproject_ex")
        return out

    def reproject_points(self, bank, P_l, P_r, uv, frame=0, min_depth=0.0, max_depth=0.0, local_T=LOCAL_TRANSFORM, mask=None):
        """generateKeypoints3DStereo (Stereo.cpp:53-117): uv (n, 2) float32 keypoints (x, y) -> (n, 3) float32, NaN = bad point"""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        n = uv.shape[0]
        Pl = np.ascontiguousarray(P_l, np.float64).reshape(12); Pr = np.ascontiguousarray(P_r, np.float64).reshape(12)
        out = np.empty((n, 3), np.float32)
        dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
        lt = None if local_T is None else np.ascontiguousarray(local_T, np.float32).reshape(12)
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        _check(self.L, self.L.u96_reproject_points(self.h, bank, frame, Pl.ctypes.data_as(dp), Pr.ctypes.data_as(dp), uv.ctypes.data, n,
                                                   None if mk is None else mk.ctypes.data, min_depth, max_depth,
                                                   None if lt is None else lt.ctypes.data_as(fp), out.ctypes.data), "u96_reproject_points")
        return out

    def bank_device_ptr(self, bank, which):
        p = ctypes.c_void_p(); pitch = ctypes.c_int(); frame = ctypes.c_size_t()
        _check(self.L, self.L.u96_bank_device_ptr(self.h, bank, which, ctypes.byref(p), ctypes.byref(pitch), ctypes.byref(frame)),
               "u96_bank_device_ptr")
        return p.value, pitch.value, frame.value

    def disp_tensor(self, bank):
        """Zero-copy torch view (n, H, W) int16 of the bank's disparity maps in HBM (u96_bank_device_ptr), e.g. as the
        send buffer of the optional multi-GPU gather; valid until the bank is submitted again."""
        import torch
        ptr, pitch, frame = self.bank_device_ptr(bank, BUF_DISP)
        n = self._n[bank]; W, H = self._geom[bank]

        class _Cai:                                    # CUDA array interface v2
            __cuda_array_interface__ = {"shape": (n, H, W), "typestr": "<i2", "data": (ptr, False), "version": 2,
                                        "strides": (frame, pitch, 2)}
        return torch.as_tensor(_Cai(), device=f"cuda:{self.device}")

    def last_stage_ms(self, bank):
        ms = (ctypes.c_float * 4)()
        _check(self.L, self.L.u96_last_stage_ms(self.h, bank, ms), "u96_last_stage_ms")
        return dict(zip(("h2d", "rect", "xsbl", "bm"), list(ms)))

    def last_stage_ms_ex(self, bank):
        ms = (ctypes.c_float * 6)()
        _check(self.L, self.L.u96_last_stage_ms_ex(self.h, bank, ms, 6), "u96_last_stage_ms_ex")
        return dict(zip(("h2d", "rect", "gftt", "xsbl", "bm", "post"), list(ms)))

    def last_aux_ms(self, which):
        """kernel ms of the last reproject (0) / reproject_points (1) / receive_uvc (2) call; needs set_profiling"""
        v = ctypes.c_float()
        _check(self.L, self.L.u96_last_aux_ms(self.h, which, ctypes.byref(v)), "u96_last_aux_ms")
        return v.value

    def kernel_launches(self):
        return int(self.L.u96_kernel_launches(self.h))


def microbench(which, device=0):
    L = load_library()
    g = ctypes.c_double()
    _check(L, L.u96_microbench(device, which, ctypes.byref(g)), "u96_microbench")
    return g.value


class Fpga:
    """Drop-in for the data plane of the reference's `class Fpga` (slam/src/core/FPGA.cpp).

    registerOpen/memoryOpen create the GPU handle; setRectImage + start() replace
    `setRectImage` + `reg->xsbl.Control |= FPGA_XSBL_SW_START` (main.cpp:165-175);
    receiveData waits for DATA_READY and copies the active bank out (FPGA.cpp:310-347).
    Firmware defaults: 640x480, block 21, 64 disparities, uniqueness off (fpga.c:150-160).
    """
    IMAGE_WIDTH, IMAGE_HEIGHT = 640, 480

    def __init__(self, device=0, width=640, height=480):
        self.device, self.IMAGE_WIDTH, self.IMAGE_HEIGHT = device, width, height
        self.fe = None

    def registerOpen(self):
        try:
            self.fe = StereoFrontEnd(self.device, self.IMAGE_WIDTH, self.IMAGE_HEIGHT, 1)
        except (U96Error, ImportError):
            return -1
        # Fpga_Init BM block (fpga.c:150-160)
        self.fe.set_bm_registers((self.IMAGE_HEIGHT << 16) + self.IMAGE_WIDTH, 0x00150040, 0)
        self.fe.set_rect_params(SHIPPED_RECT_PARAMS)
        return 0

    def registerClose(self):
        if self.fe:
            self.fe.close(); self.fe = None
        return 0

    memoryOpen = lambda self: 0 if self.fe else -1     # banks live in HBM inside the handle
    memoryClose = lambda self: 0

    def setRectImage(self, bank, imageLeft, imageRight):
        """FPGA.cpp:236-249: the pair is written into the RECT bank at call time"""
        self.fe.set_rect_image(bank, imageLeft, imageRight)

    def start(self, bank):
        """FPGA_XSBL_SW_START: run xsbl -> bm on the RECT bank."""
        self.fe.start_xsbl(bank)

    def generateKeypoints3D(self, bank, P_l, P_r, keypoints, minDepth=0.0, maxDepth=0.0):
        """Stereo.cpp:119-154 with DEPTH_METHOD_FPGA_BM: keypoints (n, 2) float32 (x, y) -> (n, 3) float32 in the body frame"""
        return self.fe.reproject_points(bank, P_l, P_r, keypoints, 0, minDepth, maxDepth)

    def captureFromSensor(self, bank, rawLeft, rawRight):
        """sensor path: rect -> xsbl -> bm (CameraStereoImages.cpp:134-149)"""
        self.fe.submit_raw(bank, rawLeft, rawRight)

    def receiveRectImages(self, bank):
        L, R = self.fe.receive_rect(bank)
        return L[0], R[0]

    def receiveDepthMap(self, bank):
        return self.fe.receive_disp(bank)[0]

    def enableGftt(self, on=True):
        """fpga->gftt.Control = FPGA_GFTT_CTRL_ENABLE when RETURN_DATA_GFTT is requested (fpga.c:162-172)"""
        self.fe.set_gftt(on)

    def receiveEigen(self, bank):
        """FPGA.cpp:281-296 -> (CV_16UC1 map, maxEigen)"""
        e, m = self.fe.receive_eigen(bank)
        return e[0], int(m[0])

    def receiveData(self):
        """-> (activeBank, rectL, rectR, disparity CV_16SC1)"""
        bank = self.fe.wait()
        L, R = self.receiveRectImages(bank)
        return bank, L, R, self.receiveDepthMap(bank)


class StereoBM:
    """cv::StereoBM-shaped front (main.cpp:198-215) running PROFILE_OPENCV on the GPU."""

    def __init__(self, numDisparities=64, blockSize=21, device=0):
        self.p = dict(num_disparities=numDisparities, block_size=blockSize, prefilter_cap=31, texture_threshold=10,
                      uniqueness_ratio=15, min_disparity=0, profile=PROFILE_OPENCV, disp12_max_diff=-1,
                      speckle_window_size=0, speckle_range=0)
        self.device, self.fe, self._shape = device, None, None

    @classmethod
    def create(cls, numDisparities=64, blockSize=21, device=0):
        return cls(numDisparities, blockSize, device)

    def setPreFilterCap(self, v): self.p["prefilter_cap"] = v
    def setBlockSize(self, v): self.p["block_size"] = v
    def setMinDisparity(self, v): self.p["min_disparity"] = v
    def setNumDisparities(self, v): self.p["num_disparities"] = v
    def setTextureThreshold(self, v): self.p["texture_threshold"] = v
    def setUniquenessRatio(self, v): self.p["uniqueness_ratio"] = v
    def setSpeckleWindowSize(self, v): self.p["speckle_window_size"] = v
    def setSpeckleRange(self, v): self.p["speckle_range"] = v
    def setDisp12MaxDiff(self, v): self.p["disp12_max_diff"] = v

    def compute(self, left, right):
        left = np.asarray(left)
        H, W = left.shape[-2:]
        if self.fe is None or self._shape != (W, H):
            if self.fe:
                self.fe.close()
            self.fe = StereoFrontEnd(self.device, W, H, 1)
            self._shape = (W, H)
        self.fe.set_bm_params(width=W, height=H, **self.p)
        self.fe.submit_rect(0, left, right)
        bank = self.fe.wait()
        return self.fe.receive_disp(bank)[0]
