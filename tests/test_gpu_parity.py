"""GPU parity tests (run with -m gpu on the B200 box): every stage through the C ABI, bit-exact against
the CPU oracle, the reference's golden vectors and the committed cv2 fixtures."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def u(libpath):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import u96_slam_b200
    return u96_slam_b200


@pytest.fixture(scope="module")
def fe640(u):
    fe = u.StereoFrontEnd(0, 640, 480, 4)
    yield fe
    fe.close()


def run_xsbl(fe, bank, xl, xr, **params):
    fe.set_bm_params(**params)
    fe.submit_xsbl(bank, xl, xr)
    assert fe.wait() == bank
    return fe.receive_disp(bank)


RTL = dict(width=640, height=480, profile=0, num_disparities=64, block_size=21, uni_enable=0, uni_mode=0, uni_thr=0,
           x_store_offset=1, rtl_extended=0, min_disparity=0)


def test_c1_xsobel_matches_reference_golden(u, fe640, golden):
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)       # Fpga_Init values, fpga.c:150-160
    fe640.submit_rect(0, golden["rect_l"], golden["rect_r"])
    b = fe640.wait()
    sl, sr = fe640.receive_xsbl(b)
    assert np.array_equal(sl[0], golden["xsbl_l"]) and np.array_equal(sr[0], golden["xsbl_r"])
    rl, rr = fe640.receive_rect(b)                                   # copy-out of the bank (receiveRectImages)
    assert np.array_equal(rl[0], golden["rect_l"]) and np.array_equal(rr[0], golden["rect_r"])


@pytest.mark.parametrize("wsz,uni,thr,mode,crc", [
    (21, 0, 0, 0, 0x3C312D26), (21, 1, 921, 0, 0xF3284A7C), (15, 0, 0, 0, 0xD0650EA3),
    (9, 1, 800, 1, None), (5, 0, 0, 0, None), (17, 0, 0, 0, None), (31, 0, 0, 0, None), (3, 0, 0, 0, None),
    (19, 0, 0, 0, None), (27, 1, 600, 0, None), (7, 0, 0, 0, None)])
def test_c1_bm_rtl_bundled_pair(u, fe640, golden, oracle, wsz, uni, thr, mode, crc):
    d = run_xsbl(fe640, 1, golden["xsbl_l"], golden["xsbl_r"], **dict(RTL, block_size=wsz, uni_enable=uni, uni_thr=thr, uni_mode=mode))[0]
    want = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=wsz, ndisp=64, uni_enb=uni, uni_thr=thr, uni_mode=mode)
    assert np.array_equal(d, want), int((d != want).sum())
    if crc is not None:                                              # SURVEY Appendix B cross-check values
        assert zlib.crc32(d.tobytes()) & 0xFFFFFFFF == crc


@pytest.mark.parametrize("D", [32, 96, 128, 256])
def test_bm_rtl_disparity_ranges(u, fe640, golden, oracle, D):
    ext = int(D > 128)
    d = run_xsbl(fe640, 0, golden["xsbl_l"], golden["xsbl_r"], **dict(RTL, num_disparities=D, rtl_extended=ext))[0]
    assert np.array_equal(d, oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=D, rtl_extended=ext))


def test_bm_rtl_x_store_offset_and_sign_extension(u, fe640, golden, oracle):
    xl, xr = golden["xsbl_l"], golden["xsbl_r"]
    d0 = run_xsbl(fe640, 0, xl, xr, **dict(RTL, x_store_offset=0))[0]
    assert np.array_equal(d0, oracle.bm_rtl(xl, xr, wsz=21, ndisp=64, x_store_offset=0))
    # D=256 without the extension reproduces the RTL's bit-15 sign extension (bm_obuf2.v:153)
    d = run_xsbl(fe640, 1, xl, xr, **dict(RTL, num_disparities=256, rtl_extended=0))[0]
    assert np.array_equal(d, oracle.bm_rtl(xl, xr, wsz=21, ndisp=256, rtl_extended=0))


def test_opencv_profile_matches_cv2_golden(u, fe640, golden, cv_golden, oracle):
    for k, want in cv_golden.items():
        if not k.startswith("D"):
            continue
        D, B, T, U = [int(s[1:]) for s in k.split("_")]
        fe640.set_bm_params(width=640, height=480, profile=u.PROFILE_OPENCV, num_disparities=D, block_size=B,
                            texture_threshold=T, uniqueness_ratio=U, prefilter_cap=31, min_disparity=0)
        fe640.submit_rect(0, golden["rect_l"], golden["rect_r"])
        b = fe640.wait()
        pl, pr = fe640.receive_xsbl(b)
        assert np.array_equal(pl[0], oracle.xsobel_cv(golden["rect_l"], 31))
        assert np.array_equal(fe640.receive_disp(b)[0], want), k


def test_stereobm_facade_matches_live_cv2(u):
    cv2 = pytest.importorskip("cv2")
    L, R = u.synth_pair(2, 3, 400, 200, 48)
    ref = cv2.StereoBM_create(48, 11)
    ref.setPreFilterCap(20); ref.setTextureThreshold(5); ref.setUniquenessRatio(12)
    ref.setSpeckleWindowSize(0); ref.setDisp12MaxDiff(-1)
    bm = u.StereoBM.create(48, 11)
    bm.setPreFilterCap(20); bm.setTextureThreshold(5); bm.setUniquenessRatio(12)
    assert np.array_equal(bm.compute(L, R), ref.compute(L, R))


def test_c2_raw_pipeline_batch_and_banks(u, fe640, oracle):
    L, R = u.synth_batch(1, 0, 4, 640, 480, 64)
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_bm_params(x_store_offset=1)
    fe640.set_rect_params(u.SHIPPED_RECT_PARAMS)
    fe640.submit_raw(0, L[:3], R[:3])                # two banks in flight, FIFO order
    fe640.submit_raw(1, L[3:], R[3:])
    assert fe640.wait() == 0 and fe640.wait() == 1
    out = {0: (fe640.receive_rect(0), fe640.receive_xsbl(0), fe640.receive_disp(0)),
           1: (fe640.receive_rect(1), fe640.receive_xsbl(1), fe640.receive_disp(1))}
    for i in range(4):
        bank, j = (0, i) if i < 3 else (1, 0)
        (gl, gr), (sl, sr), d = out[bank]
        wl, wr = oracle.rectify(L[i], u.SHIPPED_RECT_PARAMS, 0), oracle.rectify(R[i], u.SHIPPED_RECT_PARAMS, 1)
        assert np.array_equal(gl[j], wl) and np.array_equal(gr[j], wr)
        wxl, wxr = oracle.xsobel_rtl(wl), oracle.xsobel_rtl(wr)
        assert np.array_equal(sl[j], wxl) and np.array_equal(sr[j], wxr)
        assert np.array_equal(d[j], oracle.bm_rtl(wxl, wxr, wsz=21, ndisp=64))


def test_fpga_facade_file_mode(u, golden, oracle):
    """FPGA_TEST call stack (main.cpp:165-181): setRectImage -> SW_START -> receiveData."""
    f = u.Fpga()
    assert f.registerOpen() == 0 and f.memoryOpen() == 0
    for it in range(3):
        bank = it % 2
        f.setRectImage(bank, golden["rect_l"], golden["rect_r"])
        f.start(bank)
        active, rl, rr, disp = f.receiveData()
        assert active == bank and disp.dtype == np.int16 and disp.shape == (480, 640)
        assert np.array_equal(disp, oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64))
    assert f.registerClose() == 0


@pytest.mark.parametrize("B", [9, 15, 21])
def test_c3_kitti_shape_both_profiles(u, oracle, B):
    W, H, D = 1242, 375, 128
    L, R = u.synth_batch(2, 0, 2, W, H, D)
    rp = u.identity_rect_params(W, H, 700.0)
    with u.StereoFrontEnd(0, W, H, 2) as fe:
        fe.set_rect_params(rp)
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1)
        fe.submit_raw(0, L, R); b = fe.wait()
        gl, gr = fe.receive_rect(b); d = fe.receive_disp(b)
        for i in range(2):                                           # every frame of the batch
            wl, wr = oracle.rectify(L[i], rp, 0), oracle.rectify(R[i], rp, 1)
            assert np.array_equal(gl[i], wl) and np.array_equal(gr[i], wr)
            assert np.array_equal(d[i], oracle.bm_rtl(oracle.xsobel_rtl(wl), oracle.xsobel_rtl(wr), wsz=B, ndisp=D, bitserial_div=0))
        fe.set_bm_params(profile=u.PROFILE_OPENCV, texture_threshold=10, uniqueness_ratio=10, prefilter_cap=31)
        fe.submit_rect(1, L, R); b = fe.wait()
        d = fe.receive_disp(b)
        for i in range(2):
            assert np.array_equal(d[i], oracle.bm_cv(oracle.xsobel_cv(L[i]), oracle.xsobel_cv(R[i]), wsz=B, ndisp=D))


def test_c4_full_hd_256_disparities(u, oracle):
    W, H, D = 1920, 1080, 256
    L, R = u.synth_batch(3, 0, 2, W, H, D)
    with u.StereoFrontEnd(0, W, H, 2) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=21, num_disparities=D, rtl_extended=1, x_store_offset=1)
        fe.submit_rect(0, L, R); b = fe.wait()
        d = fe.receive_disp(b)
        for i in range(2):                                           # every frame of the batch
            want = oracle.bm_rtl(oracle.xsobel_rtl(L[i]), oracle.xsobel_rtl(R[i]), wsz=21, ndisp=D, rtl_extended=1, bitserial_div=0)
            assert np.array_equal(d[i], want)
        # size-independent properties at full size: borders invalid, range, idempotence, batch == single
        assert (d[:, :10] == -1).all() and (d[:, :, :D + 11] == -1).all() and d.max() < D * 16 + 8
        fe.submit_rect(1, L[1:], R[1:]); b = fe.wait()
        assert np.array_equal(fe.receive_disp(b)[0], d[1])


def rotated_rect_params(u, W, H, deg=1.2):
    """A rectifying rotation of about `deg` degrees about every axis per camera (opposite signs left / right) through the
    calibration -> 27 registers generator (formats.rect_params_from_calibration)."""
    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return Rz @ Ry @ Rx
    a = np.deg2rad(deg)
    f = 0.9 * W
    K_src = [(f * 1.01, f * 1.012, W / 2 + 3, H / 2 - 2), (f * 0.995, f * 0.993, W / 2 - 4, H / 2 + 1)]
    R_rect = [rot(0.4 * a, a, -0.7 * a), rot(-0.3 * a, -0.8 * a, 0.6 * a)]
    return u.rect_params_from_calibration(K_src, R_rect, (f * 0.97, f * 0.97, W / 2, H / 2))


@pytest.mark.parametrize("kind", ["identity", "rotated"])
def test_c4_full_hd_raw_path(u, oracle, kind):
    """the raw entry point at 1920x1080 (k_rect_remap_tma + staged transfers + cluster BM), every frame checked:
    RECT and DISP banks against the oracle for a near-identity and a rotated rectification set."""
    W, H, D = 1920, 1080, 256
    L, R = u.synth_batch(3, 4, 2, W, H, D)
    rp = u.identity_rect_params(W, H, float(W)) if kind == "identity" else rotated_rect_params(u, W, H)
    with u.StereoFrontEnd(0, W, H, 2) as fe:
        fe.set_rect_params(rp)
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=21, num_disparities=D, rtl_extended=1, x_store_offset=1)
        fe.submit_raw(0, L, R); b = fe.wait()
        gl, gr = fe.receive_rect(b); xl, xr = fe.receive_xsbl(b); d = fe.receive_disp(b)
        for i in range(2):
            wl, wr = oracle.rectify(L[i], rp, 0), oracle.rectify(R[i], rp, 1)
            assert np.array_equal(gl[i], wl) and np.array_equal(gr[i], wr), (kind, i)
            sl, sr = oracle.xsobel_rtl(wl), oracle.xsobel_rtl(wr)
            assert np.array_equal(xl[i], sl) and np.array_equal(xr[i], sr)
            assert np.array_equal(d[i], oracle.bm_rtl(sl, sr, wsz=21, ndisp=D, rtl_extended=1, bitserial_div=0)), (kind, i)
        if kind == "rotated":
            assert (gl[0] != L[0]).mean() > 0.5                      # the map really moves pixels
            assert (d[0] > 0).mean() > 0.3


def test_c4_full_hd_opencv_profile_256_disparities(u, oracle):
    """PROFILE_OPENCV above 1242 px: 1920x1080, D256 (4-CTA clusters), texture + uniqueness, every frame checked;
    then the main.cpp:210-212 post filters on the same frames."""
    W, H, D = 1920, 1080, 256
    L, R = u.synth_batch(3, 8, 2, W, H, D)
    with u.StereoFrontEnd(0, W, H, 2) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_OPENCV, block_size=21, num_disparities=D, prefilter_cap=31,
                         texture_threshold=10, uniqueness_ratio=10, disp12_max_diff=-1, speckle_window_size=0, speckle_range=0)
        fe.submit_rect(0, L, R); b = fe.wait()
        d = fe.receive_disp(b)
        for i in range(2):
            assert np.array_equal(d[i], oracle.bm_cv(oracle.xsobel_cv(L[i]), oracle.xsobel_cv(R[i]), wsz=21, ndisp=D)), i
        assert (d[0] >= 0).mean() > 0.3
        fe.set_bm_params(disp12_max_diff=1, speckle_window_size=50, speckle_range=32)
        fe.submit_rect(1, L, R); b = fe.wait()
        d2 = fe.receive_disp(b)
        want = oracle.bm_cv_post(oracle.xsobel_cv(L[1]), oracle.xsobel_cv(R[1]), wsz=21, ndisp=D)
        assert np.array_equal(d2[1], want)
        assert (d2[1] != d[1]).sum() > 100                           # the filters did something


def test_property_uniform_shift_of_right_image(u, fe640, oracle):
    """Shifting R right by k columns lowers every interior disparity by exactly k (x16)."""
    L, R = u.synth_pair(1, 7, 640, 480, 64)
    xl, xr = oracle.xsobel_rtl(L), oracle.xsobel_rtl(R)
    k = 3
    xr2 = np.zeros_like(xr); xr2[:, k:] = xr[:, :-k]
    d1 = run_xsbl(fe640, 0, xl, xr, **dict(RTL, block_size=15))[0]
    d2 = run_xsbl(fe640, 1, xl, xr2, **dict(RTL, block_size=15))[0]
    both = (d1 >= (k + 2) * 16) & (d2 >= 32) & (d1 < 60 * 16)
    assert both.sum() > 50000
    assert (d1[both] - d2[both] == 16 * k).mean() > 0.98


def test_reproject_matches_oracle(u, fe640, oracle):
    L, R = u.synth_pair(1, 2, 640, 480, 64)
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_bm_params(x_store_offset=1)
    fe640.submit_rect(0, L, R); b = fe640.wait()
    d = fe640.receive_disp(b)[0]
    sx, sy = 640 / 1241, 480 / 376                                   # StereoCameraModel.cpp:108-119 scaling
    P_l = np.array([[718.856 * sx, 0, 607.1928 * sx, 0], [0, 718.856 * sy, 185.2157 * sy, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * sx
    for decim, loc in ((1, False), (4, True), (2, True)):
        got = fe640.reproject(b, P_l, P_r, decim, loc)[0]
        want = oracle.reproject(d, P_l, P_r, decim, int(loc))
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or \
            ((got == want) | (np.isnan(got) & np.isnan(want))).all()


def _bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_reproject_points_matches_oracle_and_reference_fixture(u, fe640, oracle):
    """a16, keypoint form (generateKeypoints3DStereo, Stereo.cpp:53-117): GPU == oracle on 20 000 random keypoints per frame
    of a real disparity batch (borders, invalid -1 pixels, out-of-map points, mask, depth gates, no / default / odd
    localTransform), and GPU == the committed outputs of the reference's compiled code on the fixture's map."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import reproject_inputs
    L, R = u.synth_batch(1, 4, 3, 640, 480, 64)
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_bm_params(x_store_offset=1)
    fe640.submit_rect(1, L, R); b = fe640.wait()
    d = fe640.receive_disp(b)
    sx, sy = 640 / 1241, 480 / 376
    P_l = np.array([[718.856 * sx, 0, 607.1928 * sx, 0], [0, 718.856 * sy, 185.2157 * sy, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * sx
    rng = np.random.default_rng(7)
    n = 20000
    uv = np.stack([rng.random(n) * 640, rng.random(n) * 480], 1).astype(np.float32)
    uv[:8] = [[0, 0], [639.99, 479.99], [-0.5, -0.5], [-1.0, 5], [640.0, 5], [5, 480.0], [np.nan, 3], [1e30, 1e30]]
    uv[8:2000, 0] = rng.choice([0.0, 1.5, 74.2, 85.0, 86.0, 627.9, 628.0, 639.0], 1992)      # valid-rectangle borders (D + h, W - 2 - h + 1)
    mask = (rng.random(n) < 0.9).astype(np.uint8)
    odd_T = np.array([0.5, -0.25, 1.0, 0.125, -1.0, 0.0, 0.75, -2.0, 0.0, -1.0, 0.3, 4.0], np.float32)
    total_good = 0
    for f in range(3):
        for (mn, mx, T, mk) in ((0.0, 0.0, u.LOCAL_TRANSFORM, None), (-1.0, 0.0, None, mask), (2.0, 12.0, odd_T, mask),
                                (0.0, 4.0, np.zeros(12, np.float32), None)):
            got = fe640.reproject_points(b, P_l, P_r, uv, f, mn, mx, T, mk)
            want = oracle.reproject_points(d[f], P_l, P_r, uv, mn, mx, T, mk)
            assert _bits_equal(got, want), (f, mn, mx)
            total_good += int(np.isfinite(got[:, 0]).sum())
        assert np.isnan(got[:8][[3, 4, 5, 6, 7]]).all()
    assert total_good > 50000
    # the reference's own outputs (fixture): upload its disparity map as a bank and gather
    (W, H, disp, kuv, kP_l, kP_r, pose, gates) = reproject_inputs()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reproject_ref.npz"))
    import torch
    with u.StereoFrontEnd(0, W, H, 1) as fe:
        fe.set_bm_params(width=W, height=H, profile=0, block_size=5, num_disparities=32)
        z = np.zeros((1, H, W), np.uint8)
        fe.submit_xsbl(0, z, z); fe.wait()
        # overwrite the bank's disparity with the fixture's map (device pointer from the C ABI, wrapped zero-copy)
        view = fe.disp_tensor(0)
        view[0].copy_(torch.from_numpy(disp).cuda())
        torch.cuda.synchronize()
        for rs in (0, 1):
            Pl, Pr = g[f"P_l_resize{rs}"], g[f"P_r_resize{rs}"]
            for gi, (mn, mx) in enumerate(gates):
                assert _bits_equal(fe.reproject_points(0, Pl, Pr, kuv, 0, mn, mx), g[f"kp_resize{rs}_gate{gi}"])
            assert _bits_equal(fe.reproject_ex(0, Pl, Pr, 4, u.LOCAL_TRANSFORM, None)[0], g[f"dense_resize{rs}_local"])
            assert _bits_equal(fe.reproject_ex(0, Pl, Pr, 4, u.LOCAL_TRANSFORM, pose[None])[0], g[f"dense_resize{rs}_local_pose"])
        assert _bits_equal(fe.reproject_ex(0, g["P_l_resize0"], g["P_r_resize0"], 1, None, None)[0], g["dense_resize0_plain"])


def test_reproject_ex_local_transform_and_per_frame_pose(u, fe640, oracle):
    """dense consumer with both transforms (main.cpp:535-541), one pose per frame of the batch"""
    L, R = u.synth_batch(1, 9, 4, 640, 480, 64)
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_bm_params(x_store_offset=1)
    fe640.submit_rect(0, L, R); b = fe640.wait()
    d = fe640.receive_disp(b)
    sx, sy = 640 / 1241, 480 / 376
    P_l = np.array([[718.856 * sx, 0, 607.1928 * sx, 0], [0, 718.856 * sy, 185.2157 * sy, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * sx
    rng = np.random.default_rng(3)
    poses = np.concatenate([rng.normal(size=(4, 3, 3)), rng.normal(size=(4, 3, 1)) * 5], 2).astype(np.float32).reshape(4, 12)
    for decim in (1, 4, 8):
        got = fe640.reproject_ex(b, P_l, P_r, decim, u.LOCAL_TRANSFORM, poses)
        for f in range(4):
            assert _bits_equal(got[f], oracle.reproject_ex(d[f], P_l, P_r, decim, u.LOCAL_TRANSFORM, poses[f])), (decim, f)
    got = fe640.reproject_ex(b, P_l, P_r, 2, None, poses)
    assert _bits_equal(got[3], oracle.reproject_ex(d[3], P_l, P_r, 2, None, poses[3]))
    assert np.isfinite(got).any()


def test_set_rect_image_then_software_start(u, fe640, oracle, golden):
    """Fpga::setRectImage writes the RECT bank at call time (FPGA.cpp:236-249); FPGA_XSBL_SW_START runs xsbl -> bm on it"""
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_bm_params(x_store_offset=1)
    L, R = golden["rect_l"].copy(), golden["rect_r"].copy()
    fe640.submit_rect(0, R, L); fe640.wait()                        # leave other contents in bank 0 first
    fe640.set_rect_image(0, L, R)
    keepL, keepR = L.copy(), R.copy()
    L[:] = 0; R[:] = 0                                               # the caller's buffers are free again
    rl, rr = fe640.receive_rect(0)
    assert np.array_equal(rl[0], keepL) and np.array_equal(rr[0], keepR)
    with pytest.raises(u.U96Error) as e:
        fe640.receive_disp(0)                                        # nothing has run on this bank yet
    assert e.value.code == -4
    with pytest.raises(u.U96Error):
        fe640.wait()                                                 # and nothing is in flight
    fe640.start_xsbl(0)
    assert fe640.wait() == 0
    want = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64)
    assert np.array_equal(fe640.receive_disp(0)[0], want)
    xl, xr = fe640.receive_xsbl(0)
    assert np.array_equal(xl[0], golden["xsbl_l"]) and np.array_equal(xr[0], golden["xsbl_r"])
    f = u.Fpga(0)
    assert f.registerOpen() == 0
    f.setRectImage(1, keepL, keepR)
    a, b2 = f.receiveRectImages(1)
    assert np.array_equal(a, keepL) and np.array_equal(b2, keepR)
    f.start(1)
    bank, _, _, dm = f.receiveData()
    assert bank == 1 and np.array_equal(dm, want)
    f.registerClose()


def test_configuration_is_locked_while_a_bank_is_in_flight_and_banks_keep_their_geometry(u, oracle):
    """the setters refuse between submit and wait (a pending bank reads the map / tile plan / parameters); a
    bank is received with the geometry it was filled with even after the configuration moved on."""
    L, R = u.synth_batch(2, 0, 3, 640, 480, 64)
    with u.StereoFrontEnd(0, 640, 480, 3) as fe:
        fe.set_bm_params(width=640, height=480, profile=0, block_size=15, num_disparities=64, x_store_offset=1)
        fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
        fe.submit_raw(0, L, R)
        for call in (lambda: fe.set_bm_params(block_size=21), lambda: fe.set_rect_params(u.SHIPPED_RECT_PARAMS),
                     lambda: fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0), lambda: fe.set_stream(0)):
            with pytest.raises(u.U96Error) as e:
                call()
            assert e.value.code == -4
        assert fe.get_bm_params()["block_size"] == 15
        assert fe.wait() == 0
        want = [oracle.bm_rtl(oracle.xsobel_rtl(oracle.rectify(L[i], u.SHIPPED_RECT_PARAMS, 0)),
                              oracle.xsobel_rtl(oracle.rectify(R[i], u.SHIPPED_RECT_PARAMS, 1)), wsz=15, ndisp=64) for i in range(3)]
        fe.set_bm_params(width=320, height=240, block_size=9, num_disparities=32)       # idle now: accepted
        l2, r2 = u.synth_batch(3, 0, 2, 320, 240, 32)
        fe.submit_rect(1, l2, r2); assert fe.wait() == 1
        d0 = fe.receive_disp(0)                                      # bank 0 still holds three 640x480 maps
        assert d0.shape == (3, 480, 640)
        for i in range(3):
            assert np.array_equal(d0[i], want[i])
        rl, _ = fe.receive_rect(0)
        assert rl.shape == (3, 480, 640) and np.array_equal(rl[1], oracle.rectify(L[1], u.SHIPPED_RECT_PARAMS, 0))
        d1 = fe.receive_disp(1)
        assert d1.shape == (2, 240, 320)
        assert np.array_equal(d1[0], oracle.bm_rtl(oracle.xsobel_rtl(l2[0]), oracle.xsobel_rtl(r2[0]), wsz=9, ndisp=32))


def test_error_behaviour(u):
    with u.StereoFrontEnd(0, 320, 240, 2) as fe:
        with pytest.raises(u.U96Error) as e:
            fe.wait()
        assert e.value.code == -4                                   # nothing submitted
        with pytest.raises(u.U96Error):
            fe.set_bm_params(width=320, height=240, block_size=20)   # even block size
        with pytest.raises(u.U96Error):
            fe.set_bm_params(width=320, height=240, profile=0, block_size=21, num_disparities=48)   # RTL: multiple of 32
        with pytest.raises(u.U96Error):
            fe.set_bm_params(width=640, height=480)                  # larger than the handle
        fe.set_bm_params(width=320, height=240, profile=0, block_size=9, num_disparities=32)
        L, R = u.synth_batch(5, 0, 2, 320, 240, 32)
        with pytest.raises(u.U96Error) as e:
            fe.submit_raw(0, L, R)                                   # rect parameters never set
        assert e.value.code == -4
        fe.submit_rect(0, L, R)
        with pytest.raises(u.U96Error) as e:
            fe.submit_rect(0, L, R)                                  # bank still in flight
        assert e.value.code == -4
        assert fe.wait() == 0
        with pytest.raises(u.U96Error):
            fe.submit_rect(1, np.concatenate([L, L]), np.concatenate([R, R]))   # batch > max_batch
        assert fe.kernel_launches() > 0


def test_device_resident_submit_is_zero_copy_and_equal(u, oracle):
    import torch
    L, R = u.synth_batch(1, 0, 2, 640, 480, 64)
    with u.StereoFrontEnd(0, 640, 480, 2) as fe:
        fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
        fe.set_bm_params(x_store_offset=1)
        fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
        fe.set_stream(torch.cuda.current_stream().cuda_stream)
        dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        fe.submit_device("raw", 0, dL.data_ptr(), dR.data_ptr(), 640, 2); b = fe.wait()
        d_dev = fe.receive_disp(b)
        fe.submit_raw(1, L, R); b = fe.wait()
        assert np.array_equal(fe.receive_disp(b), d_dev)


def test_cpp_host_shim(u):
    """host/Fpga.hpp + fpga_test_main.cpp: the reference's FPGA_TEST loop in C++ over the C ABI."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "u96_slam_b200", "host", "fpga_test")
    if not os.path.exists(exe):
        sys.path.insert(0, root)
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("kpts3d")]
    assert len(lines) == 3
    for i, ln in enumerate(lines):
        kf, ln = ln.split(";", 1)
        kf = kf.split()                                              # generateKeypoints3D (Stereo.cpp:119-154) through Fpga.hpp
        assert int(kf[3]) > 100 and int(kf[5]) == int(kf[3]) and int(kf[1]) > 0.5 * int(kf[3])
        f = ln.split()
        assert int(f[3]) == i % 2                                    # bank = iteration % 2 (main.cpp:168)
        assert float(f[f.index("frac_at_12px") + 1]) > 0.95          # random texture shifted by 12 px
        assert int(f[-1]) > 10000                                    # dense points from the x4-decimated map
        assert 0 < int(f[f.index("eig_max") + 1]) <= 0xFFFF and int(f[f.index("candidates") + 1]) > 1000   # receiveEigen (FPGA.cpp:281-296)


def test_async_receive_pipeline(u, oracle):
    import torch
    L, R = u.synth_batch(1, 0, 2, 640, 480, 64)
    with u.StereoFrontEnd(0, 640, 480, 1) as fe:
        fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
        fe.set_bm_params(x_store_offset=1)
        out = [torch.empty((1, 480, 640), dtype=torch.int16).pin_memory() for _ in range(2)]
        for b in range(2):
            fe.submit_rect(b, L[b], R[b])
            fe.enqueue_receive_disp_ptr(b, out[b].data_ptr())
        assert fe.wait() == 0 and fe.wait() == 1
        for b in range(2):
            want = oracle.bm_rtl(oracle.xsobel_rtl(L[b]), oracle.xsobel_rtl(R[b]), wsz=21, ndisp=64)
            assert np.array_equal(out[b][0].numpy(), want)


def test_pipelined_async_submit_matches(u, oracle):
    """u96_submit_raw_async: chunked H2D / kernels / D2H pipeline gives the same bits as the plain path."""
    import torch
    L, R = u.synth_batch(1, 0, 4, 640, 480, 64)
    n = 310                                                      # three chunks (two BM waves = 148 frames each, then 14) over the sub-streams
    reps = (n + 3) // 4
    hL = torch.from_numpy(np.concatenate([L] * reps)[:n]).pin_memory(); hR = torch.from_numpy(np.concatenate([R] * reps)[:n]).pin_memory()
    out = torch.empty((n, 480, 640), dtype=torch.int16).pin_memory()
    with u.StereoFrontEnd(0, 640, 480, n) as fe:
        fe.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
        fe.set_bm_params(x_store_offset=1)
        fe.set_rect_params(u.SHIPPED_RECT_PARAMS)
        fe.set_gftt(True)                                        # the eigen map is produced per chunk as well
        fe.submit_host_ptr_async("raw", 0, hL.data_ptr(), hR.data_ptr(), 640, n, out.data_ptr())
        assert fe.wait() == 0
        ref = fe.receive_disp(0)
        assert np.array_equal(out.numpy(), ref)
        e, m = fe.receive_eigen(0)
        for i in (0, 1, 147, 148, 149, 295, 296, 309):
            j = i % 4
            rl, rr = oracle.rectify(L[j], u.SHIPPED_RECT_PARAMS, 0), oracle.rectify(R[j], u.SHIPPED_RECT_PARAMS, 1)
            assert np.array_equal(out[i].numpy(), oracle.bm_rtl(oracle.xsobel_rtl(rl), oracle.xsobel_rtl(rr), wsz=21, ndisp=64)), i
            we, wm = oracle.gftt_eig(rl)
            assert np.array_equal(e[i], we) and int(m[i]) == wm, i


def test_staged_host_transfers_of_unaligned_width(u, oracle):
    """Width that is not a multiple of the internal pitch (KITTI-like 330 px here): host transfers take the tight device
    staging area (one contiguous PCIe copy + on-device pitch conversion) in the plain, the chunk-pipelined and the receive
    paths; every one must deliver the same bits as the oracle."""
    import torch
    W, H, D, B = 330, 96, 64, 9
    L, R = u.synth_batch(7, 0, 4, W, H, D)
    rp = u.identity_rect_params(W, H, float(W))
    n = 70                                                       # >= 64 frames: the chunk pipeline is taken
    reps = (n + 3) // 4
    hL = torch.from_numpy(np.concatenate([L] * reps)[:n]).pin_memory(); hR = torch.from_numpy(np.concatenate([R] * reps)[:n]).pin_memory()
    out = torch.empty((n, H, W), dtype=torch.int16).pin_memory()
    with u.StereoFrontEnd(0, W, H, n) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1)
        fe.set_rect_params(rp)
        fe.submit_host_ptr_async("raw", 0, hL.data_ptr(), hR.data_ptr(), W, n, out.data_ptr())
        assert fe.wait() == 0
        assert np.array_equal(out.numpy(), fe.receive_disp(0))   # staged D2H of the pipeline == staged D2H of receive
        rl_all, rr_all = fe.receive_rect(0)
        xl_all, xr_all = fe.receive_xsbl(0)
        for i in (0, 1, 2, 3, 63, 64, 69):
            j = i % 4
            rl, rr = oracle.rectify(L[j], rp, 0), oracle.rectify(R[j], rp, 1)
            assert np.array_equal(rl_all[i], rl) and np.array_equal(rr_all[i], rr), i
            xl, xr = oracle.xsobel_rtl(rl), oracle.xsobel_rtl(rr)
            assert np.array_equal(xl_all[i], xl) and np.array_equal(xr_all[i], xr), i
            assert np.array_equal(out[i].numpy(), oracle.bm_rtl(xl, xr, wsz=B, ndisp=D)), i
        # plain (non-pipelined) path: a short batch from pageable memory
        fe.submit_raw(1, L[:3], R[:3])
        assert fe.wait() == 1
        d = fe.receive_disp(1)
        for j in range(3):
            rl, rr = oracle.rectify(L[j], rp, 0), oracle.rectify(R[j], rp, 1)
            assert np.array_equal(d[j], oracle.bm_rtl(oracle.xsobel_rtl(rl), oracle.xsobel_rtl(rr), wsz=B, ndisp=D)), j


@pytest.mark.parametrize("W,B", [(361, 25), (360, 25), (360, 27)])
def test_bm_window_blocks_at_the_tile_edge(u, oracle, W, B):
    """wsz 25 / 27 on the 5-warp tile with 134-136 centres in the last tile: the window sum of the last valid segment is built
    from whole blocks that cover one column past the tile (subtracted again by the fix-up).  Regression test for the
    racecheck finding (pad columns behind the column-sum buffers); repeated launches, bit-exact every time."""
    H, D = 72, 64
    L, R = u.synth_batch(11, 0, 2, W, H, D)
    with u.StereoFrontEnd(0, W, H, 8) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1)
        xl = np.stack([oracle.xsobel_rtl(L[0]), oracle.xsobel_rtl(L[1])] * 4)
        xr = np.stack([oracle.xsobel_rtl(R[0]), oracle.xsobel_rtl(R[1])] * 4)
        want = [oracle.bm_rtl(xl[j], xr[j], wsz=B, ndisp=D) for j in range(2)]
        for rep in range(6):
            fe.submit_xsbl(rep & 1, xl, xr)
            assert fe.wait() == (rep & 1)
            d = fe.receive_disp(rep & 1)
            for i in range(8):
                assert np.array_equal(d[i], want[i & 1]), (rep, i)


def test_c5_slam_loop_octomap(u, oracle, tmp_path):
    """BASELINE config C5: BM disparity -> x4 decimation -> reprojectTo3D -> pose -> OctoMap (main.cpp:495-561) through
    host/slam_loop (C++ shim + the reference's vendored OctoMap), point set checked against the oracle."""
    import os, subprocess, struct
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "u96_slam_b200", "host", "slam_loop")
    if not os.path.exists(exe):
        pytest.skip("slam_loop not built (needs the reference's OctoMap sources at build time)")
    N = 5
    seq = tmp_path / "seq.bin"
    frames = [u.synth_pair(1, i, 640, 480, 64, x_drift=1) for i in range(N)]
    with open(seq, "wb") as f:
        f.write(struct.pack("<3i", 640, 480, N))
        for L, R in frames:
            f.write(L.tobytes()); f.write(R.tobytes())
    step = (0.0, -0.05, 0.0)
    out = subprocess.run([exe, str(seq), str(tmp_path / "slam.bt")] + [str(s) for s in step], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    tok = out.stdout.split()
    got = {tok[i]: tok[i + 1] for i in range(0, len(tok), 2)}
    assert os.path.getsize(tmp_path / "slam.bt") > 1000 and int(got["leaf_nodes"]) > 1000
    # oracle side of the same loop
    sx, sy = 640.0 / 1241, 480.0 / 376
    P_l = np.array([[718.856 * sx, 0, 607.1928 * sx, 0], [0, 718.856 * sy, 185.2157 * sy, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * sx
    finite = inserted = 0
    checksum = 0.0
    for i, (L, R) in enumerate(frames):
        d = oracle.bm_rtl(oracle.xsobel_rtl(L), oracle.xsobel_rtl(R), wsz=21, ndisp=64)
        pts = oracle.reproject(d, P_l, P_r, 4, 1).reshape(-1, 3)
        ok = np.isfinite(pts).all(axis=1)
        finite += int(ok.sum())
        p = pts[ok]
        o = np.array([np.float32(step[k]) * np.float32(i) for k in range(3)], np.float32)
        w = (p + o).astype(np.float32)
        v = (w - o).astype(np.float32)
        rng = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]).astype(np.float32).astype(np.float64))
        keep = rng <= 25.0
        inserted += int(keep.sum())
        wk = w[keep].astype(np.float64)
        checksum += float((wk[:, 0] + 2.0 * wk[:, 1] + 3.0 * wk[:, 2]).sum())
    assert int(got["finite_points"]) == finite
    assert abs(int(got["inserted"]) - inserted) <= 2                  # float sqrt at exactly the 25 m gate
    assert abs(float(got["checksum"]) - checksum) <= 1e-6 * abs(checksum) + 200.0


def test_opencv_postfilters(u, fe640, golden, cv_golden, oracle):
    """cv::StereoBM exactly as configured at main.cpp:198-212 (validateDisparity + filterSpeckles)."""
    fe640.set_bm_params(width=640, height=480, profile=u.PROFILE_OPENCV, num_disparities=64, block_size=21, texture_threshold=10,
                        uniqueness_ratio=10, prefilter_cap=31, min_disparity=0, disp12_max_diff=1, speckle_window_size=50, speckle_range=32)
    fe640.submit_rect(0, golden["rect_l"], golden["rect_r"]); b = fe640.wait()
    got = fe640.receive_disp(b)[0]
    assert np.array_equal(got, cv_golden["maincpp_postfilter"])          # cv2.StereoBM on the reference's bundled pair
    L, R = u.synth_batch(1, 3, 3, 640, 480, 64)
    for d12, sw, sr in ((1, 0, 0), (-1, 50, 32), (2, 120, 16), (0, 30, 48)):
        fe640.set_bm_params(disp12_max_diff=d12, speckle_window_size=sw, speckle_range=sr, block_size=15)
        fe640.submit_rect(1, L, R); b = fe640.wait()
        got = fe640.receive_disp(b)
        for i in (0, 2):
            want = oracle.bm_cv_post(oracle.xsobel_cv(L[i]), oracle.xsobel_cv(R[i]), wsz=15, ndisp=64, disp12_max_diff=d12,
                                     speckle_window=sw, speckle_range=sr)
            assert np.array_equal(got[i], want), (d12, sw, sr, i, int((got[i] != want).sum()))
    fe640.set_bm_params(disp12_max_diff=-1, speckle_window_size=0, speckle_range=0)


@pytest.mark.parametrize("W,H,D,B,n", [(640, 150, 64, 21, 1), (523, 101, 64, 17, 5), (640, 90, 64, 31, 8), (700, 133, 128, 21, 2),
                                        (640, 84, 64, 21, 3), (640, 83, 64, 21, 3)])
def test_saturating_chain_in_y_bands(u, oracle, W, H, D, B, n):
    """A handful of pairs with the 10-bit saturating column sums (bm_calc_sad.v:449-466): the chain over the image height is cut into
    16-row bands -- band functions (chain from 0, chain from 1023, sum of n - o), k_bm_chain, bands from their exact start states
    (bm_fused.cuh).  Every frame == oracle on frames that saturate and recover (stripes of full contrast in the middle third, so the
    deficit of a clipped sum has to be carried across band boundaries); heights around the 64-row threshold take both paths."""
    L, R = u.synth_batch(31, 0, n, W, H, D)
    L[:, H // 3: 2 * H // 3] = np.where(L[:, H // 3: 2 * H // 3] > 127, 255, 0).astype(np.uint8)
    with u.StereoFrontEnd(0, W, H, n) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1, uni_enable=0)
        for bank in (0, 1, 0):                                        # the scratch of both banks, and its reuse
            fe.submit_rect(bank, L, R)
            b = fe.wait()
            d = fe.receive_disp(b); xl, xr = fe.receive_xsbl(b)
    sat = 0
    for i in range(n):
        want = oracle.bm_rtl(xl[i], xr[i], wsz=B, ndisp=D)
        sat += oracle.sat_events()
        assert np.array_equal(d[i], want), f"frame {i}: {int((d[i] != want).sum())} pixels"
    assert sat > 0


@pytest.mark.parametrize("D,B,cap", [(128, 31, 33), (128, 31, 34), (256, 31, 33), (128, 21, 63), (64, 31, 33)])
def test_opencv_profile_column_sums_at_the_fp16_limit(u, oracle, D, B, cap):
    """The cv::StereoBM variants with 16-bit staged rows keep their column sums as fp16 bit patterns (bm_fused.cuh), exact below 2048:
    window x 2 cap = 2046 is the largest configuration that takes them (cap 34: 2108 falls back to byte rows).  Vertical stripes
    of full contrast drive |x-Sobel| to the clip in every row, so whole columns of the cost volume sit at window x 2 cap."""
    W, H = 448, 72
    rng = np.random.default_rng(5)
    cols = np.repeat(rng.integers(0, 2, W // 2 + 1), 2)[:W] * 255                 # stripes two pixels wide, random phase
    L = np.broadcast_to(cols.astype(np.uint8), (H, W)).copy()
    R = np.roll(255 - L, 3, axis=1)
    L[::9, ::5] ^= 0x40; R[::7, ::3] ^= 0x20                                      # a little texture so that not every pixel is a tie
    with u.StereoFrontEnd(0, W, H, 1) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_OPENCV, num_disparities=D, block_size=B, texture_threshold=0,
                         uniqueness_ratio=3, prefilter_cap=cap, min_disparity=0, disp12_max_diff=-1, speckle_window_size=0)
        fe.submit_rect(0, L[None], R[None])
        b = fe.wait()
        got = fe.receive_disp(b)[0]; xl, xr = fe.receive_xsbl(b)
    assert int(np.abs(xl[0].astype(int) - xr[0].astype(int)).max()) == 2 * cap      # the clip is reached
    want = oracle.bm_cv(xl[0], xr[0], wsz=B, ndisp=D, prefilter_cap=cap, texture_threshold=0, uniqueness_ratio=3)
    assert np.array_equal(got, want), int((got != want).sum())


@pytest.mark.parametrize("W,H", [(332, 70), (330, 64), (648, 50), (2056, 40)])
def test_opencv_postfilters_ragged_widths(u, oracle, W, H):
    """validateDisparity + filterSpeckles on widths that take the other code paths of postfilter.cu: a multiple of 4 but not of 8 (scalar
    row pass, four-pixel passes), not a multiple of 4 (scalar everywhere), a multiple of 8 (vector row pass), wider than 2048."""
    D, n = 64, 2
    L, R = u.synth_batch(13, 2, n, W, H, D)
    with u.StereoFrontEnd(0, W, H, n) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_OPENCV, num_disparities=D, block_size=11, texture_threshold=5,
                         uniqueness_ratio=5, prefilter_cap=31, min_disparity=0, disp12_max_diff=1, speckle_window_size=40, speckle_range=24)
        fe.submit_rect(0, L, R)
        got = fe.receive_disp(fe.wait())
    for i in range(n):
        want = oracle.bm_cv_post(oracle.xsobel_cv(L[i]), oracle.xsobel_cv(R[i]), wsz=11, ndisp=D, texture_threshold=5, uniqueness_ratio=5,
                                 disp12_max_diff=1, speckle_window=40, speckle_range=24)
        assert np.array_equal(got[i], want), (i, int((got[i] != want).sum()))


def test_stereobm_facade_maincpp_configuration(u, golden, cv_golden):
    bm = u.StereoBM.create(16, 9)                                         # main.cpp:201
    bm.setPreFilterCap(31); bm.setBlockSize(21); bm.setMinDisparity(0); bm.setNumDisparities(64)
    bm.setTextureThreshold(10); bm.setUniquenessRatio(10)
    bm.setSpeckleWindowSize(50); bm.setSpeckleRange(32); bm.setDisp12MaxDiff(1)
    assert np.array_equal(bm.compute(golden["rect_l"], golden["rect_r"]), cv_golden["maincpp_postfilter"])


@pytest.mark.parametrize("D,B,T,U,cap", [(64, 21, 10, 10, 31), (64, 9, 0, 0, 31), (64, 5, 25, 30, 63), (128, 15, 10, 15, 31),
                                         (128, 31, 0, 5, 20), (256, 21, 10, 10, 31), (256, 7, 3, 40, 31)])
def test_opencv_profile_fast_path_slices(u, fe640, golden, oracle, D, B, T, U, cap):
    """cv::StereoBM profile on the cluster-sliced kernel (numDisparities = 64 x cluster size): winner, exact uniqueness
    and sub-pixel neighbours across slice boundaries, texture lane."""
    for src in ("golden", "synth"):
        L, R = (golden["rect_l"], golden["rect_r"]) if src == "golden" else u.synth_pair(5, D, 640, 480, min(D, 200))
        fe640.set_bm_params(width=640, height=480, profile=u.PROFILE_OPENCV, num_disparities=D, block_size=B, texture_threshold=T,
                            uniqueness_ratio=U, prefilter_cap=cap, min_disparity=0, disp12_max_diff=-1, speckle_window_size=0)
        fe640.submit_rect(0, L, R)
        d = fe640.receive_disp(fe640.wait())[0]
        want = oracle.bm_cv(oracle.xsobel_cv(L, cap), oracle.xsobel_cv(R, cap), wsz=B, ndisp=D, prefilter_cap=cap, texture_threshold=T, uniqueness_ratio=U)
        assert np.array_equal(d, want), (src, int((d != want).sum()))


@pytest.mark.parametrize("D,B,thr", [(128, 21, 921), (128, 9, 700), (256, 21, 921), (256, 15, 500)])
def test_bm_rtl_uniqueness_across_slices(u, fe640, golden, oracle, D, B, thr):
    """RTL profile with the uniqueness filter on: the approximate min2 travels through every dphase merge
    (bm_calc_upd.v), i.e. across the cluster's slice records."""
    ext = int(D > 128)
    d = run_xsbl(fe640, 0, golden["xsbl_l"], golden["xsbl_r"],
                 **dict(RTL, num_disparities=D, block_size=B, rtl_extended=ext, uni_enable=1, uni_thr=thr))[0]
    want = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=B, ndisp=D, rtl_extended=ext, uni_enb=1, uni_thr=thr)
    assert np.array_equal(d, want), int((d != want).sum())


def test_cluster_ring_under_scheduling_noise(u, oracle):
    """the DSMEM record ring of the cluster kernels (D = 128 / 256) releases its slots with a relaxed remote arrive.
    Stress it: many launches of both cluster sizes and both profiles while a second handle keeps unrelated kernels running on
    other streams (CTAs of a cluster then drift apart as their SMs are shared unevenly); every frame of every launch must equal
    the oracle / the first launch bit for bit."""
    W, H = 600, 200
    noise = u.StereoFrontEnd(0, 640, 480, 24)
    noise.set_bm_params(width=640, height=480, profile=0, block_size=9, num_disparities=64, x_store_offset=1)
    nL, nR = u.synth_batch(9, 0, 24, 640, 480, 64)
    try:
        for D, prof in ((128, 0), (256, 0), (128, 1), (256, 1)):
            L, R = u.synth_batch(4, D, 6, W, H, D)
            with u.StereoFrontEnd(0, W, H, 6) as fe:
                if prof == 0:
                    fe.set_bm_params(width=W, height=H, profile=0, block_size=21, num_disparities=D, uni_enable=1, uni_thr=900, x_store_offset=1, rtl_extended=1)
                    want = [oracle.bm_rtl(oracle.xsobel_rtl(L[i]), oracle.xsobel_rtl(R[i]), wsz=21, ndisp=D, uni_enb=1, uni_thr=900, rtl_extended=1,
                                          bitserial_div=0) for i in range(6)]
                else:
                    fe.set_bm_params(width=W, height=H, profile=1, block_size=15, num_disparities=D, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10)
                    want = [oracle.bm_cv(oracle.xsobel_cv(L[i]), oracle.xsobel_cv(R[i]), wsz=15, ndisp=D) for i in range(6)]
                for it in range(40):
                    noise.submit_rect(it & 1, nL, nR)                 # in flight on its own stream while the cluster kernel runs
                    fe.submit_rect(it & 1, L, R)
                    b = fe.wait()
                    d = fe.receive_disp(b)
                    noise.wait()
                    for i in range(6):
                        assert np.array_equal(d[i], want[i]), (D, prof, it, i)
    finally:
        noise.close()


def test_fast_and_generic_bm_kernels_agree(u, monkeypatch):
    """The generic kernel (U96_BM_GENERIC=1) and the fast path produce identical maps on a ragged width."""
    W, H, D = 1000, 131, 128
    L, R = u.synth_batch(9, 0, 3, W, H, D)
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("U96_BM_GENERIC", "1")
        with u.StereoFrontEnd(0, W, H, 3) as fe:
            res = []
            for prof in (u.PROFILE_RTL, u.PROFILE_OPENCV):
                fe.set_bm_params(width=W, height=H, profile=prof, num_disparities=D, block_size=11, x_store_offset=1,
                                 texture_threshold=10, uniqueness_ratio=10, prefilter_cap=31, uni_enable=1, uni_thr=900)
                fe.submit_rect(0, L, R)
                res.append(fe.receive_disp(fe.wait()))
            outs.append(res)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("W,H,D,B,n", [(640, 56, 64, 21, 24), (500, 40, 64, 17, 20), (640, 44, 64, 31, 20), (523, 48, 64, 15, 20),
                                        (700, 50, 128, 21, 6), (610, 40, 256, 19, 3)])
def test_fused_bm_kernel_every_frame_of_a_batch(u, oracle, W, H, D, B, n):
    """k_bm_fused with the 16-bit staged rows and the fp16 oldest-row step (bm_fused.cuh), in batches beyond the <= 18-pair rule that keeps
    a handful of saturating 64-disparity pairs on k_bm_fast: EVERY frame == oracle, on frames whose column sums hit the 10-bit
    ceiling (bm_calc_sad.v:449-466) and on ragged widths; the same frames through U96_BM_FUSED=0 agree as well (same process: the
    switch is read once, so the comparison runs through the generic kernel instead)."""
    L, R = u.synth_batch(21, 0, n, W, H, D)
    L[:, H // 4: 3 * H // 4] = np.where(L[:, H // 4: 3 * H // 4] > 127, 255, 0).astype(np.uint8)      # strong edges: saturating sums
    with u.StereoFrontEnd(0, W, H, n) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1,
                         rtl_extended=int(D > 128), uni_enable=0)
        fe.submit_rect(0, L, R)
        b = fe.wait()
        d = fe.receive_disp(b); xl, xr = fe.receive_xsbl(b)
    sat = 0
    for i in range(n):
        want = oracle.bm_rtl(xl[i], xr[i], wsz=B, ndisp=D, rtl_extended=int(D > 128))
        sat += oracle.sat_events()
        assert np.array_equal(d[i], want), f"frame {i}"
    if B * 63 > 1023:
        assert sat > 0


def test_uvc_payload(u, fe640, golden, oracle):
    """u96_receive_uvc == the firmware's UVC frame (xusb_main.c:293-376) for the RECT, XSBL and BM streams."""
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    L2 = np.stack([golden["rect_l"], golden["rect_r"]]); R2 = np.stack([golden["rect_r"], golden["rect_l"]])
    fe640.submit_rect(0, L2, R2)
    b = fe640.wait()
    sl, sr = fe640.receive_xsbl(b); d = fe640.receive_disp(b)
    for i in range(2):
        assert np.array_equal(fe640.receive_uvc(b, u.UVC_RECT)[i], oracle.pack_uvc(1, L2[i], R2[i]))
        assert np.array_equal(fe640.receive_uvc(b, u.UVC_XSBL)[i], oracle.pack_uvc(2, sl[i], sr[i]))
        assert np.array_equal(fe640.receive_uvc(b, u.UVC_BM)[i], oracle.pack_uvc(3, disp=d[i]))
    # ragged width (W not a multiple of 8)
    W, H = 650, 37
    L, R = u.synth_batch(4, 0, 1, W, H, 64)
    with u.StereoFrontEnd(0, W, H, 1) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=9, num_disparities=64, x_store_offset=1)
        fe.submit_rect(0, L, R); b = fe.wait()
        assert np.array_equal(fe.receive_uvc(b, u.UVC_BM)[0], oracle.pack_uvc(3, disp=fe.receive_disp(b)[0]))
        assert np.array_equal(fe.receive_uvc(b, u.UVC_RECT)[0], oracle.pack_uvc(1, L[0], R[0]))


def test_gftt_min_eigenvalue_map(u, fe640, golden, oracle):
    """u96_set_gftt + u96_receive_eigen == the oracle's reading of dvp/rtl/gftt*.v (map and gftt.Max), on the bundled pair,
    through the raw path (GFTT reads the RECT bank), in a batch, and on ragged shapes."""
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.set_gftt(True)
    try:
        fe640.submit_rect(0, golden["rect_l"], golden["rect_r"]); b = fe640.wait()
        e, m = fe640.receive_eigen(b)
        want, wmax = oracle.gftt_eig(golden["rect_l"])
        assert np.array_equal(e[0], want) and int(m[0]) == wmax
        d = fe640.receive_disp(b)[0]                                   # the BM result is unaffected
        assert np.array_equal(d, oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64))
    finally:
        fe640.set_gftt(False)
    with pytest.raises(u.U96Error):                                    # bank filled without GFTT: nothing to receive
        fe640.submit_rect(1, golden["rect_l"], golden["rect_r"]); fe640.wait(); fe640.receive_eigen(1)
    rng = np.random.default_rng(11)
    for (W, H, n) in ((640, 480, 5), (53, 37, 3), (130, 9, 2), (1242, 375, 2)):
        L = rng.integers(0, 256, (n, H, W), dtype=np.uint8); L[0] = (np.kron(rng.integers(0, 2, (H // 3 + 1, W // 3 + 1)), np.ones((3, 3))) * 255).astype(np.uint8)[:H, :W]
        R = rng.integers(0, 256, (n, H, W), dtype=np.uint8)
        with u.StereoFrontEnd(0, W, H, n) as fe:
            D = 32 if W < 200 else 64
            fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=5, num_disparities=D, x_store_offset=1)
            fe.set_gftt(True)
            fe.submit_rect(0, L, R); b = fe.wait()
            e, m = fe.receive_eigen(b)
            for i in range(n):
                want, wmax = oracle.gftt_eig(L[i])
                assert np.array_equal(e[i], want), (W, H, i)
                assert int(m[i]) == wmax
    # sensor path: the map is computed from the RECTIFIED left image
    L, R = u.synth_batch(1, 0, 2, 640, 480, 64)
    fe640.set_rect_params(u.SHIPPED_RECT_PARAMS); fe640.set_gftt(True)
    try:
        fe640.submit_raw(0, L, R); b = fe640.wait()
        rl, _ = fe640.receive_rect(b); e, m = fe640.receive_eigen(b)
        for i in range(2):
            want, wmax = oracle.gftt_eig(rl[i])
            assert np.array_equal(e[i], want) and int(m[i]) == wmax
    finally:
        fe640.set_gftt(False)


def test_randomised_shapes_and_parameters_both_profiles(u, oracle):
    """Seeded sweep over ragged widths/heights, window sizes, disparity ranges (fast path, cluster path and the generic
    kernel), uniqueness settings and both profiles, on noise images (ties, saturation, flat areas): bit-exact every time."""
    rng = np.random.default_rng(20261017)
    cases = []
    for _ in range(14):
        D = int(rng.choice([32, 64, 96, 128, 160, 256]))
        B = int(rng.choice([5, 7, 9, 11, 15, 17, 21, 25, 31]))
        W = int(D + B + 2 + rng.integers(3, 260)); H = int(B + rng.integers(1, 70))
        cases.append((W, H, D, B))
    cases += [(64 + 21 + 2, 21, 64, 21), (64 + 31 + 3, 31, 64, 31), (777, 33, 128, 9)]       # minimal valid sizes, one centre column
    for i, (W, H, D, B) in enumerate(cases):
        kind = i % 3
        if kind == 0:                                   # noise
            L = rng.integers(0, 256, (2, H, W), dtype=np.uint8); R = rng.integers(0, 256, (2, H, W), dtype=np.uint8)
        elif kind == 1:                                 # flat areas + a few edges: ties everywhere
            L = np.full((2, H, W), 90, np.uint8); R = np.full((2, H, W), 90, np.uint8)
            L[:, :, W // 3:] = 200; R[:, :, W // 2:] = 200; L[1, H // 2:] = 10
        else:                                           # 0/255 blocks: column sums saturate at 10 bit
            L = (np.kron(rng.integers(0, 2, (2, H // 2 + 1, W // 2 + 1)), np.ones((2, 2))) * 255).astype(np.uint8)[:, :H, :W]
            R = np.roll(L, -int(rng.integers(0, D)), axis=2)
        with u.StereoFrontEnd(0, W, H, 2) as fe:
            uni = int(rng.integers(0, 2)); thr = int(rng.integers(0, 1024)); mode = int(rng.integers(0, 2)); ext = int(D > 128 or rng.integers(0, 2))
            xo = int(rng.integers(0, 2))
            fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, uni_enable=uni, uni_thr=thr,
                             uni_mode=mode, rtl_extended=ext, x_store_offset=xo)
            fe.submit_rect(0, L, R); b = fe.wait()
            d = fe.receive_disp(b); sl, sr = fe.receive_xsbl(b)
            for k in range(2):
                xl, xr = oracle.xsobel_rtl(L[k]), oracle.xsobel_rtl(R[k])
                assert np.array_equal(sl[k], xl) and np.array_equal(sr[k], xr), ("xsbl", W, H)
                want = oracle.bm_rtl(xl, xr, wsz=B, ndisp=D, uni_enb=uni, uni_thr=thr, uni_mode=mode, rtl_extended=ext, x_store_offset=xo)
                assert np.array_equal(d[k], want), ("rtl", W, H, D, B, uni, thr, mode, ext, xo, int((d[k] != want).sum()))
            if B >= 5 and B * B * 62 <= 65535:
                cap = int(rng.integers(1, 32)); tex = int(rng.integers(0, 40)); uq = int(rng.integers(0, 30))
                fe.set_bm_params(profile=u.PROFILE_OPENCV, prefilter_cap=cap, texture_threshold=tex, uniqueness_ratio=uq)
                fe.submit_rect(1, L, R); b = fe.wait()
                d = fe.receive_disp(b)
                for k in range(2):
                    want = oracle.bm_cv(oracle.xsobel_cv(L[k], cap), oracle.xsobel_cv(R[k], cap), wsz=B, ndisp=D, prefilter_cap=cap,
                                        texture_threshold=tex, uniqueness_ratio=uq)
                    assert np.array_equal(d[k], want), ("cv", W, H, D, B, cap, tex, uq, int((d[k] != want).sum()))


def test_zero_copy_disparity_view(u, fe640, golden):
    """StereoFrontEnd.disp_tensor: a torch view of the DISP bank in HBM (the send buffer of the optional multi-GPU gather)
    holds exactly what u96_receive_disp copies out, also for a pitch wider than the image."""
    import torch
    fe640.set_bm_registers((480 << 16) + 640, 0x00150040, 0)
    fe640.submit_rect(0, np.stack([golden["rect_l"]] * 2), np.stack([golden["rect_r"]] * 2)); b = fe640.wait()
    t = fe640.disp_tensor(b)
    assert t.is_cuda and t.dtype == torch.int16 and tuple(t.shape) == (2, 480, 640)
    assert np.array_equal(t.cpu().numpy(), fe640.receive_disp(b))
    W, H = 650, 40                                               # pitch 768 > W: strided view
    L, R = u.synth_batch(4, 0, 2, W, H, 64)
    with u.StereoFrontEnd(0, W, H, 2) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=9, num_disparities=64, x_store_offset=1)
        fe.submit_rect(0, L, R); b = fe.wait()
        t = fe.disp_tensor(b)
        assert not t.is_contiguous() and np.array_equal(t.contiguous().cpu().numpy(), fe.receive_disp(b))
