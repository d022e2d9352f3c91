"""CPU tests: the oracle against every pin the reference offers -- no GPU needed."""
import os
import zlib

import numpy as np
import pytest

from oracle_py import SHIPPED_RECT, REF_LIB_PATH, REF_STEREO_LIB_PATH, RefFpga


def test_kat1_xsobel_matches_reference_golden(oracle, golden):
    # data/ref_rect_{l,r} -> data/ref_xsbl_{l,r}: the only golden pair the reference ships (xsbl2.v)
    for side in "lr":
        got = oracle.xsobel_rtl(golden["rect_" + side])
        assert np.array_equal(got, golden["xsbl_" + side])
    assert golden["xsbl_l"][0].max() == 0 and golden["xsbl_l"][-1].max() == 0      # invalid lines stay 0
    assert (golden["xsbl_l"][1:-1, 0] == 32).all() and (golden["xsbl_l"][1:-1, -1] == 32).all()


def test_opencv_prefilter_differs_from_rtl(oracle, golden):
    # cap-31 rule mismatches the RTL goldens on most pixels
    cv = oracle.xsobel_cv(golden["rect_l"], 31)
    assert (cv != golden["xsbl_l"]).mean() > 0.5


def test_rect_remap_matches_reference_function(oracle):
    # fixture = output of the reference's own rect_remap() (fpga.c:303-366) compiled into oracle/_ref
    m = np.load(os.path.join(os.path.dirname(__file__), "golden", "rect_remap_ref.npz"))
    for lr, n in ((0, "l"), (1, "r")):
        xs, ys = oracle.rect_remap(SHIPPED_RECT, lr, 640, 480)
        assert np.array_equal(xs, m["xs_" + n]) and np.array_equal(ys, m["ys_" + n])
    # shipped parameters map inside the source image
    assert xs.min() >= 0 and xs.max() < 640 * 32 and ys.min() >= 0 and ys.max() < 480 * 32


@pytest.mark.skipif(not os.path.exists(REF_LIB_PATH), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_command_stream_reproduces_shipped_cmd_dat(oracle):
    """KAT from the reference's own assets: rect_remap() + rect_cmd_gen() of the reference's fpga.c, compiled
    from where it lies and run on the shipped parameter set, regenerate src/dvp/sim/cmd.dat (17 104 words) exactly once the
    firmware's 4 KiB-boundary burst split (issue_cmd, fpga.c:519-541, absent from the testbench dump) is undone.  The
    stream encodes, per destination row, the source-row transitions and span lengths of the inverse map, so this pins the
    compiled reference map -- which the oracle's orc_rect_remap equals bit for bit (previous tests) -- to a reference file."""
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "rect_cmd_dat.npz"))["words"]
    got = RefFpga().rect_cmd_stream(SHIPPED_RECT, 640, 480)
    merged, splits, ydst = [], 0, 0
    for v in (int(x) for x in got):
        if not v & 0xFF:                                # Y command: {lr[17], ydst[16:8]}
            ydst = (v >> 8) & 0x1FF
            merged.append(v)
        elif (merged[-1] & 0xFF) and not (v >> 7) & 1 and (v >> 8) == (merged[-1] >> 8) + (merged[-1] & 0x7F) \
                and (ydst * 640 + (v >> 8)) % 4096 == 0:
            merged[-1] += v & 0x7F                      # second half of a burst that crossed a 4 KiB boundary
            splits += 1
        else:
            ydst += (v >> 7) & 1                        # ydif: the destination row advances (cmd_x, fpga.c:572-603)
            merged.append(v)
    assert len(got) == len(want) + splits and splits > 0
    assert np.array_equal(np.array(merged, np.uint32), want)
    # the implicit contract the GPU path keeps from the command stream: every destination pixel exactly once
    assert sum(v & 0x7F for v in merged if v & 0xFF) == 2 * 640 * 480
    # and the map the stream was generated from is the oracle's map
    ref = RefFpga().rect_remap(SHIPPED_RECT, 640, 480)
    for lr in (0, 1):
        xs, ys = oracle.rect_remap(SHIPPED_RECT, lr, 640, 480)
        assert np.array_equal(xs, ref[lr][0]) and np.array_equal(ys, ref[lr][1])


@pytest.mark.skipif(not os.path.exists(REF_LIB_PATH), reason="oracle/_ref not built (no /root/reference here)")
def test_rect_remap_live_reference_random_params(oracle):
    from oracle_py import RefFpga
    ref = RefFpga()
    rng = np.random.default_rng(7)
    for _ in range(3):
        d = {k: (np.array(v) + rng.integers(-2000, 2000, np.array(v).shape)).tolist() for k, v in SHIPPED_RECT.items()}
        d["c"] = [320, 240]
        d["f2inv"] = [d["f2inv"][0]] * 2 if False else d["f2inv"]
        # the reference C code reads f2inv/c2_f2/c per channel, the RTL shares channel 0's: feed identical values
        got = [oracle.rect_remap(d, lr, 160, 120) for lr in (0, 1)]
        want = ref.rect_remap(d, 160, 120)
        for lr in (0, 1):
            assert np.array_equal(got[lr][0], want[lr][0]) and np.array_equal(got[lr][1], want[lr][1])


def test_rect_interp_identity_and_clamp(oracle):
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (24, 40), dtype=np.uint8)
    ys, xs = np.meshgrid(np.arange(24) * 32, np.arange(40) * 32, indexing="ij")
    assert np.array_equal(oracle.rect_interp(src, xs, ys), src)                  # integer coordinates: identity
    out = oracle.rect_interp(src, xs + 16, ys)                                   # half-pixel: rounded mean
    want = ((src[:, :-1].astype(int) * 16 * 32 + src[:, 1:].astype(int) * 16 * 32 >> 9) + 1) >> 1
    assert np.array_equal(out[:, :-1], want)
    assert (oracle.rect_interp(src, xs - 64, ys) [:, :1] == 0).all()             # outside taps read 0


def test_kat3_diven_closed_forms(oracle):
    rng = np.random.default_rng(3)
    # rect: diven#(26,26,26,24)(2^24, lw) == floor(2^48/lw)
    for lw in rng.integers(1 << 22, 1 << 25, 3000):
        assert oracle.diven(26, 26, 26, 24, 1 << 24, int(lw)) == ((1 << 48) // int(lw)) & ((1 << 26) - 1)
    # sub-pixel: diven#(18,18,8,17)(n, d) == floor(128 n / d) for d > 0, |n| <= d/2, n of either sign.
    # (bm_calc_frac.v:80-95: a negative divisor implies neg_val, i.e. dividend 0 -- checked below;
    #  for n != 0 with d < 0 the non-restoring divider is NOT floor, but that input cannot occur.)
    for _ in range(20000):
        d = int(rng.integers(1, 1 << 16)) * 2
        n = int(rng.integers(-(d // 2), d // 2 + 1))
        q = oracle.diven(18, 18, 8, 17, n & 0x3FFFF, d & 0x3FFFF)
        assert q == ((128 * n) // d) & 0xFF, (n, d, q)
    # 0 / +-d -> 0
    for d in (2, -2, 500, -131070):
        assert oracle.diven(18, 18, 8, 17, 0, d & 0x3FFFF) == 0
    # uniqueness: diven#(17,17,11,16)(a, b) == floor(1024 a / b) (a <= b), b == 0 -> 2047
    for _ in range(20000):
        b = int(rng.integers(1, 1 << 16)); a = int(rng.integers(0, b + 1))
        assert oracle.diven(17, 17, 11, 16, a, b) == (1024 * a // b) & 0x7FF
    # min2 == 0 implies min1 == 0 (min1 <= min2 always): the only reachable zero-divisor case
    assert oracle.diven(17, 17, 11, 16, 0, 0) == 2047
    assert oracle.diven(17, 17, 11, 16, 77, 77) == 1024          # equal minima: & 0x3FF -> 0 -> passes the filter


# CRCs of an independent numpy reading of the RTL (non-authoritative cross-check)
@pytest.mark.parametrize("wsz,uni,thr,valid,total,crc", [
    (21, 0, 0, 253088, 194655560, 0x3C312D26),
    (21, 1, 921, 157291, 133771940, 0xF3284A7C),
    (15, 0, 0, 259128, 190344441, 0xD0650EA3),
])
def test_bm_rtl_cross_check_with_survey(oracle, golden, wsz, uni, thr, valid, total, crc):
    d = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=wsz, ndisp=64, uni_enb=uni, uni_thr=thr)
    assert int((d >= 0).sum()) == valid and int(d.astype(np.int64).sum()) == total
    assert zlib.crc32(d.tobytes()) & 0xFFFFFFFF == crc
    # closed-form divisions == bit-serial diven on real data
    d2 = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=wsz, ndisp=64, uni_enb=uni, uni_thr=thr, bitserial_div=0)
    assert np.array_equal(d, d2)


def _rtl_bm_inputs(kind, W, H, D, seed):
    import u96_slam_b200 as u
    from oracle_py import Oracle
    rng = np.random.default_rng(seed)
    if kind == "synth":                                              # textured pair with a block-wise disparity field
        L, R = u.synth_pair(seed, 1, W, H, D)
        o = Oracle()
        return o.xsobel_rtl(L), o.xsobel_rtl(R)
    if kind == "flat":                                               # low-amplitude noise: ties, adjacent minima, zero SADs, guard-lane wins
        base = rng.integers(30, 34, (H, W + D)).astype(np.uint8)
        return np.ascontiguousarray(base[:, D:]), np.ascontiguousarray(base[:, D - 3:W + D - 3] ^ (rng.random((H, W)) < 0.02))
    # saturating: full-swing 6-bit noise (column sums hit 1023 at wsz 21) with a flat band
    a = rng.integers(0, 64, (H, W)).astype(np.uint8); b = rng.integers(0, 64, (H, W)).astype(np.uint8)
    a[H // 4:3 * H // 4] = 63; b[H // 4:3 * H // 4] = 0
    return a, b


@pytest.mark.parametrize("D,W,H", [(64, 200, 70), (128, 300, 60), (256, 420, 56), (32, 120, 50), (96, 240, 48), (160, 300, 48)])
@pytest.mark.parametrize("wsz", [5, 15, 21])
def test_bm_rtl_oracle_agrees_with_independent_numpy_reading(oracle, D, W, H, wsz):
    """a6-a14 beyond D = 64 rest on more than one reader.  tests/rtl_bm_numpy.py was written from the
    Verilog alone; it must equal the C oracle for every dphase count 1..8, uniqueness off / on in both modes, both store
    offsets and both output extensions, on textured, tie-heavy and saturating inputs."""
    from rtl_bm_numpy import bm_rtl_numpy
    n_uni = 0
    for kind in ("synth", "flat", "sat"):
        xl, xr = _rtl_bm_inputs(kind, W, H, D, 100 + D + wsz)
        for (enb, mode, thr, xo, ext) in ((0, 0, 0, 1, 0), (1, 0, 921, 1, int(D > 128)), (1, 1, 600, 0, 1), (1, 0, 0, 1, 1)):
            want = oracle.bm_rtl(xl, xr, wsz=wsz, ndisp=D, uni_enb=enb, uni_mode=mode, uni_thr=thr, x_store_offset=xo, rtl_extended=ext)
            got = bm_rtl_numpy(xl, xr, wsz, D, enb, mode, thr, xo, ext)
            assert np.array_equal(got, want), (kind, enb, mode, thr, xo, ext, int((got != want).sum()))
            if enb and kind == "synth":
                off = oracle.bm_rtl(xl, xr, wsz=wsz, ndisp=D, x_store_offset=xo, rtl_extended=ext)
                n_uni += int((off != want).sum())
        if kind == "sat" and wsz == 21:
            assert oracle.sat_events() > 0                           # the 10-bit ceiling really was hit
    assert n_uni > 0                                                 # the uniqueness filter really changed pixels


def test_bm_rtl_independent_reading_on_the_bundled_pair_all_ranges(oracle, golden):
    """The same cross-check on the reference's own ref_xsbl images at D = 64 / 128 / 256, wsz 15 / 21."""
    from rtl_bm_numpy import bm_rtl_numpy
    xl, xr = golden["xsbl_l"][100:260], golden["xsbl_r"][100:260]    # a 160-row band keeps the CPU suite short
    for D in (64, 128, 256):
        for wsz in (15, 21):
            for (enb, thr) in ((0, 0), (1, 921)):
                want = oracle.bm_rtl(xl, xr, wsz=wsz, ndisp=D, uni_enb=enb, uni_thr=thr, rtl_extended=1)
                assert np.array_equal(bm_rtl_numpy(xl, xr, wsz, D, enb, 0, thr, 1, 1), want), (D, wsz, enb)
                assert (want >= 0).mean() > 0.05


def test_bm_rtl_saturation_facts(oracle, golden):
    # 10-bit column-sum saturation happens on ref_xsbl at wsz 21, never at wsz <= 16
    oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64)
    assert oracle.sat_events() > 0
    oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=15, ndisp=64)
    assert oracle.sat_events() == 0


def test_bm_rtl_layout(oracle, golden):
    d = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64, x_store_offset=1)
    assert (d[:10] == -1).all() and (d[-10:] == -1).all()            # hwsz invalid rows top/bottom
    assert (d[:, :64 + 10 + 1] == -1).all() and (d[:, -10:] == -1).all()   # ndisp+hwsz+1 leading, hwsz trailing
    d0 = oracle.bm_rtl(golden["xsbl_l"], golden["xsbl_r"], wsz=21, ndisp=64, x_store_offset=0)
    assert np.array_equal(d0[:, :-1], d[:, 1:])                      # A1: same values one column to the left
    assert d.max() < 64 * 16 + 8


def test_bm_cv_matches_cv2_golden(oracle, golden, cv_golden):
    pl, pr = oracle.xsobel_cv(golden["rect_l"], 31), oracle.xsobel_cv(golden["rect_r"], 31)
    for k, want in cv_golden.items():
        if not k.startswith("D"):
            continue
        D, B, T, U = [int(s[1:]) for s in k.split("_")]
        got = oracle.bm_cv(pl, pr, wsz=B, ndisp=D, texture_threshold=T, uniqueness_ratio=U)
        assert got.dtype == np.int16 and np.array_equal(got, want), k
    assert (cv_golden["D64_B21_T10_U10"] == -16).any()               # KAT-4: invalid = -16


def test_bm_cv_matches_live_cv2_on_synthetic(oracle):
    cv2 = pytest.importorskip("cv2")
    import u96_slam_b200 as u
    L, R = u.synth_pair(2, 5, 320, 120, 48)
    for D, B, T, U in ((48, 9, 10, 15), (32, 15, 0, 0)):
        bm = cv2.StereoBM_create(D, B)
        bm.setPreFilterCap(25); bm.setTextureThreshold(T); bm.setUniquenessRatio(U)
        bm.setSpeckleWindowSize(0); bm.setDisp12MaxDiff(-1); bm.setMinDisparity(0)
        want = bm.compute(L, R)
        got = oracle.bm_cv(oracle.xsobel_cv(L, 25), oracle.xsobel_cv(R, 25), wsz=B, ndisp=D, prefilter_cap=25,
                           texture_threshold=T, uniqueness_ratio=U)
        assert np.array_equal(got, want)


def test_reproject_matches_float_formula(oracle):
    rng = np.random.default_rng(5)
    disp = rng.integers(-16, 64 * 16, (48, 64)).astype(np.int16)
    P_l = np.array([[718.856 * 640 / 1241, 0, 607.19 * 640 / 1241, 0], [0, 718.856 * 480 / 376, 185.22 * 480 / 376, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * 640 / 1241
    for decim in (1, 4):
        out = oracle.reproject(disp, P_l, P_r, decim, 0)
        d = disp[::decim, ::decim].astype(np.float32) / np.float32(16)
        bad = d <= 0
        assert np.isnan(out[bad]).all()
        c = np.float32(P_r[0, 2] - P_l[0, 2])
        dc = (d + c).astype(np.float32)
        Wx = ((P_l[0, 3] / P_l[0, 0] - P_r[0, 3] / P_r[0, 0]) / dc.astype(np.float64)).astype(np.float32)
        Z = (P_l[0, 0] * Wx.astype(np.float64)).astype(np.float32)
        u0 = (np.arange(out.shape[1]) * decim).astype(np.float32)[None, :]
        X = ((u0.astype(np.float64) - P_l[0, 2]) * Wx.astype(np.float64)).astype(np.float32)
        assert np.array_equal(out[..., 2][~bad], Z[~bad]) and np.array_equal(out[..., 0][~bad], X[~bad])
        loc = oracle.reproject(disp, P_l, P_r, decim, 1)
        assert np.array_equal(loc[..., 0][~bad], out[..., 2][~bad]) and np.array_equal(loc[..., 1][~bad], -out[..., 0][~bad])


def _reproject_fixture():
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_golden import reproject_inputs
    g = np.load(os.path.join(here, "golden", "reproject_ref.npz"))
    return reproject_inputs(), {k: g[k] for k in g.files}


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_reprojection_oracle_matches_reference_fixture(oracle):
    """a16 pinned: the oracle equals the outputs of the reference's own Stereo.cpp / StereoCameraModel.cpp / Transform.cpp
    (compiled unmodified, fixture made by tests/golden/make_golden.py) bit for bit, NaNs and signed zeros included."""
    import u96_slam_b200 as u
    (W, H, disp, uv, P_l, P_r, pose, gates), g = _reproject_fixture()
    assert np.array_equal(g["local_transform"], u.LOCAL_TRANSFORM)
    for rs in (0, 1):
        Pl, Pr = g[f"P_l_resize{rs}"], g[f"P_r_resize{rs}"]
        for gi, (mn, mx) in enumerate(gates):
            want = g[f"kp_resize{rs}_gate{gi}"]
            assert _same_bits(oracle.reproject_points(disp, Pl, Pr, uv, mn, mx), want)
            assert np.isfinite(want[:, 0]).sum() > (0 if (rs, gi) == (0, 3) else 200)
        assert _same_bits(oracle.reproject_ex(disp, Pl, Pr, 4, u.LOCAL_TRANSFORM, None), g[f"dense_resize{rs}_local"])
        assert _same_bits(oracle.reproject_ex(disp, Pl, Pr, 4, u.LOCAL_TRANSFORM, pose), g[f"dense_resize{rs}_local_pose"])
    assert _same_bits(oracle.reproject_ex(disp, g["P_l_resize0"], g["P_r_resize0"], 1, None, None), g["dense_resize0_plain"])
    # the legacy entry point = localTransform through transformPoint
    assert _same_bits(oracle.reproject(disp, g["P_l_resize1"], g["P_r_resize1"], 4, 1), g["dense_resize1_local"])


def test_kitti_projection_loader_matches_reference_fixture(tmp_path):
    """formats.load_projection_kitti == StereoCameraModel::load (KITTI route + 640x480 rescale, StereoCameraModel.cpp:68-119)"""
    import u96_slam_b200 as u
    from oracle_py import RefStereo
    (W, H, disp, uv, P_l, P_r, pose, gates), g = _reproject_fixture()
    calib = str(tmp_path / "calib.txt")
    RefStereo.write_kitti_calib(calib, P_l, P_r)
    for rs in (0, 1):
        Pl, Pr = u.load_projection_kitti(calib, do_resize=bool(rs))
        for mine, ref in ((Pl, g[f"P_l_resize{rs}"]), (Pr, g[f"P_r_resize{rs}"])):
            for (i, j) in ((0, 0), (1, 1), (0, 2), (1, 2), (0, 3)):      # the ten getters of StereoCameraModel.h:24-33
                assert mine[i, j] == ref[i, j]


@pytest.mark.skipif(not os.path.exists(REF_STEREO_LIB_PATH), reason="oracle/_ref/libstereo_ref.so not built (reference tree absent)")
def test_reprojection_oracle_matches_live_compiled_reference(oracle, tmp_path):
    """The same pin on fresh random inputs: oracle == the reference's compiled functions, called live."""
    from oracle_py import RefStereo, LOCAL_TRANSFORM
    r = RefStereo()
    t, is_null = r.local_transform()
    assert not is_null and np.array_equal(t, LOCAL_TRANSFORM)
    rng = np.random.default_rng(11)
    for trial in range(3):
        W, H = (640, 480) if trial == 0 else (int(rng.integers(40, 300)), int(rng.integers(30, 200)))
        fx = float(rng.uniform(300, 900)); fy = fx * float(rng.uniform(0.9, 1.1))
        P_l = np.array([[fx, 0, W * rng.uniform(0.4, 0.6), 0], [0, fy, H * rng.uniform(0.4, 0.6), 0], [0, 0, 1, 0]], np.float64)
        P_r = P_l.copy(); P_r[0, 3] = -fx * rng.uniform(0.05, 0.6); P_r[0, 2] += (0.0 if trial < 2 else 3.5)
        calib = str(tmp_path / f"calib{trial}.txt")
        r.write_kitti_calib(calib, P_l, P_r)
        Pl, Pr = r.model_load(calib, trial == 1)
        disp = rng.integers(-64, 256 * 16, (H, W)).astype(np.int16)
        disp[rng.random((H, W)) < 0.3] = -1
        n = 12000 if trial == 0 else 3000
        uv = np.stack([rng.random(n) * W, rng.random(n) * H], 1).astype(np.float32)
        uv[:3] = [[0, 0], [W - 0.001, H - 0.001], [-0.25, -0.75]]
        for mn, mx in ((0.0, 0.0), (-1.0, 0.0), (1.0, 20.0), (0.0, 3.0)):
            assert _same_bits(oracle.reproject_points(disp, Pl, Pr, uv, mn, mx), r.keypoints3d(calib, trial == 1, uv, disp, mn, mx))
        pose = np.concatenate([rng.normal(size=(3, 3)), rng.normal(size=(3, 1)) * 3], 1).astype(np.float32).reshape(12)
        dd = np.ascontiguousarray(disp[::4, ::4][:H // 4, :W // 4])
        assert _same_bits(oracle.reproject_ex(disp, Pl, Pr, 4, LOCAL_TRANSFORM, pose), r.dense_cloud(calib, trial == 1, dd, 4, True, pose))
        assert _same_bits(oracle.reproject_ex(disp, Pl, Pr, 1, None, None), r.dense_cloud(calib, trial == 1, disp, 1, False, None))


def test_postfilters_match_opencv_public_functions(oracle, golden, cv_golden):
    """f1 row: validateDisparity + filterSpeckles as cv::StereoBM::compute applies them (main.cpp:210-212)."""
    cv2 = pytest.importorskip("cv2")
    import ctypes
    from oracle_py import BmCvParams
    pl, pr = oracle.xsobel_cv(golden["rect_l"], 31), oracle.xsobel_cv(golden["rect_r"], 31)
    # (a) the reference's exact configuration on the reference's bundled pair: bit-exact vs cv2.StereoBM
    assert np.array_equal(oracle.bm_cv_post(pl, pr), cv_golden["maincpp_postfilter"])
    # (b) the two filters against OpenCV's public functions on the same inputs
    H, W = pl.shape
    p = BmCvParams(21, 64, 31, 10, 10)
    disp = np.empty((H, W), np.int16); cost = np.zeros((H, W), np.int16)
    i16p, u8p = ctypes.POINTER(ctypes.c_int16), ctypes.POINTER(ctypes.c_uint8)
    oracle.L.orc_bm_cv_cost(pl.ctypes.data_as(u8p), pr.ctypes.data_as(u8p), W, H, ctypes.byref(p), disp.ctypes.data_as(i16p), cost.ctypes.data_as(i16p))
    mine = disp.copy(); oracle.L.orc_validate_disparity(mine.ctypes.data_as(i16p), cost.ctypes.data_as(i16p), W, H, 0, 64, 1)
    theirs = disp.copy(); cv2.validateDisparity(theirs, cost, 0, 64, 1)
    assert np.array_equal(mine, theirs) and (mine != disp).sum() > 1000
    for size, diff in ((50, 32), (200, 16), (10, 64)):
        a = disp.copy(); oracle.L.orc_filter_speckles(a.ctypes.data_as(i16p), W, H, -16, size, diff)
        b = disp.copy(); cv2.filterSpeckles(b, -16, size, diff)
        assert np.array_equal(a, b) and (a != disp).sum() > 100


def test_dat_writer_reproduces_reference_files_byte_for_byte(golden):
    """formats.write_dat(golden image) == the reference's data/ref_*.dat text (SHA-256 recorded by make_golden.py);
    read_dat is its inverse."""
    import hashlib
    import json
    import u96_slam_b200.formats as fm
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dat_sha256.json")))
    for key, name in (("rect_l", "ref_rect_l"), ("rect_r", "ref_rect_r"), ("xsbl_l", "ref_xsbl_l"), ("xsbl_r", "ref_xsbl_r")):
        b = fm.write_dat(golden[key])
        assert len(b) == want[name]["bytes"] and hashlib.sha256(b).hexdigest() == want[name]["sha256"], name
        assert np.array_equal(fm.read_dat(b, width=640), golden[key])


def test_uvc_payload_oracle_follows_firmware_loops(oracle, golden):
    """orc_pack_uvc against a literal numpy transcription of Xusb_ReceiveData's index arithmetic (xusb_main.c:313-372)."""
    L, R = golden["rect_l"], golden["rect_r"]
    f = oracle.pack_uvc(1, L, R)
    flat = f.reshape(-1)
    rows, cols = np.mgrid[0:480, 0:640]
    for lr, img in ((0, L), (1, R)):
        dst = (rows * 1280 + cols + lr * 640) * 2
        assert np.array_equal(flat[dst], img) and (flat[dst + 1] == 0x80).all()
    d = (np.arange(480 * 640, dtype=np.int32).reshape(480, 640) * 37 % 2200 - 100).astype(np.int16)
    d[::7, ::5] = -1
    f = oracle.pack_uvc(3, disp=d).reshape(-1)
    dl, dr = (rows * 1280 + cols) * 2, (rows * 1280 + cols + 640) * 2
    assert np.array_equal(f[dl], (d >> 4).astype(np.uint8)) and (f[dr] == 0).all() and (f[dl + 1] == 0x80).all() and (f[dr + 1] == 0x80).all()


def _rect_intp_numpy(src, xs, ys):
    """A second reading of rect_intp.v:288-412, written from the Verilog alone and vectorised (the C oracle walks pixel by pixel):
    u10.5 source coordinates -> four taps (outside the image: 0), weights xf / 32-xf / yf / 32-yf as u1.5, products u1.10, taps x
    weights u8.10, the two row sums, their sum, bits [17:9] + 1, bit 9 set -> 0xFF else bits [8:1]."""
    H, W = src.shape
    xs = xs.astype(np.int64); ys = ys.astype(np.int64)
    xi, yi, xf, yf = xs >> 5, ys >> 5, xs & 31, ys & 31

    def tap(y, x):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return np.where(ok, src[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)].astype(np.int64), 0)
    ul_, ur_, dl_, dr_ = tap(yi, xi), tap(yi, xi + 1), tap(yi + 1, xi), tap(yi + 1, xi + 1)
    xfi, yfi = 32 - xf, 32 - yf
    ulr = (ul_ * (xfi * yfi) + ur_ * (xf * yfi)) & 0x3FFFF                      # ulr_lim[17:0]
    dlr = (dl_ * (xfi * yf) + dr_ * (xf * yf)) & 0x3FFFF
    udlr = ulr + dlr
    rnd = ((udlr >> 9) & 0x1FF) + 1                                             # udlr_lim[8:0] + 1
    return np.where(rnd & 0x200, 255, (rnd >> 1) & 0xFF).astype(np.uint8)


def test_rect_interpolation_oracle_agrees_with_independent_numpy_reading(oracle, golden):
    """a3 has no reference input/output pair: the C oracle is cross-checked with a second, vectorised reading of the RTL on
    the shipped map (keystone, both cameras), on random maps that leave the image, and at the rounding / 0xFF limiter corner."""
    import u96_slam_b200 as u
    rng = np.random.default_rng(11)
    src = golden["rect_l"]
    H, W = src.shape
    for cam in (0, 1):
        xs, ys = oracle.rect_remap(u.SHIPPED_RECT_PARAMS, cam, W, H)
        assert np.array_equal(oracle.rect_interp(src, xs, ys), _rect_intp_numpy(src, xs, ys))
    noise = rng.integers(0, 256, (H, W), dtype=np.uint8)
    xs = (np.arange(W)[None, :] * 32 + rng.integers(-200, 200, (H, W))).astype(np.int16)
    ys = (np.arange(H)[:, None] * 32 + rng.integers(-200, 200, (H, W))).astype(np.int16)
    assert np.array_equal(oracle.rect_interp(noise, xs, ys), _rect_intp_numpy(noise, xs, ys))
    white = np.full((H, W), 255, np.uint8)                                       # 255 x 1.0 = 0x3FC00: bits [17:9] + 1 stays below bit 9
    assert (_rect_intp_numpy(white, xs % (32 * 8), ys % (32 * 8))[:4, :4] == 255).all()
    assert np.array_equal(oracle.rect_interp(white, xs, ys), _rect_intp_numpy(white, xs, ys))


def test_rect_registers_from_calibration(oracle):
    """rect_params_from_calibration: identity calibration reproduces the near-identity register set, and for a rotated rig
    the fixed-point map of rect_remap (fpga.c:303-366) follows K R^T K'^-1 to within the u10.5 output resolution."""
    import u96_slam_b200 as u
    W, H, f = 640, 480, 700.0
    ident = u.rect_params_from_calibration([(f, f, W / 2, H / 2)] * 2, [np.eye(3)] * 2, (f, f, W / 2, H / 2))
    assert ident == u.identity_rect_params(W, H, f)

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return Rz @ Ry @ Rx
    K = [(690.0, 705.0, 318.0, 236.0), (702.0, 699.0, 318.0, 236.0)]
    Rr = [rot(0.01, -0.02, 0.015), rot(-0.008, 0.012, -0.02)]
    Kn = (650.0, 650.0, 322.0, 241.0)
    p = u.rect_params_from_calibration(K, Rr, Kn)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    for cam in range(2):
        mx, my = oracle.rect_remap(p, cam, W, H)
        v = np.stack([(xs - Kn[2]) / Kn[0], (ys - Kn[3]) / Kn[1], np.ones_like(xs)], -1) @ Rr[cam]      # R^T applied to each ray
        wx = v[..., 0] / v[..., 2] * K[cam][0] + round(K[0][2]); wy = v[..., 1] / v[..., 2] * K[cam][1] + round(K[0][3])
        assert np.abs(mx / 32.0 - wx).max() < 0.06 and np.abs(my / 32.0 - wy).max() < 0.06


def _gftt_numpy(img):
    """Independent vectorised reading of dvp/rtl/gftt_{sbl,eig,box,obuf}.v (second restatement, int64 numpy)."""
    p = img.astype(np.int64); H, W = p.shape
    dx = np.zeros((H, W), np.int64); dy = np.zeros((H, W), np.int64)
    dx[1:-1, 1:-1] = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    dy[1:-1, 1:-1] = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    ax, ay = np.abs(dx), np.abs(dy)

    def box(v):
        h = np.zeros_like(v)
        h[:, 1:-1] = v[:, :-2] + v[:, 1:-1] + v[:, 2:]          # columns 0 and W-1 forced to 0 (gftt_box.v:185)
        s = np.zeros_like(v)
        s[2:-2] = h[1:-3] + h[2:-2] + h[3:-1]
        return np.minimum(s, 0xFFFF)

    a, c, b = box((ax * ax) >> 6), box((ay * ay) >> 6), box((ax * ay) >> 6)
    s = np.minimum(((a - c) ** 2 >> 10) + ((b * b) >> 8), 0x3FFFFF) << 10
    root = np.floor(np.sqrt(s.astype(np.float64))).astype(np.int64)
    root = np.where(root * root > s, root - 1, root); root = np.where((root + 1) ** 2 <= s, root + 1, root)
    e = (a + c) - (root & 0xFFFF)
    out = np.where(e < 0, 0, np.where(e & 0x10000, 0xFFFF, e)).astype(np.uint16)
    out[:2] = 0; out[-2:] = 0
    return out


def test_gftt_oracle_agrees_with_independent_numpy_reading(oracle, golden):
    """the reference ships no eigen dump (parity unpinned); two independent readings of the RTL agree,
    and the structural facts of the RTL hold: border rows/columns are zero, max = gftt.Max over the written rows."""
    rng = np.random.default_rng(5)
    imgs = [golden["rect_l"], golden["rect_r"], rng.integers(0, 256, (37, 53), dtype=np.uint8),
            (np.kron(rng.integers(0, 2, (14, 22)), np.ones((3, 3))) * 255).astype(np.uint8)]   # 0/255 blocks drive the 16/22-bit limits
    for img in imgs:
        got, mx = oracle.gftt_eig(img)
        assert np.array_equal(got, _gftt_numpy(img))
        assert mx == int(got.max())
        assert not got[:2].any() and not got[-2:].any() and not got[:, 0].any() and not got[:, -1].any()
    assert oracle.gftt_eig(np.full((16, 16), 77, np.uint8))[1] == 0                 # flat image: no corners
    assert oracle.gftt_eig(imgs[3])[1] == 0xFFFF                                    # the output limiter is reached


def test_gftt_map_tracks_opencv_min_eigenvalue(oracle, golden):
    """Sanity against the algorithm the RTL approximates (cv::cornerMinEigenVal, blockSize 3, Sobel 3): the maps are
    strongly rank-correlated on the reference's bundled image (not equal: the RTL drops the sign of dx*dy and truncates)."""
    cv2 = pytest.importorskip("cv2")
    img = golden["rect_l"]
    got = oracle.gftt_eig(img)[0][2:-2, 1:-1].astype(np.float64).ravel()
    ref = cv2.cornerMinEigenVal(img, 3, ksize=3)[2:-2, 1:-1].astype(np.float64).ravel()
    top = ref >= np.quantile(ref, 0.99)
    assert np.mean(got[top] >= np.quantile(got, 0.95)) > 0.9
