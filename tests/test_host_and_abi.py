"""CPU tests: host logic, the C-ABI surface (symbols only -- no compute without a GPU) and the
world_size-2 sharding path over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(libpath):
    hdr = open(os.path.join(ROOT, "include", "u96_stereo.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(u96_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 36, names
    lib = ctypes.CDLL(libpath)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.u96_abi_version.restype = ctypes.c_int
    assert lib.u96_abi_version() == int(re.search(r"#define U96_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "u96_stereo.h")).read()).group(1)) >= 5
    lib.u96_strerror.restype = ctypes.c_char_p
    assert b"fallback" in lib.u96_strerror(-6)


def test_no_cpu_fallback(libpath):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import u96_slam_b200 as u
    with pytest.raises(u.U96Error) as e:
        u.StereoFrontEnd(0, 640, 480, 1)
    assert e.value.code == -6
    assert u.Fpga().registerOpen() == -1            # reference convention: 0 / -1 (FPGA.cpp:27-60)


def test_product_does_not_touch_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the oracle."""
    pkg = os.path.join(ROOT, "u96_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(dp, f)


def test_params_struct_layout():
    from u96_slam_b200.stereo import BmParams, RectParams
    assert ctypes.sizeof(BmParams) == 17 * 4
    assert ctypes.sizeof(RectParams) == (4 + 2 + 2 + 2 + 18) * 4
    p = RectParams.from_dict(__import__("u96_slam_b200").SHIPPED_RECT_PARAMS)
    assert p.f[1][0] == 39609530 and p.rot[1][2][2] == 16568783 and p.c2_f2[1] == 5932596


def test_synth_is_deterministic_and_shaped():
    import u96_slam_b200 as u
    a = u.synth_pair(1, 3, 640, 480, 64); b = u.synth_pair(1, 3, 640, 480, 64); c = u.synth_pair(1, 4, 640, 480, 64)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], c[0])
    assert a[0].shape == (480, 640) and a[0].dtype == np.uint8 and 60 < a[0].mean() < 200
    L, R = u.synth_pair(2, 0, 256, 64, 32)
    # inside a disparity block R[y][x - delta] == L[y][x] up to the +-1 noise added to R
    delta = 4 + ((0 * 7 + 0 * 13 + 0) % 20)
    x = np.arange(40, 60)
    assert np.abs(R[5, x - delta].astype(int) - L[5, x].astype(int)).max() <= 2


def test_shard_frames_partition():
    sys.path.insert(0, ROOT)
    from bench import shard_frames
    for world in (1, 2, 4, 8):
        parts = [shard_frames(37, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


GLOO_WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from bench import shard_frames, gather_disparity
import u96_slam_b200 as u
from oracle_py import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
o = Oracle()
mine = shard_frames(6, rank, world)
# each rank works on its own frames only (no data-path collective); the checksum of checksums is gathered
cs = 0
maps = []
for i in mine:
    L, R = u.synth_pair(1, i, 160, 96, 32)
    d = o.bm_rtl(o.xsobel_rtl(L), o.xsobel_rtl(R), wsz=9, ndisp=32)
    cs += int(d.astype(np.int64).sum())
    maps.append(d)
# the optional exchange step (SURVEY 8e): all disparity maps onto rank 0, re-interleaved into stream order
got = gather_disparity(torch.from_numpy(np.stack(maps)), rank, world)
gsum = -1
if rank == 0:
    full = np.empty((6,) + maps[0].shape, np.int16)
    for r in range(world):
        full[shard_frames(6, r, world)] = got[r].numpy()
    gsum = int(full.astype(np.int64).sum()) + int(full[5].astype(np.int64).sum())      # order-sensitive: frame 5 counted twice
t = torch.tensor([cs, len(mine)], dtype=torch.int64)
dist.barrier()
dist.all_reduce(t)
ms = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)           # bench.py: max-over-ranks time
if rank == 0:
    print(json.dumps({"sum": int(t[0]), "frames": int(t[1]), "max_ms": float(ms[0]), "gathered": gsum}))
dist.destroy_process_group()
"""


def test_frame_sharding_world_size_2_gloo(tmp_path, oracle):
    """N>1 path on CPU: 2 ranks over gloo shard the stream frame-wise; the union equals the 1-rank run."""
    import json
    import u96_slam_b200 as u
    w = tmp_path / "worker.py"
    w.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(w), ROOT],
                         capture_output=True, text=True, env=env, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    want, last = 0, 0
    for i in range(6):
        L, R = u.synth_pair(1, i, 160, 96, 32)
        last = int(oracle.bm_rtl(oracle.xsobel_rtl(L), oracle.xsobel_rtl(R), wsz=9, ndisp=32).astype(np.int64).sum())
        want += last
    assert res == {"sum": want, "frames": 6, "max_ms": 2.0, "gathered": want + last}


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "c2"], capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")


def test_bench_fails_loudly_without_a_gpu_and_numa_binding_is_optional():
    """bench.py's product arm has no CPU fallback; the rank-to-NUMA binding degrades to None when the topology is not exposed."""
    sys.path.insert(0, ROOT)
    import bench
    import torch
    if not torch.cuda.is_available():
        assert bench.bind_to_gpu_numa_node(0) is None        # no CUDA device: nothing to bind to, and no exception
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                             capture_output=True, text=True, timeout=280)
        assert out.returncode != 0 and out.stdout.strip() == "" and "no CPU fallback" in out.stderr


def test_projection_matrix_loaders(tmp_path):
    """KITTI calib.txt and OpenCV-yml projection matrices (StereoCameraModel.cpp:19-122), incl. the 640x480 rescale."""
    import u96_slam_b200 as u
    k = tmp_path / "calib.txt"
    k.write_text("P0: 7.188560000000e+02 0 6.071928000000e+02 0 0 7.188560000000e+02 1.852157000000e+02 0 0 0 1 0\n"
                 "P1: 7.188560000000e+02 0 6.071928000000e+02 -3.861448000000e+02 0 7.188560000000e+02 1.852157000000e+02 0 0 0 1 0\n"
                 "P2: 1 0 0 0 0 1 0 0 0 0 1 0\n")
    Pl, Pr = u.load_projection_kitti(str(k))
    sx, sy = 640.0 / 1241, 480.0 / 376
    assert Pl.shape == (3, 4) and np.isclose(Pl[0, 0], 718.856 * sx) and np.isclose(Pl[1, 2], 185.2157 * sy)
    assert np.isclose(Pr[0, 3], -386.1448 * sx) and Pr[2, 2] == 1.0
    Pl2, _ = u.load_projection_kitti(str(k), do_resize=False)
    assert Pl2[0, 0] == 718.856
    yml = ("%YAML:1.0\n---\nimage_width: 1280\nimage_height: 960\ncamera_name: {n}\n"
           "projection_matrix: !!opencv-matrix\n   rows: 3\n   cols: 4\n   dt: d\n"
           "   data: [ 800., 0., 650., {tx}, 0., 800.,\n       470., 0., 0., 0., 1., 0. ]\n")
    (tmp_path / "l.yml").write_text(yml.format(n="left", tx="0."))
    (tmp_path / "r.yml").write_text(yml.format(n="right", tx="-96."))
    Pl, Pr = u.load_projection_opencv_yml(str(tmp_path / "l.yml"), str(tmp_path / "r.yml"))
    assert np.allclose(Pl, [[400, 0, 325, 0], [0, 400, 235, 0], [0, 0, 1, 0]]) and np.isclose(Pr[0, 3], -48.0)
    (tmp_path / "bad.yml").write_text(yml.format(n="x", tx="0.").replace("cols: 4", "cols: 3"))
    with pytest.raises(ValueError):
        u.load_projection_opencv_yml(str(tmp_path / "bad.yml"), str(tmp_path / "r.yml"))
