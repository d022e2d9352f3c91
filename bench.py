#!/usr/bin/env python
"""bench.py -- throughput of the dense-stereo front end on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4] [--reps R] [--no-configs]

A "step" is one pass of the hot path (rect -> x-Sobel -> SAD block matching -> 16x disparity) over one batch of
synthetic stereo pairs.  Headline workload = BASELINE.json configs[1] (C2): 640x480, 64 disparities, block 21, RTL
profile (what the FPGA computes), full remap+Sobel+BM pipeline.

  value   : frames/s with the batch already resident in HBM (CUDA events on the launching stream): the MEDIAN of R
            repetitions of the K-step block (`reps` holds every block)
  e2e     : frames/s through the C ABI with pinned HOST buffers (H2D + kernels + D2H every step), beside the PCIe ceiling
            of the same transfers issued by ALL ranks at once
  roofline: the BM kernel against the measured integer-pipe issue rate (SURVEY 8d convention) + what ncu says binds it
  configs : the other BASELINE configurations as sub-records of the same line -- c1 (bundled pair, B15 + B21, both profiles,
            parity asserted against the reference's own vectors / cv2 fixtures), c3 (KITTI shape, B 9/15/21), c4 (1080p,
            D256) and `opencv` = the reference's CPU mode like for like (cv::StereoBM profile + validateDisparity +
            filterSpeckles, slam/src/core/main.cpp:197-217) end to end, next to cv2 timed in the same run
  cpu_baseline: the CPU oracle port of the same pipeline and cv2.StereoBM on this box's host cores (bounded sample)

Frames are sharded frame-wise over ranks (one process per GPU, torch.distributed only for the barrier and the
max-over-ranks time): weak scaling, no collective on the data path.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, D, block, seed, frames per step per GPU)
    "c2": dict(W=640, H=480, D=64, B=21, seed=1, batch=296, name="C2 synthetic 640x480 D64 B21 raw->rect->xsbl->bm (RTL profile)"),
    "c3": dict(W=1242, H=375, D=128, B=15, seed=2, batch=148, name="C3 KITTI-shape 1242x375 D128 B15 raw->rect->xsbl->bm (RTL profile)"),
    "c4": dict(W=1920, H=1080, D=256, B=21, seed=3, batch=37, name="C4 1920x1080 D256 B21 raw->rect->xsbl->bm (RTL-extended profile)"),
}
# survey cross-check values of the RTL profile on the bundled pair -- used by the c1 sub-record
C1_CRC = {21: "3c312d26", 15: "d0650ea3"}
# the reference's CPU-mode parameter set (slam/src/core/main.cpp:198-212)
MAINCPP = dict(prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10, disp12_max_diff=1, speckle_window_size=50, speckle_range=32)


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the NUMA node its GPU hangs off, BEFORE the pinned staging buffers are allocated
    (first touch then places them on that node), so that N ranks do not all stream through one memory controller.
    Returns the node, or None when the topology is not exposed (VMs often report -1)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_frames(total, rank, world):
    """Frame indices of `rank` when `total` frames are dealt round-robin over `world` GPUs."""
    return list(range(rank, total, world))


def gather_disparity(t, rank, world, dst=0):
    """The optional exchange step of SURVEY 8(e): every rank's disparity maps (same shape on all ranks) collected on rank
    `dst` -- NCCL send/recv over NVLink on the GPUs, gloo on CPU tensors in the tests.  Returns the list on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    tb = t.contiguous().view(torch.uint8)            # int16 is not a collective dtype (NCCL, gloo): ship the bytes
    out = [torch.empty_like(tb) for _ in range(world)] if rank == dst else None
    dist.gather(tb, out, dst=dst)
    return [o.view(t.dtype) for o in out] if out is not None else None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.p, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_baseline(wl, frames, threads):
    """The oracle port of the same pipeline (rectify + x-Sobel + BM, RTL profile) on host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_py import Oracle
    import u96_slam_b200 as u
    from concurrent.futures import ThreadPoolExecutor
    o = Oracle()
    W, H, D, B = wl["W"], wl["H"], wl["D"], wl["B"]
    rp = u.SHIPPED_RECT_PARAMS if (W, H) == (640, 480) else u.identity_rect_params(W, H, float(W))
    pairs = [u.synth_pair(wl["seed"], i, W, H, D) for i in range(frames)]

    def one(pr):
        L, R = pr
        rl, rr = o.rectify(L, rp, 0), o.rectify(R, rp, 1)
        return o.bm_rtl(o.xsobel_rtl(rl), o.xsobel_rtl(rr), wsz=B, ndisp=D, rtl_extended=int(D > 128), bitserial_div=0)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:          # ctypes releases the GIL: real parallelism
        list(ex.map(one, pairs))
    dt = time.perf_counter() - t0
    return frames / dt


def cpu_cv2_reference(wl, frames, threads, pairs=None):
    """cv::StereoBM exactly as the reference's CPU mode configures it (slam/src/core/main.cpp:198-215),
    i.e. including validateDisparity/filterSpeckles; None when cv2 is not importable."""
    try:
        import cv2
    except ImportError:
        return None
    import u96_slam_b200 as u
    W, H, D, B = wl["W"], wl["H"], wl["D"], wl["B"]
    cv2.setNumThreads(threads)
    bm = cv2.StereoBM_create(16, 9)
    bm.setPreFilterCap(31); bm.setBlockSize(B); bm.setMinDisparity(0); bm.setNumDisparities(D)
    bm.setTextureThreshold(10); bm.setUniquenessRatio(10)
    bm.setSpeckleWindowSize(50); bm.setSpeckleRange(32); bm.setDisp12MaxDiff(1)
    if pairs is None:
        pairs = [u.synth_pair(wl["seed"], i, W, H, D) for i in range(min(frames, 8))]
    bm.compute(*pairs[0])
    t0 = time.perf_counter()
    for i in range(frames):
        bm.compute(*pairs[i % len(pairs)])
    return frames / (time.perf_counter() - t0)


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W, H, D = wl["W"], wl["H"], wl["D"]
    per_step = max(4, min(64, int(2e9 / (W * H * D))))         # bounded sample per step
    fps_list = []
    kind, sample = "reference", ""
    for step in range(args.warmup + args.steps):
        fps = cpu_cv2_reference(wl, per_step, cores)
        if fps is None:
            kind = "port"
            fps = cpu_oracle_baseline(wl, max(cores, 8), cores)
        if step >= args.warmup:
            fps_list.append(fps)
    fps = float(np.mean(fps_list))
    if kind == "reference":
        sample = (f"cv2.StereoBM (OpenCV, the routine the reference CPU mode calls at slam/src/core/main.cpp:198-215, "
                  f"post-filters on) x{per_step} frames/step, setNumThreads({cores})")
    else:
        sample = f"oracle port (rectify+xsobel+bm_rtl), {max(cores, 8)} frames/step on {cores} threads"
    ms = 1e3 * per_step / fps
    line = {"impl": "reference", "metric": "disparity_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "mpix_disp_per_s": fps * W * H * D / 1e6,
            "config": {"workload": wl["name"], "frames_per_step": per_step,
                       "note": "the reference has no CPU implementation of the FPGA (RTL) arithmetic: its CPU mode is cv::StereoBM + post filters "
                               "without run-time rectification; the like-for-like GPU number is `configs.opencv` / `like_for_like` in the b200 line"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state shared by the measured configurations."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import u96_slam_b200 as u
        self.torch, self.dist, self.u, self.args = torch, dist, u, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.int_peak = None                         # T lane-op/s, measured once

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def alu_peak(self):
        if self.int_peak is None:
            from u96_slam_b200.stereo import microbench
            self.int_peak = microbench(1, self.local) / 1e3      # VABSDIFF4 / ALU-pipe issue rate, T lane-op/s
        return self.int_peak


def stats(vals):
    v = sorted(vals)
    med = float(np.median(v))
    return {"median": med, "min": v[0], "max": v[-1], "spread": (v[-1] - v[0]) / med if med else None, "n": len(v)}


def measure(cx, wl, hL, hR, *, profile, entry, steps, warmup, reps, rect_params=None, bm_extra=None, e2e=True, e2e_reps=3,
            stages=True):
    """Resident and end-to-end rates of one configuration on this rank's frames hL/hR ((nb, H, W) u8 each).
    Returns a dict; every time is the max over ranks, every rate the whole job's."""
    torch, u = cx.torch, cx.u
    W, H, D, B = wl["W"], wl["H"], wl["D"], wl["B"]
    nb = hL.shape[0]
    kind = {"raw": "raw", "rect": "rect"}[entry]

    def configure(fe):
        p = dict(width=W, height=H, profile=profile, block_size=B, num_disparities=D, min_disparity=0)
        if profile == u.PROFILE_RTL:
            p.update(uni_enable=0, uni_mode=0, uni_thr=0, x_store_offset=1, rtl_extended=int(D > 128))
        else:
            p.update(prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10, disp12_max_diff=-1, speckle_window_size=0, speckle_range=0)
        p.update(bm_extra or {})
        fe.set_bm_params(**p)
        if entry == "raw":
            fe.set_rect_params(rect_params)

    # ------------------------------------------------------------------ resident
    fe = u.StereoFrontEnd(cx.local, W, H, nb)
    configure(fe)
    stream = torch.cuda.current_stream()
    fe.set_stream(stream.cuda_stream)
    fe.set_profiling(True)
    dL = torch.from_numpy(hL).cuda(); dR = torch.from_numpy(hR).cuda()

    def run_steps(k, stage=None):
        # the two banks are used the way the reference's producer uses them (bank = iteration % 2, main.cpp:168): step i+1 is
        # submitted before step i is waited for, so the host round trip of u96_wait never leaves the GPU idle between steps
        for i in range(k):
            b = i & 1
            if i >= 2:
                assert fe.wait() == b
                if stage is not None:
                    for kk, v in fe.last_stage_ms_ex(b).items():
                        stage[kk] += v
            fe.submit_device(kind, b, dL.data_ptr(), dR.data_ptr(), W, nb)
        for i in range(max(k - 2, 0), k):
            b = fe.wait()
            if stage is not None:
                for kk, v in fe.last_stage_ms_ex(b).items():
                    stage[kk] += v

    run_steps(warmup)
    blocks, stage, launches = [], {k: 0.0 for k in ("h2d", "rect", "gftt", "xsbl", "bm", "post")}, 0
    for r in range(reps):
        cx.barrier()
        l0 = fe.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_steps(steps, stage if stages else None)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = fe.kernel_launches() - l0
        blocks.append(cx.max_over_ranks(e0.elapsed_time(e1)))
    for k in stage:
        stage[k] /= (steps * reps)
    ms_block = stats(blocks)
    frames_block = nb * steps * cx.world
    out = {"value": frames_block / (ms_block["median"] * 1e-3), "ms_per_step": ms_block["median"] / steps,
           "reps": {"block_ms": blocks, **ms_block, "what": f"{reps} repetitions of the {steps}-step block; value = median block"},
           "frames_per_step_per_gpu": nb, "gpu_launches": int(launches), "stage_ms_per_step": stage if stages else None}
    if stages and stage["bm"] > 0:
        peak = cx.alu_peak()
        ops = 6.0 * W * H * D * nb                               # 6 integer lane-ops per pixel-disparity
        out["bm_roofline"] = {"kernel_ms": stage["bm"], "achieved": ops / (stage["bm"] * 1e-3) / 1e12, "peak": peak, "unit": "Tlaneop/s",
                              "frac": ops / (stage["bm"] * 1e-3) / 1e12 / peak, "algorithmic_ops_per_launch": ops}
    keep = fe
    del dL, dR

    # ------------------------------------------------------------------ end to end (host buffers through the C ABI)
    if e2e:
        keep.close()
        fe2 = u.StereoFrontEnd(cx.local, W, H, nb)
        configure(fe2)
        pL = torch.from_numpy(hL).pin_memory(); pR = torch.from_numpy(hR).pin_memory()
        pD = [torch.empty((nb, H, W), dtype=torch.int16).pin_memory() for _ in range(2)]
        in_ptr = (pL.data_ptr(), pR.data_ptr())
        wc_bufs = []
        if cx.args.wc_inputs:
            # write-combined pinned staging for the INPUTS (the host only ever writes them front to back): bypasses the CPU caches,
            # which frees snoop bandwidth on hosts whose DMA reads compete with N ranks' memory traffic
            import ctypes
            lib = u.load_library()
            ptrs = []
            for src in (hL, hR):
                pp = ctypes.c_void_p()
                rc = lib.u96_host_alloc_wc(ctypes.byref(pp), src.nbytes)
                if rc != 0:
                    raise RuntimeError(f"u96_host_alloc_wc: {rc}")
                ctypes.memmove(pp, src.ctypes.data, src.nbytes)
                ptrs.append(pp.value); wc_bufs.append(pp)
            in_ptr = tuple(ptrs)

        def e2e_loop(k):
            # two banks in flight: H2D + kernels + D2H of bank b are queued back to back on its stream, the host only
            # waits for the older bank, so both copy engines and the SMs overlap
            for i in range(k):
                b = i & 1
                if i >= 2:
                    assert fe2.wait() == b
                fe2.submit_host_ptr_async(kind, b, in_ptr[0], in_ptr[1], W, nb, pD[b].data_ptr())
            for _ in range(min(k, 2)):
                fe2.wait()

        e2e_steps = max(4, steps)
        e2e_loop(3)
        vals = []
        for r in range(e2e_reps):
            cx.barrier()
            t0 = time.perf_counter()
            e2e_loop(e2e_steps)
            torch.cuda.synchronize()
            dt = cx.max_over_ranks(time.perf_counter() - t0)
            vals.append(nb * e2e_steps * cx.world / dt)
        checksum = int(pD[0][0].to(torch.int64).sum().item())
        fe2.close()
        for pp in wc_bufs:
            u.load_library().u96_host_free(pp)
        st = stats(vals)
        out["e2e"] = {"value": st["median"], "unit": "frames/s", "h2d_bytes_per_step": int(2 * W * H * nb),
                      "d2h_bytes_per_step": int(2 * W * H * nb), "steps": e2e_steps, "reps": vals, "spread": st["spread"], "checksum": checksum,
                      "input_memory": "write-combined pinned" if cx.args.wc_inputs else "pinned",
                      "timing": "wall clock around u96_submit_*_async/u96_wait over two banks, synchronize on both sides, max over ranks; median of the repetitions"}
        out["_pinned"] = (pL, pR, pD)
    else:
        keep.close()
    return out


def pcie_all_ranks(cx, pL, pR, pD):
    """PCIe ceiling of the end-to-end loop: the same pinned buffers copied both ways at once by EVERY rank at the same
    time (barrier on both sides), no kernels."""
    torch = cx.torch
    dI = torch.empty_like(pL, device="cuda"); dO = torch.empty_like(pD[0], device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    best = 0.0
    for _ in range(3):
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(s_in):
                dI.copy_(pL, non_blocking=True); dI.copy_(pR, non_blocking=True)
            with torch.cuda.stream(s_out):
                pD[1].copy_(dO, non_blocking=True)
        torch.cuda.synchronize()
        dt = cx.max_over_ranks(time.perf_counter() - t0)          # the slowest rank closes the step, like in the e2e loop
        best = max(best, 2 * 2 * pL.numel() / dt / 1e9)
    del dI, dO
    return best                                                   # GB/s per direction per rank, all ranks busy


def synth_frames(u, wl, rank, world, pool):
    """This rank's frames of the stream: round-robin sharding, a pool of distinct frames tiled to the batch."""
    W, H, D, nb = wl["W"], wl["H"], wl["D"], wl["batch"]
    pool = min(nb, pool)
    mine = shard_frames(pool * world, rank, world)
    Lp, Rp = zip(*(u.synth_pair(wl["seed"], i, W, H, D) for i in mine))
    reps = (nb + pool - 1) // pool
    return np.concatenate([np.stack(Lp)] * reps)[:nb], np.concatenate([np.stack(Rp)] * reps)[:nb]


def strip(d):
    d = dict(d)
    d.pop("_pinned", None)
    return d


def sub_c1(cx, steps, reps):
    """BASELINE config 1: the reference's bundled pair (data/ref_rect_{l,r} -> ref_xsbl_{l,r}), block 15 and 21, both profiles.
    Parity is asserted in the run against reference-held vectors and committed fixtures (no oracle involved): the x-Sobel
    images equal ref_xsbl bit for bit, the RTL disparity matches the survey's CRC-32 cross-check values, the OPENCV
    disparity equals the committed cv2.StereoBM outputs."""
    u = cx.u
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_rect_xsbl.npz"))
    cvg = np.load(os.path.join(ROOT, "tests", "golden", "cv2_bm_golden.npz"))
    L, R = g["rect_l"], g["rect_r"]
    parity = {}
    with u.StereoFrontEnd(cx.local, 640, 480, 1) as fe:
        for B in (21, 15):
            fe.set_bm_params(width=640, height=480, profile=u.PROFILE_RTL, block_size=B, num_disparities=64, uni_enable=0, x_store_offset=1,
                             rtl_extended=0, min_disparity=0)
            fe.submit_rect(0, L, R); b = fe.wait()
            xl, xr = fe.receive_xsbl(b)
            parity[f"rtl_b{B}_xsobel_vs_ref_xsbl"] = bool(np.array_equal(xl[0], g["xsbl_l"]) and np.array_equal(xr[0], g["xsbl_r"]))
            crc = "%08x" % zlib.crc32(fe.receive_disp(b)[0].tobytes())
            parity[f"rtl_b{B}_disp_crc32"] = {"got": crc, "survey": C1_CRC[B], "ok": crc == C1_CRC[B]}
            fe.set_bm_params(profile=u.PROFILE_OPENCV, prefilter_cap=31, texture_threshold=10, uniqueness_ratio=10, disp12_max_diff=-1,
                             speckle_window_size=0, speckle_range=0)
            fe.submit_rect(1, L, R); b = fe.wait()
            parity[f"opencv_b{B}_vs_cv2_fixture"] = bool(np.array_equal(fe.receive_disp(b)[0], cvg[f"D64_B{B}_T10_U10"]))
        fe.set_bm_params(block_size=21, **MAINCPP)
        fe.submit_rect(0, L, R); b = fe.wait()
        parity["opencv_b21_postfilters_vs_cv2_fixture"] = bool(np.array_equal(fe.receive_disp(b)[0], cvg["maincpp_postfilter"]))
    ok = all((v["ok"] if isinstance(v, dict) else v) for v in parity.values())
    assert ok, f"C1 parity failed: {parity}"
    nb = 296
    hL = np.ascontiguousarray(np.broadcast_to(L, (nb, 480, 640))); hR = np.ascontiguousarray(np.broadcast_to(R, (nb, 480, 640)))
    rec = {"workload": "C1 bundled ref_rect pair x296, rect -> xsbl -> bm", "parity": parity, "parity_ok": ok, "runs": {}}
    for B in (15, 21):
        for prof, name in ((u.PROFILE_RTL, "rtl"), (u.PROFILE_OPENCV, "opencv")):
            wl = dict(W=640, H=480, D=64, B=B, seed=0, batch=nb)
            m = measure(cx, wl, hL, hR, profile=prof, entry="rect", steps=steps, warmup=2, reps=reps, e2e=(B == 15 and prof == u.PROFILE_RTL), e2e_reps=2)
            rec["runs"][f"{name}_b{B}"] = {"value": m["value"], "ms_per_step": m["ms_per_step"], "kernel_ms": m["stage_ms_per_step"]["bm"],
                                           "roofline_frac": m.get("bm_roofline", {}).get("frac"), "spread": m["reps"]["spread"],
                                           **({"e2e": m["e2e"]["value"]} if "e2e" in m else {})}
    return rec


def sub_sweep(cx, key, blocks, steps, reps, pool):
    """BASELINE configs 3 / 4: raw -> rect -> x-Sobel -> BM (RTL profile) on the synthetic stream of that shape."""
    u = cx.u
    base = dict(WORKLOADS[key])
    W, H = base["W"], base["H"]
    hL, hR = synth_frames(u, base, cx.rank, cx.world, pool)
    rp = u.identity_rect_params(W, H, float(W))
    rec = {"workload": base["name"].replace(f"B{base['B']}", "B" + "/".join(str(b) for b in blocks)), "frames_per_step_per_gpu": base["batch"], "runs": {}}
    for B in blocks:
        wl = dict(base, B=B)
        last = (B == blocks[-1])
        m = measure(cx, wl, hL, hR, profile=u.PROFILE_RTL, entry="raw", steps=steps, warmup=2, reps=reps, rect_params=rp, e2e=last, e2e_reps=2)
        r = {"value": m["value"], "mpix_disp_per_s": m["value"] * W * H * base["D"] / 1e6, "ms_per_step": m["ms_per_step"],
             "kernel_ms": m["stage_ms_per_step"]["bm"], "roofline_frac": m.get("bm_roofline", {}).get("frac"),
             "rect_ms": m["stage_ms_per_step"]["rect"], "xsbl_ms": m["stage_ms_per_step"]["xsbl"], "spread": m["reps"]["spread"]}
        if "e2e" in m:
            r["e2e"] = m["e2e"]["value"]
            r["e2e_bytes_per_step"] = m["e2e"]["h2d_bytes_per_step"] + m["e2e"]["d2h_bytes_per_step"]
        rec["runs"][f"b{B}"] = r
    return rec


def sub_opencv(cx, steps, reps, cv2_threads):
    """The reference's CPU mode like for like (slam/src/core/main.cpp:197-217): cv::StereoBM profile, block 21, 64 disparities,
    preFilterCap 31, texture 10, uniqueness 10, disp12MaxDiff 1, speckle 50/32 -- rectified pairs in HOST memory in, 16x disparity
    in HOST memory out (u96_submit_rect_async), next to cv2.StereoBM with the same settings on the same frames in this run."""
    u = cx.u
    wl = dict(W=640, H=480, D=64, B=21, seed=1, batch=296)
    pool = 74
    mine = shard_frames(pool * cx.world, cx.rank, cx.world)
    pairs = [u.synth_pair(1, i, 640, 480, 64) for i in mine]
    hL = np.concatenate([np.stack([p[0] for p in pairs])] * 4); hR = np.concatenate([np.stack([p[1] for p in pairs])] * 4)      # 296 = 2 full waves of the BM grid
    m = measure(cx, wl, hL, hR, profile=u.PROFILE_OPENCV, entry="rect", steps=steps, warmup=2, reps=reps, bm_extra=MAINCPP, e2e=True, e2e_reps=3)
    rec = {"workload": "cv::StereoBM profile + validateDisparity + filterSpeckles (main.cpp:198-212), 640x480 D64 B21, 296 rectified pairs per step "
                       "from a 74-frame pool, host buffers in and out",
           "value": m["value"], "ms_per_step": m["ms_per_step"], "e2e": m["e2e"]["value"], "e2e_spread": m["e2e"]["spread"],
           "kernel_ms": m["stage_ms_per_step"]["bm"], "postfilter_ms": m["stage_ms_per_step"]["post"], "xsbl_ms": m["stage_ms_per_step"]["xsbl"],
           "roofline_frac": m.get("bm_roofline", {}).get("frac"), "gpu_launches": m["gpu_launches"]}
    if cx.rank == 0 and cv2_threads:
        cv = cpu_cv2_reference(wl, 48, cv2_threads, pairs=pairs[:16])
        if cv is not None:
            rec["cv2_frames_per_s"] = cv
            rec["cv2_threads"] = cv2_threads
            rec["e2e_over_cv2"] = m["e2e"]["value"] / cv
            rec["resident_over_cv2"] = m["value"] / cv
    return rec


def aux_kernels(cx):
    """Kernel times of the stages that are not part of a submit: dense reprojection, keypoint reprojection, UVC packing, GFTT
    (640x480, 64 pairs), as fractions of the HBM copy peak for their algorithmic bytes."""
    u = cx.u
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    n, W, H = 64, 640, 480
    L, R = u.synth_batch(1, 0, 8, W, H, 64)
    hL = np.concatenate([L] * 8); hR = np.concatenate([R] * 8)
    sx, sy = 640 / 1241, 480 / 376
    P_l = np.array([[718.856 * sx, 0, 607.1928 * sx, 0], [0, 718.856 * sy, 185.2157 * sy, 0], [0, 0, 1, 0]])
    P_r = P_l.copy(); P_r[0, 3] = -386.1448 * sx
    out = {}
    with u.StereoFrontEnd(cx.local, W, H, n) as fe:
        fe.set_bm_registers((H << 16) + W, 0x00150040, 0)
        fe.set_profiling(True)
        fe.set_gftt(True)
        for _ in range(3):
            fe.submit_rect(0, hL, hR); b = fe.wait()
        st = fe.last_stage_ms_ex(b)
        px = n * W * H
        out["gftt"] = {"ms": st["gftt"], "frames": n, "hbm_frac": 3.0 * px / (st["gftt"] * 1e-3) / 1e9 / peak, "bytes_per_px": 3}
        for _ in range(2):
            fe.reproject_ex(b, P_l, P_r, 1, u.LOCAL_TRANSFORM, None)
        ms = fe.last_aux_ms(0)
        out["reproject_dense"] = {"ms": ms, "frames": n, "hbm_frac": 14.0 * px / (ms * 1e-3) / 1e9 / peak, "bytes_per_px": 14}
        for _ in range(2):
            fe.reproject_ex(b, P_l, P_r, 4, u.LOCAL_TRANSFORM, None)
        ms = fe.last_aux_ms(0)
        out["reproject_x4_decimated"] = {"ms": ms, "frames": n, "points": px // 16}
        rng = np.random.default_rng(0)
        uv = np.stack([rng.random(2000) * W, rng.random(2000) * H], 1).astype(np.float32)
        for _ in range(3):
            fe.reproject_points(b, P_l, P_r, uv, 0)
        out["reproject_points_2000"] = {"ms": fe.last_aux_ms(1)}
        for _ in range(2):
            fe.receive_uvc(b, u.UVC_BM)
        ms = fe.last_aux_ms(2)
        out["uvc_bm"] = {"ms": ms, "frames": n, "hbm_frac": 6.0 * px / (ms * 1e-3) / 1e9 / peak, "bytes_per_px": 6}
    return out


def single_pair_latency(cx, W, H, D, B, rp):
    """One iteration of the slam loop through the reference-shaped calls (main.cpp:165-181)."""
    u = cx.u
    L, R = u.synth_pair(1, 0, W, H, D)
    l1, r1 = L[None].copy(), R[None].copy()
    d1 = np.empty((1, H, W), np.int16)
    with u.StereoFrontEnd(cx.local, W, H, 1) as fe:
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, min_disparity=0, uni_enable=0,
                         x_store_offset=1, rtl_extended=int(D > 128))
        fe.set_rect_params(rp)
        ts = []
        for i in range(60):
            t0 = time.perf_counter()
            fe.submit_raw(i & 1, l1, r1)
            b = fe.wait()
            fe.receive_disp(b, out=d1)
            ts.append(time.perf_counter() - t0)
        fe.set_profiling(True)
        fe.submit_raw(0, l1, r1); b = fe.wait()
        st = fe.last_stage_ms_ex(b)
    return {"ms_median": 1e3 * float(np.median(ts[10:])), "ms_p90": 1e3 * float(np.quantile(ts[10:], 0.9)),
            "kernel_ms": {k: st[k] for k in ("rect", "xsbl", "bm")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5, help="repetitions of the K-step block; value = median")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1 / c3 / c4 / opencv sub-records")
    ap.add_argument("--wc-inputs", action="store_true", help="end-to-end loop reads its inputs from write-combined pinned memory")
    ap.add_argument("--gather", action="store_true", help="N>1: also time the optional gather of all disparity maps onto rank 0")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)
    args.reps = max(args.reps, 1)

    import torch
    import torch.distributed as dist
    # stdout carries exactly ONE line, the JSON: anything a library writes to fd 1 meanwhile (NCCL prints its version banner
    # there) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    cx = Ctx(args)
    u, rank, world, local = cx.u, cx.rank, cx.world, cx.local
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W, H, D, B, nb = wl["W"], wl["H"], wl["D"], wl["B"], wl["batch"]
    rp = u.SHIPPED_RECT_PARAMS if (W, H) == (640, 480) else u.identity_rect_params(W, H, float(W))
    hL, hR = synth_frames(u, wl, rank, world, 16)
    in_bytes = 2 * nb * W * H

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    head = measure(cx, wl, hL, hR, profile=u.PROFILE_RTL, entry="raw", steps=args.steps, warmup=args.warmup, reps=args.reps,
                   rect_params=rp, e2e=True, e2e_reps=max(3, min(args.reps, 5)))
    clocks = sampler.stop() if rank == 0 else None
    pL, pR, pD = head.pop("_pinned")
    pcie = pcie_all_ranks(cx, pL, pR, pD)
    ceiling = pcie * 1e9 / (2.0 * W * H) * world
    head["e2e"]["pcie_all_ranks"] = {"gbs_per_direction_per_rank": pcie, "gbs_per_direction_total": pcie * world, "ranks": world,
                                     "ceiling_frames_per_s": ceiling, "e2e_fraction_of_ceiling": head["e2e"]["value"] / ceiling,
                                     "note": "pinned H2D of the step's inputs and D2H of its disparity maps issued together by every rank at once, "
                                             "no kernels, slowest rank closes the step"}
    head["e2e"]["pcie"] = {"bidir_gbs_per_direction": pcie, "ceiling_frames_per_s": ceiling}
    head["e2e"]["host_numa_node_rank0"] = numa_node
    del pL, pR, pD

    gather = None
    if args.gather and world > 1:
        # optional, off the hot path: one step's disparity maps of every rank onto rank 0 over NVLink (zero-copy send buffer)
        fe = u.StereoFrontEnd(local, W, H, nb)
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, x_store_offset=1, rtl_extended=int(D > 128))
        fe.set_rect_params(rp)
        fe.submit_raw(0, hL, hR); fe.wait()
        t_d = fe.disp_tensor(0)
        for _ in range(3):                             # the first send/recv pairs connect the P2P channels lazily
            gather_disparity(t_d, rank, world)
        cx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            got = gather_disparity(t_d, rank, world)
        g1.record()
        torch.cuda.synchronize()
        tg = cx.max_over_ranks(g0.elapsed_time(g1) / 5)
        if rank == 0:
            nbytes = t_d.numel() * 2
            gather = {"ms": tg, "bytes_per_rank": nbytes, "gbs_into_rank0": (world - 1) * nbytes / tg / 1e6,
                      "ok": bool(len(got) == world and torch.equal(got[0], t_d))}
        del t_d
        fe.close()

    # ------------------------------------------------------------------ the other BASELINE configurations
    configs, cfg_err = {}, {}
    if not args.no_configs and args.workload == "c2":
        sub_steps, sub_reps = max(3, min(args.steps, 6)), 3
        cores = os.cpu_count() or 1
        plan = [("c1", lambda: sub_c1(cx, sub_steps, sub_reps)),
                ("c3", lambda: sub_sweep(cx, "c3", (9, 15, 21), sub_steps, sub_reps, 8)),
                ("c4", lambda: sub_sweep(cx, "c4", (21,), sub_steps, sub_reps, 8)),
                ("opencv", lambda: sub_opencv(cx, sub_steps, sub_reps, 0 if args.no_cpu_baseline else cores))]
        for name, fn in plan:
            try:
                configs[name] = fn()
            except Exception as e:                      # a failing sub-record must not take the headline with it -- but it is reported
                cfg_err[name] = f"{type(e).__name__}: {e}"
                cx.barrier()

    if rank == 0:
        aux, lat = None, None
        try:
            aux = aux_kernels(cx)
            lat = {"what": "one host pair: u96_submit_raw + u96_wait + u96_receive_disp (pageable host buffers), 640x480 D64",
                   "b21": single_pair_latency(cx, 640, 480, 64, 21, u.SHIPPED_RECT_PARAMS),
                   "b15": single_pair_latency(cx, 640, 480, 64, 15, u.SHIPPED_RECT_PARAMS)}
        except Exception as e:
            cfg_err["aux"] = f"{type(e).__name__}: {e}"
        # ---- roofline of the dominant kernel (k_bm_fused<RTL, saturating, 64 disparities>): integer pipe by the SURVEY 8(d) convention, measured issue rate as the peak ----
        stage = head["stage_ms_per_step"]
        bmr = head["bm_roofline"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_note = None, "no capture for this build"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "bm_traffic.json")))
            from u96_slam_b200 import build as _build
            if tr.get("workload") == args.workload and tr.get("src_sha16") == _build.source_sha16():
                traffic = tr.get("dram_bytes_per_frame") * nb
                traffic_note = f"ncu --set full capture of these very kernel sources ({tr.get('source')})"
            elif tr.get("workload") == args.workload:
                traffic_note = f"profiles/bm_traffic.json was captured with other kernel sources ({tr.get('src_sha16')}): not reported"
        except (OSError, TypeError):
            pass
        bm_ms = stage["bm"]
        roofline = {"kernel": "k_bm_fused<RTL,SAT,8>", "bound": "int", "achieved": bmr["achieved"], "peak": bmr["peak"], "unit": "Tlaneop/s",
                    "frac": bmr["frac"],
                    "peak_source": "measured live: u96_microbench VABSDIFF4 issue rate (ALU pipe, 64 lanes/clk/SM); "
                                   "MEASURED_PEAKS.json has no integer figure",
                    "algorithmic_ops_per_launch": bmr["algorithmic_ops_per_launch"], "kernel_ms": bm_ms, "traffic": traffic, "traffic_note": traffic_note,
                    "binding_resource": "no unit saturated (ncu: shared-memory data pipe 75 % of peak, issue slots 65 %, ALU pipe 54 %, FMA pipe 25 %): "
                                        "two CTA barriers per image row and dependent shared-memory loads leave 5 compute warps per scheduler "
                                        "latency bound (profiles/r02_summary.md)",
                    "hbm": {"achieved": 4.0 * W * H * nb / (bm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": 4.0 * W * H * nb / (bm_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                    "stage_ms_per_step": stage,
                    "other_kernels": {"rect": {"ms": stage["rect"], "hbm_frac": 4.0 * W * H * nb / (stage["rect"] * 1e-3) / 1e9 / hbm_peak,
                                               "binding_resource": "issue slots / ALU pipe (ncu: 75 % / 57 %), not HBM"},
                                      "xsobel": {"ms": stage["xsbl"], "hbm_frac": 4.0 * W * H * nb / (stage["xsbl"] * 1e-3) / 1e9 / hbm_peak},
                                      **(aux or {})}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_s = max(8, min(4 * cores, int(20.0 * cores / (0.25 * W * H * D / 19.66e6))))
            v = cpu_oracle_baseline(wl, n_s, cores)
            cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"C oracle port of the same pipeline (rectify+xsobel+bm_rtl) on {n_s} frames, {cores} threads"}
            cv = cpu_cv2_reference(wl, 32, cores)
            if cv is not None:
                cpu["cv2_stereobm_frames_per_s"] = cv
                cpu["cv2_stereobm_1thread_frames_per_s"] = cpu_cv2_reference(wl, 12, 1)      # 1 thread and nproc
            cpu["port_1thread_frames_per_s"] = cpu_oracle_baseline(wl, 2, 1)
            cpu["cpu_model"] = next((ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")), "unknown")
        line = {"metric": "disparity_frames_per_s", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "mpix_disp_per_s": head["value"] * W * H * D / 1e6,
                "reps": head["reps"],
                "config": {"workload": wl["name"], "frames_per_step_per_gpu": nb, "sharding": "frame-wise round-robin, no collective",
                           "l2": f"inputs {in_bytes / 1e6:.0f} MB per step > 126 MB L2" if in_bytes > 126e6 else
                                 f"inputs {in_bytes / 1e6:.0f} MB per step (<L2; intermediates {7 * in_bytes / 2e6:.0f} MB)"},
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "single_pair_latency": lat, "configs": {k: strip(v) for k, v in configs.items()}}
        if "opencv" in configs and "e2e_over_cv2" in configs["opencv"]:
            o = configs["opencv"]
            line["like_for_like"] = {"what": "reference CPU mode (cv::StereoBM + validateDisparity + filterSpeckles, main.cpp:197-217): GPU end to end "
                                             "with host buffers vs cv2 on all host threads, same frames, same run",
                                     "gpu_e2e_frames_per_s": o["e2e"], "cv2_frames_per_s": o["cv2_frames_per_s"], "ratio": o["e2e_over_cv2"]}
        if cfg_err:
            line["config_errors"] = cfg_err
        if gather is not None:
            line["gather"] = gather
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
