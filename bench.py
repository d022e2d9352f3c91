#!/usr/bin/env python
"""bench.py -- throughput of the dense-stereo front end on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4]

A "step" is one pass of the hot path (rect -> x-Sobel -> SAD block matching -> 16x disparity) over
one batch of synthetic stereo pairs.  Default workload = BASELINE.json configs[1] (C2): 640x480,
64 disparities, block 21, RTL profile, full remap+Sobel+BM pipeline.

  value  : frames/s with the batch already resident in HBM (CUDA events on the launching stream)
  e2e    : frames/s through the C ABI with pinned HOST buffers (H2D + kernels + D2H every step)
  roofline: the BM kernel against the measured integer-pipe issue rate (and its HBM fraction)
  cpu_baseline: the CPU oracle port of the same pipeline on this box's host cores (bounded sample)

Frames are sharded frame-wise over ranks (one process per GPU, torch.distributed only for the
barrier and the max-over-ranks time): weak scaling, no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, D, block, seed, frames per step per GPU)
    "c2": dict(W=640, H=480, D=64, B=21, seed=1, batch=296, name="C2 synthetic 640x480 D64 B21 raw->rect->xsbl->bm (RTL profile)"),
    "c3": dict(W=1242, H=375, D=128, B=15, seed=2, batch=148, name="C3 KITTI-shape 1242x375 D128 B15 raw->rect->xsbl->bm (RTL profile)"),
    "c4": dict(W=1920, H=1080, D=256, B=21, seed=3, batch=32, name="C4 1920x1080 D256 B21 raw->rect->xsbl->bm (RTL-extended profile)"),
}


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


def cpu_cv2_reference(wl, frames, threads):
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
            "config": {"workload": wl["name"], "frames_per_step": per_step},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true", help="N>1: also time the optional gather of all disparity maps onto rank 0")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import u96_slam_b200 as u
    from u96_slam_b200.stereo import microbench, BUF_DISP

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: anything a library writes to fd 1 meanwhile (NCCL prints its version banner
    # there) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W, H, D, B, nb = wl["W"], wl["H"], wl["D"], wl["B"], wl["batch"]
    rp = u.SHIPPED_RECT_PARAMS if (W, H) == (640, 480) else u.identity_rect_params(W, H, float(W))
    # this rank's frames of the stream: round-robin sharding, a pool of distinct frames tiled to the batch
    pool = min(nb, 16)
    mine = shard_frames(pool * world, rank, world)
    Lp, Rp = zip(*(u.synth_pair(wl["seed"], i, W, H, D) for i in mine))
    reps = (nb + pool - 1) // pool
    hL = np.concatenate([np.stack(Lp)] * reps)[:nb]; hR = np.concatenate([np.stack(Rp)] * reps)[:nb]

    def configure(fe):
        fe.set_bm_params(width=W, height=H, profile=u.PROFILE_RTL, block_size=B, num_disparities=D, min_disparity=0,
                         uni_enable=0, uni_mode=0, uni_thr=0, x_store_offset=1, rtl_extended=int(D > 128))
        fe.set_rect_params(rp)

    # ------------------------------------------------------------------ resident (value)
    fe = u.StereoFrontEnd(local, W, H, nb)
    configure(fe)
    stream = torch.cuda.current_stream()
    fe.set_stream(stream.cuda_stream)
    fe.set_profiling(True)
    dL = torch.from_numpy(hL).cuda(); dR = torch.from_numpy(hR).cuda()
    in_bytes = dL.numel() + dR.numel()

    def step(i):
        fe.submit_device("raw", i & 1, dL.data_ptr(), dR.data_ptr(), W, nb)
        fe.wait()

    def run_steps(k, stage=None):
        # the two banks are used the way the reference's producer uses them (bank = iteration % 2, main.cpp:168): step i+1 is
        # submitted before step i is waited for, so the host round trip of u96_wait never leaves the GPU idle between steps
        for i in range(k):
            b = i & 1
            if i >= 2:
                assert fe.wait() == b
                if stage is not None:
                    for kk, v in fe.last_stage_ms(b).items():
                        stage[kk] += v
            fe.submit_device("raw", b, dL.data_ptr(), dR.data_ptr(), W, nb)
        for i in range(max(k - 2, 0), k):
            b = fe.wait()
            if stage is not None:
                for kk, v in fe.last_stage_ms(b).items():
                    stage[kk] += v

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"h2d": 0.0, "rect": 0.0, "xsbl": 0.0, "bm": 0.0}
    torch.cuda.synchronize()
    e0.record(stream)
    run_steps(args.steps, stage)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = fe.kernel_launches() - l0
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    frames_total = nb * args.steps * world
    fps = frames_total / (ms_total_max * 1e-3)
    for k in stage:
        stage[k] /= args.steps
    gather = None
    if args.gather and world > 1:
        # optional, off the hot path: the step's disparity maps of every rank onto rank 0 over NVLink (zero-copy send buffer)
        t_d = fe.disp_tensor((args.steps - 1) & 1)
        for _ in range(3):                             # the first send/recv pairs connect the P2P channels lazily
            gather_disparity(t_d, rank, world)
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(5):
            got = gather_disparity(t_d, rank, world)
        g1.record(stream)
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1) / 5], device="cuda", dtype=torch.float64)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = len(got) == world and torch.equal(got[0], t_d)           # rank 0's own slot round-trips
            nbytes = t_d.numel() * 2
            gather = {"ms": float(tg.item()), "bytes_per_rank": nbytes, "gbs_into_rank0": (world - 1) * nbytes / float(tg.item()) / 1e6,
                      "frames_per_s_incl_gather": frames_total / ((ms_total_max + args.steps * float(tg.item())) * 1e-3), "ok": bool(ok)}
        del t_d
    fe.close()
    del dL, dR

    # ------------------------------------------------------------------ end to end (host buffers through the C ABI)
    fe2 = u.StereoFrontEnd(local, W, H, nb)
    configure(fe2)
    pL = torch.from_numpy(hL).pin_memory(); pR = torch.from_numpy(hR).pin_memory()
    pD = [torch.empty((nb, H, W), dtype=torch.int16).pin_memory() for _ in range(2)]

    def e2e_loop(k):
        # two banks in flight: H2D + kernels + D2H of bank b are queued back to back on its stream, the host only
        # waits for the older bank, so both copy engines and the SMs overlap
        for i in range(k):
            b = i & 1
            if i >= 2:
                assert fe2.wait() == b
            fe2.submit_host_ptr_async("raw", b, pL.data_ptr(), pR.data_ptr(), W, nb, pD[b].data_ptr())
        for _ in range(min(k, 2)):
            fe2.wait()

    e2e_steps = max(4, args.steps)
    e2e_loop(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_fps = nb * e2e_steps * world / float(t_e2e.item())
    checksum = int(pD[0][0].to(torch.int64).sum().item())
    fe2.close()
    # PCIe ceiling of that loop: the same pinned buffers copied both ways at once, no kernels (explains e2e vs value)
    pcie = None
    if rank == 0:
        dI = torch.empty_like(pL, device="cuda"); dO = torch.empty_like(pD[0], device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(2):
                with torch.cuda.stream(s_in):
                    dI.copy_(pL, non_blocking=True); dI.copy_(pR, non_blocking=True)
                with torch.cuda.stream(s_out):
                    pD[1].copy_(dO, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, 2 * 2 * pL.numel() / (time.perf_counter() - t0) / 1e9)
        pcie = {"bidir_gbs_per_direction": best, "ceiling_frames_per_s": best * 1e9 / (2.0 * W * H),
                "note": "pinned H2D of the step's inputs and D2H of its disparity maps issued together, no kernels"}
        del dI, dO

    # single-pair latency through the reference-shaped calls (what one iteration of the slam loop pays: main.cpp:165-181)
    latency = None
    if rank == 0:
        fe3 = u.StereoFrontEnd(local, W, H, 1)
        configure(fe3)
        l1, r1 = hL[:1].copy(), hR[:1].copy()
        d1 = np.empty((1, H, W), np.int16)
        ts = []
        for i in range(40):
            t0 = time.perf_counter()
            fe3.submit_raw(i & 1, l1, r1)
            b = fe3.wait()
            fe3.receive_disp(b, out=d1)
            ts.append(time.perf_counter() - t0)
        fe3.close()
        latency = {"ms_median": 1e3 * float(np.median(ts[8:])), "ms_p90": 1e3 * float(np.quantile(ts[8:], 0.9)),
                   "what": "one host pair: u96_submit_raw + u96_wait + u96_receive_disp (pageable host buffers)"}

    if rank == 0:
        # ---- roofline of the dominant kernel (k_bm): integer pipe, measured issue rate as the peak ----
        int_peak = microbench(1, local) / 1e3            # VABSDIFF4 / ALU-pipe issue rate, T lane-op/s
        bm_ms = stage["bm"]
        algo_ops = 6.0 * W * H * D * nb                  # 6 integer lane-ops per pixel-disparity
        achieved = algo_ops / (bm_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "bm_traffic.json")))
            if tr.get("workload") == args.workload:
                traffic = tr.get("dram_bytes_per_frame") * nb
        except (OSError, TypeError):
            pass
        roofline = {"kernel": "k_bm", "bound": "int", "achieved": achieved, "peak": int_peak, "unit": "Tlaneop/s",
                    "frac": achieved / int_peak,
                    "peak_source": "measured live: u96_microbench VABSDIFF4 issue rate (ALU pipe, 64 lanes/clk/SM); "
                                   "MEASURED_PEAKS.json has no integer figure",
                    "algorithmic_ops_per_launch": algo_ops, "kernel_ms": bm_ms, "traffic": traffic,
                    "hbm": {"achieved": 4.0 * W * H * nb / (bm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": 4.0 * W * H * nb / (bm_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                    "stage_ms_per_step": stage}
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
        line = {"metric": "disparity_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "mpix_disp_per_s": fps * W * H * D / 1e6,
                "config": {"workload": wl["name"], "frames_per_step_per_gpu": nb, "sharding": "frame-wise round-robin, no collective",
                           "l2": f"inputs {in_bytes / 1e6:.0f} MB per step > 126 MB L2" if in_bytes > 126e6 else
                                 f"inputs {in_bytes / 1e6:.0f} MB per step (<L2; intermediates {7 * in_bytes / 2e6:.0f} MB)"},
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(2 * W * H * nb),
                        "d2h_bytes_per_step": int(2 * W * H * nb), "steps": e2e_steps, "checksum": checksum,
                        "timing": "wall clock around u96_submit_raw_async/u96_wait over two banks, synchronize on both sides",
                        "pcie": pcie, "host_numa_node_rank0": numa_node},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "single_pair_latency": latency}
        if gather is not None:
            line["gather"] = gather
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
