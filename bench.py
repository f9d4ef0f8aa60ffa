#!/usr/bin/env python
"""bench.py — frame-pairs/sec of the line front end (extract + match + RANSAC pose), BASELINE.json's metric.

Workload (configs[1]): the synthetic fr1/xyz-shape 640x480 RGB-D stream, line-only odometry. One "step" is
one batch of B consecutive frames of the stream: every frame is extracted once (LSD -> 3D lines -> MSLD ->
MLE), matched against its predecessor (the predecessor of the first frame is the cached last frame of the
previous step) and registered (500-iteration RANSAC + LM refinement) -> B pose records.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]      the CUDA path through the C ABI
  python bench.py --impl reference ...                                 the reference's CPU path (oracle/)
  python bench.py --workload cfg3|cfg4|cfg5 ...                        the other BASELINE configs (lines kept under profiles/)

`value`  : device-resident inputs (already in HBM when the timed region starts).
`e2e`    : same steps through the host-buffer entry point: pinned host RGB u8 + the 16-bit depth the sensor / TUM PNG
           delivers (lsl_extract_batch_u16; metres + NaN conversion on the device) -> H2D inside the timed region,
           pose records D2H (what Node::Node + Node::matchNodePair callers see).
Multi-GPU: one process per GPU (torchrun), the stream is sharded (rank r owns its own batches, weak scaling),
no data-path collective; the pose records are all-gathered with NCCL at the end of every step (graph-insert time).
Only the cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 640, 480
B_FRAME = W * H * (3 + 4)          # compulsory HBM bytes per extracted frame: RGB u8 + depth f32
METRIC = "frame-pairs/sec (extract+match+RANSAC pose) on 640x480 RGB-D"

def _hbm_peak(peaks, fallback: float = 6650.0) -> float:
    """HBM GB/s out of the driver-written MEASURED_PEAKS.json (key `hbm_gbs`; any numeric entry whose key names HBM /
    copy bandwidth is accepted, nested or not); the profiling recipe's fallback otherwise."""
    def walk(d, path=""):
        if isinstance(d, dict):
            for k, v in d.items():
                kl = path + "/" + str(k).lower()
                if isinstance(v, (int, float)) and ("hbm" in kl or "copy" in kl) and v > 100:
                    # exact key first, then sustained figures (the kernel is timed inside a long step), then the rest
                    yield (0 if kl == "/hbm_gbs" else 1 if "sustain" in kl else 2, float(v))
                else:
                    yield from walk(v, kl)
        elif isinstance(d, list):
            for v in d:
                yield from walk(v, path)
    found = sorted(walk(peaks))
    return found[0][1] if found else fallback


def palindrome(u: int, n: int, phase: int = 0):
    """Frame order 0..u-1,u-2..1,0,1.. so that consecutive frames are always neighbours of the real stream."""
    period = list(range(u)) + list(range(u - 2, 0, -1)) if u > 1 else [0]
    return [period[(phase + k) % len(period)] for k in range(n)]


def _render_one(args):
    """Pool worker: one frame of a synthetic stream (+ SIFT point features for cfg 3)."""
    from lineslam_b200 import synth
    seed, idx, Wd, Hd, traj_name, want_sift = args
    global _SCENE
    if "_SCENE" not in globals() or _SCENE[0] != seed:
        _SCENE = (seed, synth.Scene(seed))
    traj = synth.trajectory_orbit if traj_name == "orbit" else synth.trajectory_xyz
    img, dep, R, p = synth.make_frame(_SCENE[1], idx, seed, Wd, Hd, traj)
    pts = None
    if want_sift:   # cfg 3: cv2 SIFT on the rendered frame -> xyz from the depth map; raw SIFT rows, RootSIFT on device
        import cv2
        K = synth.camera_K(Wd, Hd)
        kp, desc = cv2.SIFT_create(nfeatures=600).detectAndCompute(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), None)
        kp, desc = kp[:600], (desc[:600] if desc is not None else np.zeros((0, 128), np.float32))
        uv = np.array([k.pt for k in kp], np.float64).reshape(-1, 2)
        ui = np.clip(np.rint(uv[:, 0]).astype(int), 0, Wd - 1); vi = np.clip(np.rint(uv[:, 1]).astype(int), 0, Hd - 1)
        z = dep[vi, ui].astype(np.float64)
        xyz1 = np.ones((len(kp), 4), np.float32)
        xyz1[:, 0] = ((uv[:, 0] - K[0, 2]) / K[0, 0] * z).astype(np.float32)
        xyz1[:, 1] = ((uv[:, 1] - K[1, 2]) / K[1, 1] * z).astype(np.float32)
        xyz1[:, 2] = z.astype(np.float32)
        pts = (xyz1, np.ascontiguousarray(desc, np.float32))
    return img, dep, pts


def make_unique_frames(u: int, rank: int, world: int = 1, Wd: int = W, Hd: int = H, traj: str = "xyz", sift: bool = False,
                       stride: int = 1):
    """u distinct consecutive frames of the synthetic stream, rendered in a process pool (0.45 s per VGA frame per core).
    Must run BEFORE CUDA is initialised in this process (the pool forks)."""
    import multiprocessing as mp
    from lineslam_b200 import synth
    workers = max(1, min((os.cpu_count() or 1) // max(world, 1), u))
    start = rank * 7
    jobs = [(2000, start + k * stride, Wd, Hd, traj, sift) for k in range(u)]
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_render_one, jobs, chunksize=max(1, u // (workers * 4)))
    else:
        res = [_render_one(j) for j in jobs]
    imgs = np.stack([r[0] for r in res]); deps = np.stack([r[1] for r in res])
    pts = [r[2] for r in res] if sift else None
    K = synth.camera_K(Wd, Hd)
    return (imgs, deps, K, pts) if sift else (imgs, deps, K)


def depth_to_u16(deps: np.ndarray) -> np.ndarray:
    """The 16-bit TUM depth the synthetic f32 planes came from (they are quantised to 1/5000 m; NaN = 0)."""
    raw = np.nan_to_num(deps.astype(np.float64) * 5000.0, nan=0.0)
    return np.rint(raw).astype(np.uint16)


class ClockSampler:
    """SM clock and throttle reasons every 500 ms while the timed regions run (B200_PROFILING.md recipe): ONE
    long-lived `nvidia-smi -lms 500` child started before the warm-up (its start-up attaches to the driver once,
    outside the timed regions) and killed afterwards. Rows are filtered to the timed regions by wall clock.
    (An in-process NVML thread was measured to stall the CUDA driver for up to 1.7 s at random.)"""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        import tempfile
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = vis.split(",")[index] if vis and index < len(vis.split(",")) else str(index)
        self.path = tempfile.mktemp(prefix="lsl_clocks_", suffix=".csv")
        self.windows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", phys, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "500"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def start(self):
        time.sleep(1.5 if self.proc else 0.0)     # let the child finish its driver attach before any timing

    def window(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}
        time.sleep(0.6)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for line in open(self.path, errors="ignore"):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(c[2]), float(c[3])
            except Exception:
                continue
            if not any(a - 0.05 <= ts <= b + 0.05 for a, b in self.windows):
                continue
            sm.append(clk); mx = cmax
            for name, v in zip(self.NAMES, c[4:8]):
                if v.lower() == "active":
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 500"}


# ------------------------------------------------------------------------------------------ CPU arm ----
def cpu_pairs_per_sec(imgs, deps, K, n_pairs: int, threads: int):
    """The reference's CPU path (oracle restatement, TEST INFRASTRUCTURE used here only as the timed baseline):
    `threads` independent single-threaded streams side by side — the best the host cores can do on this
    embarrassingly parallel workload. Returns (pairs/s, seconds, pairs)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    pyoracle.lib()
    u = len(imgs)
    per = max(1, n_pairs // threads)

    def stream(t):
        order = palindrome(u, per + 1, phase=t)
        _, prev, _ = pyoracle.stream_step(imgs[order[0]], deps[order[0]], K, 1, None)
        t0 = time.perf_counter()
        found = 0
        for k in range(1, per + 1):
            _, cur, rec = pyoracle.stream_step(imgs[order[k]], deps[order[k]], K, 1 + k, prev)
            found += int(rec["found"])
            prev = cur
        return time.perf_counter() - t0, found

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(stream, range(threads)))
    wall = max(r[0] for r in res)
    total = per * threads
    return total / wall, time.perf_counter() - t0, total, sum(r[1] for r in res)


def tum_depth_planes(deps):
    """(u16 raw, f32 metres) as the reference's loader makes them from the 16-bit PNG (openni_listener.cpp:1233-1244)."""
    raw = depth_to_u16(deps)
    conv = np.where(raw == 0, np.float32(np.nan), raw.astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
    return raw, conv


WORKLOADS = {
    "cfg2": "cfg2 fr1/xyz-shape synthetic stream 640x480, line-only odometry",
    "cfg3": "cfg3 fr2/desk-shape synthetic stream 640x480, line + SIFT-point fusion (Node::matchNodePair both modalities)",
    "cfg4": "cfg4 loop-closure batch: 1 query x 256 keyframes line matching + RANSAC, keyframes block-wise over the ranks",
    "cfg5": "cfg5 1280x960 stream, line odometry + levmar refine (computeRelativeMotion_Ransac + optimizeRelmotion) per edge",
}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    imgs, deps, K = make_unique_frames(min(args.unique, max(24, threads + 2)), 0)
    _, deps = tum_depth_planes(deps)
    per_step = threads * 2
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_pairs_per_sec(imgs, deps, K, threads, threads)
    tot_pairs, tot_time = 0, 0.0
    for _ in range(args.steps):
        pps, sec, pairs, _ = cpu_pairs_per_sec(imgs, deps, K, per_step, threads)
        tot_pairs += pairs
        tot_time += pairs / pps
    value = tot_pairs / tot_time
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["cfg2"] + " (CPU reference path)",
                   "pairs_per_step": per_step, "unique_frames": len(imgs)},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                         "per_core": value / threads,
                         "sample": f"{per_step} pairs per step: {threads} independent single-thread streams x 2 pairs "
                                   f"(oracle restatement of detect3DLines + lineMatching + getTransform_PtsLines_ransac)"},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm ----
# FP64-pipe / issue-slot utilisation of the dominant kernels from the committed ncu capture of this build
# (profiles/r2_ncu_summary.md): what actually bounds the path (it is not HBM).
NCU_PIPE = {}
try:
    NCU_PIPE = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_pipe.json")))
except Exception:
    pass


def run_cuda(args, rank, world, local_rank):
    wl = args.workload
    big = wl == "cfg5"
    Wd, Hd = (1280, 960) if big else (W, H)
    b_frame = Wd * Hd * (3 + 4)
    B = args.batch if args.batch > 0 else {"cfg2": 1184, "cfg3": 592, "cfg4": 256, "cfg5": 148}[wl]
    dev_sift = wl == "cfg3" and args.points == "device"
    U = min(args.unique, B) if wl != "cfg4" else 257
    # ---- inputs first: the render pool forks, CUDA must not be initialised yet
    t_r = time.perf_counter()
    pts = None
    split = wl == "cfg2" and world > 1 and args.split == "stream"
    if split:   # ONE stream dealt block-wise to the ranks: every block must start at the same phase of the palindromic tiling
        U = B // 4 + 1 if B % 4 == 0 and B >= 8 else U
        assert B % (2 * U - 2) == 0, "--split stream needs a batch that is a multiple of the tiling period 2 * unique - 2"
    if wl == "cfg2":
        imgs, deps, K = make_unique_frames(U, 0 if split else rank, world)
    elif wl == "cfg3":
        if dev_sift:   # point features detected on the device inside every extract call (k_sift.cu)
            imgs, deps, K = make_unique_frames(U, rank, world, traj="orbit")
        else:
            imgs, deps, K, pts = make_unique_frames(U, rank, world, traj="orbit", sift=True)
    elif wl == "cfg4":   # 1 query + 256 keyframes sampled from the cfg-3 orbit; every rank renders only its block (+ the query on rank 0)
        from lineslam_b200.shard import shard_pairs, pad_records
        from lineslam_b200 import synth
        lo, hi, per = shard_pairs(256, world, rank)
        K = synth.camera_K(W, H)
    else:
        imgs, deps, K = make_unique_frames(U, rank, world, Wd, Hd, traj="orbit")
    if wl == "cfg4":
        import multiprocessing as mp
        idxs = ([0] if rank == 0 else []) + [40 + 11 * k for k in range(lo, hi)]      # query = orbit frame 0, keyframes every 11th frame
        workers = max(1, min((os.cpu_count() or 1) // max(world, 1), len(idxs)))
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_render_one, [(2000, i, W, H, "orbit", False) for i in idxs])
        imgs = np.stack([r[0] for r in res]); deps = np.stack([r[1] for r in res])
    raw16, deps = tum_depth_planes(deps)
    render_s = time.perf_counter() - t_r

    import torch
    import torch.distributed as dist
    from lineslam_b200 import api
    from lineslam_b200 import shard as shard_mod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = api.default_params()
    ctx = api.Context(params=p, device=local_rank, max_batch=max(B if wl != "cfg4" else len(imgs), 1), max_w=Wd, max_h=Hd)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if dev_sift:
        ctx.set_point_detector("SIFT", p.max_keypoints if hasattr(p, "max_keypoints") else 600, root_sift=True)
    if world > 1:  # library-owned NCCL communicator; the id travels over torch.distributed (plumbing)
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], world, rank)

    state = {"prev": None, "step": 0, "found": 0, "pairs": 0, "lines": []}
    per_step = []
    ktime = {}

    def note_ktimes(pair_stage):
        for k, v in ctx.kernel_times().items():
            if v > 0 and (k.startswith(("match", "pose", "relmotion")) == pair_stage):
                ktime.setdefault(k, []).append(v)

    if wl == "cfg4":
        # ---- setup (untimed): every rank extracts its keyframe block once; rank 0 also the query
        frames = ctx.extract_batch(imgs, deps, K, seeds=np.arange(1, len(imgs) + 1))
        query0 = frames[0] if rank == 0 else None
        keyframes = frames[1:] if rank == 0 else frames
        n_loc = len(keyframes)
        kf_ids = np.arange(lo, hi, dtype=np.int32) + 100
        state["lines"] = [f.num_lines for f in frames]
        pairs_per_step_total = 256

        def one_step(e2e):
            q = ctx.bcast_frame(query0, 0) if world > 1 else query0          # ncclBroadcast of the query record
            recs = ctx.match_pair_batch([q] * n_loc, keyframes, np.full(n_loc, 1000, np.int32), kf_ids,
                                        np.arange(n_loc, dtype=np.uint32) + 1 + lo)
            note_ktimes(True)
            if world > 1:
                allr = ctx.allgather_poses(None, n_loc) if n_loc == per else ctx.allgather_poses(pad_records(recs, per))
                assert len(allr) == world * per
                if q is not query0:
                    q.free()
            state["found"] += int(recs["found"].sum()); state["pairs"] += n_loc
            state["step"] += 1
    else:
        nbuf = 2
        host_i, host_d16, dev_i, dev_d = [], [], [], []
        for pb in range(nbuf):
            order = palindrome(U, B, pb * B)
            hi_ = torch.from_numpy(np.stack([imgs[i] for i in order])).pin_memory()
            hd16 = torch.from_numpy(np.stack([raw16[i] for i in order]).view(np.int16)).pin_memory()
            host_i.append(hi_); host_d16.append(hd16)
            dev_i.append(hi_.cuda()); dev_d.append(torch.from_numpy(np.stack([deps[i] for i in order])).cuda())
        orders = [palindrome(U, B, pb * B) for pb in range(nbuf)]
        torch.cuda.synchronize()
        pairs_per_step_total = world * B

        def one_step(e2e):
            s_ = state["step"]
            b = s_ % nbuf
            blk = shard_mod.stream_block(s_, world, rank, B)[0] if split else s_   # block of the (shared or own) stream this step extracts
            seeds = np.arange(1, B + 1, dtype=np.uint32) + blk * B
            if e2e:
                frames = ctx.extract_batch(host_i[b].numpy(), host_d16[b].numpy().view(np.uint16), K, seeds)
            else:
                frames = ctx.extract_batch_dev(dev_i[b].data_ptr(), 3, dev_d[b].data_ptr(), B, Wd, Hd, K, seeds)
            note_ktimes(False)
            if pts is not None:      # cfg 3: the detectors' output enters as an input (Node::Node, src/node.cpp:219-310)
                ctx.set_points_batch(frames, [pts[i][0] for i in orders[b]], [pts[i][1] for i in orders[b]], root_sift=True)
            recv = None
            if split:   # the head pair of this block needs the tail record of the previous block: ring shift over NVLink
                recv = ctx.shift_frame(frames[-1])
                if s_ == 0:   # once, untimed warm-up: what arrived is bit for bit what the neighbour extracted
                    import hashlib
                    hs = [None] * world
                    dist.all_gather_object(hs, (hashlib.sha1(frames[-1].lines().tobytes()).hexdigest(), hashlib.sha1(recv.lines().tobytes()).hexdigest()))
                    assert all(hs[r][1] == hs[(r - 1) % world][0] for r in range(world)), "lsl_shift_frame: records differ from the sender's"
                    state["shift_checked"] = True
                head = recv if rank > 0 else state["prev"]
            else:
                head = state["prev"]
            trains = [head] + frames[:-1] if head is not None else [frames[0]] + frames[:-1]
            ids = np.arange(B, dtype=np.int32) + blk * B + 1
            if pipelined:   # the pair stage of the PREVIOUS batch ran under this batch's extraction: collect it, then start this one
                finish_pairs()
                ctx.match_pair_batch_begin(frames, trains, ids, ids - 1, seeds)
                state["pending"] = (frames, head if not split or rank == 0 else None, recv if split and rank > 0 else None)
                if split:
                    state["prev"] = recv if rank == 0 else None
                else:
                    state["prev"] = frames[-1]
            else:
                recs = ctx.match_pair_batch(frames, trains, ids, ids - 1, seeds)
                after_pairs(recs, frames)
                old = state["prev"]
                if split:   # rank 0 keeps what it received (the last rank's tail precedes the head of its next block)
                    state["prev"] = recv if rank == 0 else None
                    if rank > 0:
                        recv.free()
                    frames[-1].free()
                else:
                    state["prev"] = frames[-1]
                for f in frames[:-1]:
                    f.free()
                if old is not None:
                    old.free()
            state["step"] += 1

        def after_pairs(recs, frames):
            note_ktimes(True)
            if wl == "cfg5":         # levmar refine per edge (computeRelativeMotion_Ransac + optimizeRelmotion, motion.cpp:367-526)
                ctx.relmotion_batch(B)
                note_ktimes(True)
            if world > 1:
                recs_all = ctx.allgather_poses(None, B)     # graph-insert-time exchange, from the device buffer
                assert len(recs_all) == world * B
            state["found"] += int(recs["found"].sum()); state["pairs"] += B
            if len(state["lines"]) < 4 * B:
                state["lines"] += [f.num_lines for f in frames]

        def finish_pairs():
            """Collects the batch in flight (records, exchange, statistics) and releases its frames."""
            pend = state.get("pending")
            if pend is None:
                return
            frames_p, head_p, recv_p = pend
            recs = ctx.match_pair_batch_end()
            after_pairs(recs, frames_p)
            keep = state["prev"]
            for f in frames_p:
                if f is not keep:
                    f.free()
            for f in (head_p, recv_p):
                if f is not None and f is not keep and f not in frames_p:
                    f.free()
            state["pending"] = None

    def timed(e2e: bool, steps: int):
        for v in ktime.values():
            v.clear()
        state["found"] = state["pairs"] = 0
        st0 = ctx.stats()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        gc.collect()
        gc.disable()              # no collector pauses inside the timed region (592 frame handles are created per step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        w0 = time.time()
        per_step.clear()
        for _ in range(steps):
            ts = time.perf_counter()
            one_step(e2e)
            per_step.append(round(1e3 * (time.perf_counter() - ts), 2))
        if wl != "cfg4" and pipelined:
            finish_pairs()        # the last batch's pair stage completes inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.window(w0, time.time())
        gc.enable()
        ms = e0.elapsed_time(e1)
        ms_rank = ms
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        st1 = ctx.stats()
        return ms, wall, st0, st1, ms_rank

    pipelined = wl != "cfg4" and not args.no_pipeline
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("LSL_BENCH_NOCLOCKS") else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step(False)
    if pipelined:
        finish_pairs()
    ms_dev, wall_dev, s0, s1, _ = timed(False, args.steps)
    steps_dev = list(per_step)
    kt = {k: float(np.mean(v)) for k, v in ktime.items() if v}
    found_frac = state["found"] / max(state["pairs"], 1)
    launches = int(s1.kernel_launches - s0.kernel_launches)
    one_step(True)  # warm the host-buffer path (pinned staging is the caller's here)
    if pipelined:
        finish_pairs()
    ms_e2e, wall_e2e, h0, h1, ms_e2e_rank = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None

    pairs_total = pairs_per_step_total * args.steps
    value = pairs_total / (ms_dev / 1e3)
    e2e_value = pairs_total / (ms_e2e / 1e3)
    h2d_step = int((h1.h2d_bytes - h0.h2d_bytes) // args.steps)
    h2d_gbs = [h2d_step / (ms_e2e_rank / args.steps / 1e3) / 1e9]
    if world > 1:   # per-rank H2D rate: shows a host-side (NUMA / PCIe root) limiter in the scaling run
        g = [None] * world
        dist.all_gather_object(g, h2d_gbs[0])
        h2d_gbs = g
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = _hbm_peak(peaks)
        dom = max(kt, key=kt.get) if kt else "lsd_region_kernel"
        units = B if wl != "cfg4" else max(len(imgs) - (1 if rank == 0 else 0), 1)
        alg_bytes = units * b_frame if wl != "cfg4" else int(np.sum(state["lines"])) * 1040
        achieved = alg_bytes / (kt.get(dom, float("nan")) / 1e3) / 1e9
        ln = np.array(state["lines"] if state["lines"] else [0])
        line = {
            "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak" if wl != "cfg4" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl] + ", 1 B200 per rank", "image": f"{Wd}x{Hd}",
                       ("batch_frames_per_gpu" if wl != "cfg4" else "pairs_per_step"): B, "unique_frames": U if wl != "cfg4" else 257,
                       "ransac_iters": 500,
                       "l2": (f"inputs larger than L2: {B * b_frame / 1e6:.0f} MB of RGB+depth per step per GPU, two alternating batch buffers"
                              if wl != "cfg4" else "feature records of 256 keyframes + LM scratch (> 126 MB per step), flushed by the scratch writes"),
                       "pairs_found_frac": found_frac, "lines_per_frame": float(ln.mean()),
                       "lines_per_frame_spread": [int(ln.min()), int(np.percentile(ln, 50)), int(ln.max())],
                       "render_s": round(render_s, 1),
                       **({"pipeline": "pair stage of batch k on the pair stream under the extraction of batch k + 1 (lsl_match_pair_batch_begin / _end); all K pair batches complete inside the timed region"} if pipelined else {}),
                       **({"split": "one stream, blocks of %d frames dealt round-robin to the ranks, tail records shifted rank to rank by ncclSend / ncclRecv (lsl_shift_frame); received == sent checked: %s" % (B, state.get("shift_checked", False))} if split else {}),
                       **({"point_features": "SIFT detected on the device inside lsl_extract_batch (k_sift.cu), max 600, RootSIFT" if dev_sift
                           else "cv2 SIFT on the host, uploaded with lsl_frames_set_points_batch, RootSIFT on the device"} if wl == "cfg3" else {})},
            "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d_step,
                    "d2h_bytes_per_step": int((h1.d2h_bytes - h0.d2h_bytes) // args.steps), "ms_per_step": ms_e2e / args.steps,
                    "h2d_gbs_per_rank": [round(float(x), 2) for x in h2d_gbs],
                    "input": "pinned host RGB u8 + 16-bit depth (TUM PNG values), converted on the device" if wl != "cfg4" else "features resident (pre-distributed keyframes); query record broadcast"},
            "gpu_launches": launches,
            "host_ms_each_step": {"value": steps_dev, "e2e": list(per_step)},
            "kernel_ms_per_step": kt,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (NCU_PIPE.get(dom, {}).get("dram_bytes_per_frame", 0) * units or None) if wl != "cfg4" else None,
                         "traffic_unit": "bytes per launch (ncu dram read + write of this build, profiles/r2_ncu_pipe.json)",
                         "algorithmic_bytes": alg_bytes,
                         "actual_bound": {"kind": "fp64 dependency latency", "kernel": dom, **{k: v for k, v in NCU_PIPE.get(dom, {}).items() if k != "dram_bytes_per_frame"},
                                          "source": NCU_PIPE.get("_source", "profiles/r2_ncu_pipe.json")},
                         "note": f"algorithmic bytes = {units} frames x {b_frame} B (RGB u8 + depth f32) per launch / CUDA-event time of "
                                 f"{dom}; peak = MEASURED_PEAKS.json hbm_gbs ({'measured' if peaks else 'fallback 6650 GB/s of B200_PROFILING.md'}); "
                                 f"the path is FP64-latency / dependency bound (100 LM iterations per line, sequential region "
                                 f"growing), not HBM bound: see DESIGN.md section 4 and profiles/r2_summary.md"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu and wl == "cfg2":
            threads = os.cpu_count() or 1
            pps, sec, pairs, _ = cpu_pairs_per_sec(imgs, deps, K, threads * 8, threads)
            line["cpu_baseline"] = {"value": pps, "unit": "frame-pairs/s", "cores": threads, "kind": "port", "per_core": pps / threads,
                                    "sample": f"{pairs} pairs of the same stream: {threads} independent single-thread streams "
                                              f"of the oracle restatement, {sec:.1f} s of wall time"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("LSL_BENCH_BATCH", 0)), help="frames per step per GPU (0: the workload's default, 1184 = 8 x 148 SMs for cfg2)")
    ap.add_argument("--no-pipeline", action="store_true", help="run the pair stage of a batch to completion before the next extraction (default: lsl_match_pair_batch_begin / _end, the pair stage of batch k runs under the extraction of batch k + 1)")
    ap.add_argument("--split", default="streams", choices=["streams", "stream"], help="cfg2 on N > 1 GPUs: one independent stream per rank (default), or ONE stream dealt block-wise to the ranks with the block tails shifted rank to rank (lsl_shift_frame)")
    ap.add_argument("--points", default="device", choices=["device", "host"], help="cfg3: SIFT on the device inside the extract call (default) or cv2 SIFT on the host, uploaded with lsl_frames_set_points_batch")
    ap.add_argument("--unique", type=int, default=296, help="distinct consecutive rendered frames of the stream (tiled palindromically into a batch)")
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS), help="BASELINE.json config (cfg2 = the metric's config, the default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
