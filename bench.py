#!/usr/bin/env python
"""bench.py — frame-pairs/sec of the line front end (extract + match + RANSAC pose), BASELINE.json's metric.

Workload (configs[1]): the synthetic fr1/xyz-shape 640x480 RGB-D stream, line-only odometry. One "step" is
one batch of B consecutive frames of the stream: every frame is extracted once (LSD -> 3D lines -> MSLD ->
MLE), matched against its predecessor (the predecessor of the first frame is the cached last frame of the
previous step) and registered (500-iteration RANSAC + LM refinement) -> B pose records.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]      the CUDA path through the C ABI
  python bench.py --impl reference ...                                 the reference's CPU path (oracle/)

`value`  : device-resident inputs (already in HBM when the timed region starts).
`e2e`    : same steps through the host-buffer entry point: pinned host RGB+depth -> H2D inside the timed
           region, pose records D2H (what Node::Node + Node::matchNodePair callers see).
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
# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at batch 592
# (profiles/r1_final2_ncu_full_raw_b592.csv for the first three kernels, profiles/r1_end_ncu_full_raw_b592.csv for the
# others), per frame of the launch
NCU_DRAM_BYTES_PER_FRAME = {"line_mle_kernel": (0.1764e9 + 0.1708e9) / 592, "lsd_region_kernel": (1.4525e9 + 0.0903e9) / 592,
                            "line3d_ransac_kernel": (0.3238e9 + 0.1165e9) / 592, "lsd_nfa_kernel": (0.2759e9 + 0.0110e9) / 592,
                            "ll_angle_kernel": (0.9328e9 + 3.925e9) / 592, "line_msld_kernel": (0.6691e9 + 0.0586e9) / 592}


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


def make_unique_frames(u: int, rank: int):
    from lineslam_b200 import synth
    imgs, deps, _ = synth.make_stream(u, scene_seed=2000, start=rank * 7)
    return imgs, deps, synth.camera_K(W, H)


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


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    imgs, deps, K = make_unique_frames(args.unique, 0)
    per_step = threads * 2
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_pairs_per_sec(imgs, deps, K, threads, threads)
    t0 = time.perf_counter()
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
        "config": {"workload": "cfg2 fr1/xyz-shape synthetic stream 640x480, line-only odometry (CPU reference path)",
                   "pairs_per_step": per_step, "unique_frames": args.unique},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} pairs per step: {threads} independent single-thread streams x 2 pairs "
                                   f"(oracle restatement of detect3DLines + lineMatching + getTransform_PtsLines_ransac)"},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm ----
def run_cuda(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from lineslam_b200 import api
    from lineslam_b200.records import POSE_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, U = args.batch, args.unique
    imgs, deps, K = make_unique_frames(U, rank)
    ctx = api.Context(device=local_rank, max_batch=B, max_w=W, max_h=H)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:  # library-owned NCCL communicator; the id travels over torch.distributed (plumbing)
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], world, rank)

    # host (pinned) and device copies of two batch layouts so that consecutive steps read different addresses
    def batch_arrays(phase):
        order = palindrome(U, B, phase)
        return np.stack([imgs[i] for i in order]), np.stack([deps[i] for i in order])
    nbuf = 2
    host_i, host_d, dev_i, dev_d = [], [], [], []
    for p in range(nbuf):
        bi, bd = batch_arrays(p * B)
        hi = torch.from_numpy(bi).pin_memory()
        hd = torch.from_numpy(bd).pin_memory()
        host_i.append(hi); host_d.append(hd)
        dev_i.append(hi.cuda(non_blocking=False)); dev_d.append(hd.cuda(non_blocking=False))
    torch.cuda.synchronize()

    state = {"prev": None, "step": 0, "found": 0, "pairs": 0, "lines": 0}
    per_step = []
    ktime = {}

    def one_step(e2e: bool):
        s = state["step"]
        b = s % nbuf
        seeds = np.arange(1, B + 1, dtype=np.uint32) + s * B
        if e2e:
            frames = ctx.extract_batch(host_i[b].numpy(), host_d[b].numpy(), K, seeds)
        else:
            frames = ctx.extract_batch_dev(dev_i[b].data_ptr(), 3, dev_d[b].data_ptr(), B, W, H, K, seeds)
        for k, v in ctx.kernel_times().items():
            if v > 0 and not k.startswith(("match", "pose")):
                ktime.setdefault(k, []).append(v)
        trains = [state["prev"]] + frames[:-1] if state["prev"] is not None else [frames[0]] + frames[:-1]
        ids = np.arange(B, dtype=np.int32) + s * B + 1
        recs = ctx.match_pair_batch(frames, trains, ids, ids - 1, seeds)
        for k, v in ctx.kernel_times().items():
            if v > 0 and k.startswith(("match", "pose")):
                ktime.setdefault(k, []).append(v)
        if world > 1:
            recs_all = ctx.allgather_poses(recs)     # graph-insert-time exchange
            assert len(recs_all) == world * B
        state["found"] += int(recs["found"].sum()); state["pairs"] += B
        state["lines"] += sum(f.num_lines for f in frames[:4])
        old = state["prev"]
        state["prev"] = frames[-1]
        for f in frames[:-1]:
            f.free()
        if old is not None:
            old.free()
        state["step"] += 1
        return recs

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
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.window(w0, time.time())
        gc.enable()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        st1 = ctx.stats()
        return ms, wall, st0, st1

    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("LSL_BENCH_NOCLOCKS") else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step(False)
    ms_dev, wall_dev, s0, s1 = timed(False, args.steps)
    steps_dev = list(per_step)
    kt = {k: float(np.mean(v)) for k, v in ktime.items() if v}
    found_frac = state["found"] / max(state["pairs"], 1)
    launches = int(s1.kernel_launches - s0.kernel_launches)
    one_step(True)  # warm the host-buffer path (pinned staging is the caller's here)
    ms_e2e, wall_e2e, h0, h1 = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None

    pairs_total = world * B * args.steps
    value = pairs_total / (ms_dev / 1e3)
    e2e_value = pairs_total / (ms_e2e / 1e3)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = _hbm_peak(peaks)
        dom = max(kt, key=kt.get) if kt else "lsd_region_kernel"
        achieved = B * B_FRAME / (kt.get(dom, float("nan")) / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2 fr1/xyz-shape synthetic stream 640x480, line-only odometry, 1 B200 per rank",
                       "batch_frames_per_gpu": B, "unique_frames": U, "ransac_iters": 500,
                       "l2": f"inputs larger than L2: {B * B_FRAME / 1e6:.0f} MB of RGB+depth per step per GPU, two alternating batch buffers",
                       "pairs_found_frac": found_frac, "lines_per_frame": state["lines"] / max(4 * (state["step"]), 1)},
            "e2e": {"value": e2e_value, "unit": "frame-pairs/s",
                    "h2d_bytes_per_step": int((h1.h2d_bytes - h0.h2d_bytes) // args.steps),
                    "d2h_bytes_per_step": int((h1.d2h_bytes - h0.d2h_bytes) // args.steps), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "host_ms_each_step": {"value": steps_dev, "e2e": list(per_step)},
            "kernel_ms_per_step": kt,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (NCU_DRAM_BYTES_PER_FRAME[dom] * B if dom in NCU_DRAM_BYTES_PER_FRAME else None),
                         "traffic_unit": "bytes per launch (ncu dram read + write, profiles/r1_final2_ncu_full_raw_b592.csv)",
                         "algorithmic_bytes": B * B_FRAME,
                         "note": f"algorithmic bytes = {B} frames x {B_FRAME} B (RGB u8 + depth f32) per launch / CUDA-event time of "
                                 f"{dom}; peak = MEASURED_PEAKS.json hbm_gbs ({'measured' if peaks else 'fallback 6650 GB/s of B200_PROFILING.md'}); "
                                 f"the path is FP64-latency / dependency bound (100 LM iterations per line, sequential region "
                                 f"growing), not HBM bound: see DESIGN.md section 4 and profiles/r1_summary.md"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            pps, sec, pairs, _ = cpu_pairs_per_sec(imgs, deps, K, threads * 8, threads)
            line["cpu_baseline"] = {"value": pps, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
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
    ap.add_argument("--batch", type=int, default=int(os.environ.get("LSL_BENCH_BATCH", 592)), help="frames per step per GPU")
    ap.add_argument("--unique", type=int, default=12, help="distinct rendered frames (tiled palindromically into a batch)")
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
