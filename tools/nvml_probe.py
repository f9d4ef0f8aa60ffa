import sys, os, time, threading, subprocess
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from lineslam_b200 import api
import bench, gc
mode = sys.argv[1]
B = 592
imgs, deps, K = bench.make_unique_frames(8, 0)
order = bench.palindrome(8, B)
bi = np.stack([imgs[i] for i in order]); bd = np.stack([deps[i] for i in order])
hi = torch.from_numpy(bi).pin_memory(); hd = torch.from_numpy(bd).pin_memory()
ctx = api.Context(max_batch=B)
halt = threading.Event(); lat = []
def nvml_thread(which):
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not halt.is_set():
        t0 = time.perf_counter()
        if which in ("clock", "both"): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        t1 = time.perf_counter()
        if which in ("reasons", "both"): pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        t2 = time.perf_counter()
        lat.append((1e3*(t1-t0), 1e3*(t2-t1)))
        halt.wait(0.5)
proc = None
if mode in ("clock", "reasons", "both"):
    th = threading.Thread(target=nvml_thread, args=(mode,), daemon=True); th.start()
elif mode == "smi":
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv", "-lms", "500"], stdout=open("/tmp/clk.csv", "w"))
    time.sleep(2)
prev = None; times = []
gc.disable()
for s in range(16):
    seeds = np.arange(1, B + 1, dtype=np.uint32)
    t0 = time.perf_counter()
    frames = ctx.extract_batch(hi.numpy(), hd.numpy(), K, seeds)
    trains = [prev if prev is not None else frames[0]] + frames[:-1]
    ids = np.arange(B, dtype=np.int32) + 1
    recs = ctx.match_pair_batch(frames, trains, ids, ids - 1, seeds)
    old = prev; prev = frames[-1]
    for f in frames[:-1]: f.free()
    if old is not None: old.free()
    times.append(round(1e3*(time.perf_counter()-t0), 1))
halt.set()
if proc: proc.terminate(); print(open("/tmp/clk.csv").read()[-300:])
print(mode, times)
if lat: print("nvml call ms max (clock, reasons):", max(l[0] for l in lat), max(l[1] for l in lat), "n", len(lat))
