"""Probe: extraction of step s+1 (context A, thread 1) overlapped with the pair stage of step s (context B, thread 2)."""
import sys, os, time, threading, queue
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lineslam_b200 import api
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
imgs, deps, K = bench.make_unique_frames(12, 0)
order = bench.palindrome(12, B)
bi = np.stack([imgs[i] for i in order]); bd = np.stack([deps[i] for i in order])
di, dd = torch.from_numpy(bi).cuda(), torch.from_numpy(bd).cuda()
ca, cb = api.Context(max_batch=B), api.Context(max_batch=B)
q = queue.Queue(maxsize=2)
def producer(n):
    for s in range(n):
        seeds = np.arange(1, B + 1, dtype=np.uint32) + s
        q.put((s, ca.extract_batch_dev(di.data_ptr(), 3, dd.data_ptr(), B, 640, 480, K, seeds), seeds))
    q.put(None)
found = []
def consumer():
    prev = None
    while True:
        item = q.get()
        if item is None: break
        s, frames, seeds = item
        trains = [prev if prev is not None else frames[0]] + frames[:-1]
        ids = np.arange(B, dtype=np.int32) + 1
        recs = cb.match_pair_batch(frames, trains, ids, ids - 1, seeds)
        found.append(int(recs["found"].sum()))
        old = prev; prev = frames[-1]
        for f in frames[:-1]: f.free()
        if old is not None: old.free()
for n in (3, STEPS):   # warm-up, then timed
    found.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tp = threading.Thread(target=producer, args=(n,)); tc = threading.Thread(target=consumer)
    tp.start(); tc.start(); tp.join(); tc.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pipelined extract || pair: {STEPS} steps x {B} frames in {dt*1e3:.1f} ms -> {STEPS*B/dt:.1f} pairs/s ({dt*1e3/STEPS:.1f} ms/step); found {found}")
