"""Probe: two contexts driven from two host threads (kernels of two batches overlap on the GPU)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lineslam_b200 import api
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
NCTX = int(sys.argv[2]) if len(sys.argv) > 2 else 2
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
STAGGER = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
imgs, deps, K = bench.make_unique_frames(8, 0)
order = bench.palindrome(8, B)
bi = np.stack([imgs[i] for i in order]); bd = np.stack([deps[i] for i in order])
di, dd = torch.from_numpy(bi).cuda(), torch.from_numpy(bd).cuda()
ctxs = [api.Context(max_batch=B) for _ in range(NCTX)]
def worker(ctx, nsteps, out, delay=0.0):
    time.sleep(delay)
    prev = None
    for s in range(nsteps):
        seeds = np.arange(1, B + 1, dtype=np.uint32)
        frames = ctx.extract_batch_dev(di.data_ptr(), 3, dd.data_ptr(), B, 640, 480, K, seeds)
        trains = [prev if prev is not None else frames[0]] + frames[:-1]
        ids = np.arange(B, dtype=np.int32) + 1
        recs = ctx.match_pair_batch(frames, trains, ids, ids - 1, seeds)
        old = prev; prev = frames[-1]
        for f in frames[:-1]: f.free()
        if old is not None: old.free()
        out.append(int(recs["found"].sum()))
for c in ctxs: worker(c, 2, [])   # warm-up
torch.cuda.synchronize()
outs = [[] for _ in ctxs]
t0 = time.perf_counter()
ths = [threading.Thread(target=worker, args=(c, STEPS, o, k * STAGGER)) for k, (c, o) in enumerate(zip(ctxs, outs))]
for t in ths: t.start()
for t in ths: t.join()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{NCTX} contexts x {STEPS} steps x {B} frames: {dt*1e3:.1f} ms -> {NCTX*STEPS*B/dt:.1f} pairs/s; found {[sum(o) for o in outs]}")
