"""Host-side phase timing of the stream step (probe, not a bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lineslam_b200 import api, synth
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
imgs, deps, K = bench.make_unique_frames(8, 0)
order = bench.palindrome(8, B)
bi = np.stack([imgs[i] for i in order]); bd = np.stack([deps[i] for i in order])
hi = torch.from_numpy(bi).pin_memory(); hd = torch.from_numpy(bd).pin_memory()
di, dd = hi.cuda(), hd.cuda()
ctx = api.Context(max_batch=B)
prev = None
for s in range(6):
    e2e = s >= 3
    seeds = np.arange(1, B + 1, dtype=np.uint32)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if e2e: frames = ctx.extract_batch(hi.numpy(), hd.numpy(), K, seeds)
    else: frames = ctx.extract_batch_dev(di.data_ptr(), 3, dd.data_ptr(), B, 640, 480, K, seeds)
    t1 = time.perf_counter(); dev_ext = ctx.last_timing()[0]
    trains = [prev if prev is not None else frames[0]] + frames[:-1]
    ids = np.arange(B, dtype=np.int32) + 1
    recs = ctx.match_pair_batch(frames, trains, ids, ids - 1, seeds)
    t2 = time.perf_counter(); dev_pair = ctx.last_timing()[0]
    old = prev; prev = frames[-1]
    for f in frames[:-1]: f.free()
    if old is not None: old.free()
    t3 = time.perf_counter()
    print(f"step {s} e2e={e2e}: extract {1e3*(t1-t0):.1f} ms (device {dev_ext:.1f}) pair {1e3*(t2-t1):.1f} ms (device {dev_pair:.1f}) free {1e3*(t3-t2):.1f} ms", flush=True)
kt = ctx.kernel_times(); print({k: round(v, 2) for k, v in kt.items()})
