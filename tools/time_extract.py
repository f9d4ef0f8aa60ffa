"""Times lsl_extract_batch_dev at several batch sizes (device-resident inputs). Not a bench: a probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lineslam_b200 import api, synth

B = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8,64").split(",")]
imgs, deps, poses = synth.make_stream(8, scene_seed=2000)
K = synth.camera_K()
for b in B:
    ctx = api.Context(max_batch=b, max_w=640, max_h=480)
    reps = (b + 7) // 8
    di = torch.from_numpy(np.concatenate([imgs] * reps)[:b]).cuda()
    dd = torch.from_numpy(np.concatenate([deps] * reps)[:b]).cuda()
    torch.cuda.synchronize()
    for it in range(3):
        t0 = time.time()
        fr = ctx.extract_batch_dev(di.data_ptr(), 3, dd.data_ptr(), b, 640, 480, K, seeds=np.arange(1, b + 1))
        t1 = time.time()
        tot, rg = ctx.last_timing()
        print(f"batch {b}: wall {1e3*(t1-t0):.1f} ms  device {tot:.1f} ms  region_grow {rg:.1f} ms  "
              f"segs {[len(f.segments()) for f in fr[:4]]} lines {[f.num_lines for f in fr[:4]]}  -> {b/(tot/1e3):.1f} frames/s", flush=True)
        del fr
    ctx.close()
