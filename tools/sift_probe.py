"""Dumps the device SIFT output of the reference TUM frame (gpurun_out/sift_dump.npz) for offline comparison with the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lineslam_b200 import api
import test_gpu_sift as T
tum, dep, K = T._tum_inputs()
H, W = tum.shape[:2]
ctx = api.Context(max_batch=1, max_w=W, max_h=H)
ctx.set_point_detector("SIFT", int(sys.argv[1]) if len(sys.argv) > 1 else 2000, root_sift=False)
fr = ctx.extract_batch(tum[None], dep[None], K, seeds=[9])[0]
xyz, desc, kp = fr.points()
print("n", fr.num_points, flush=True)
t = time.time()
fr2 = ctx.extract_batch(tum[None], dep[None], K, seeds=[9])[0]
print("second extract", time.time() - t, ctx.kernel_times().get("sift_kernels"), flush=True)
xyz2, desc2, kp2 = fr2.points()
print("identical", np.array_equal(kp, kp2), np.array_equal(desc, desc2), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "sift_dump.npz"), xyz=xyz, desc=desc, kp=kp)
