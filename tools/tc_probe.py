"""Probe: featureMatching of a loop-closure-shaped batch (1 query x 256 keyframes, 600 x 600 x 128 RootSIFT rows per pair) with
the tensor-core pre-filter (default) or the exact scalar kernel (LSL_MATCH_TC=0). Prints the device time of the matcher
(CUDA events of the library: LSL_K_MATCHPTS) and the candidate statistics. Usage: python tools/tc_probe.py [npairs]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lineslam_b200 import api, synth
from lineslam_b200.records import LINE_DTYPE

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(5)
P, base = synth.make_landmarks(2000, n=4000, dim=128)
ctx = api.Context(max_batch=1, max_w=64, max_h=64)
def frame(seed):
    r = np.random.default_rng(seed)
    idx = r.permutation(4000)[:600]
    d = np.abs(base[idx] + r.normal(0, 0.01, (600, 128)).astype(np.float32))
    d = np.sqrt(d / d.sum(1, keepdims=True)).astype(np.float32)
    x = np.ones((600, 4), np.float32); x[:, :3] = r.normal(0, 1, (600, 3))
    # two dummy lines so that the pair stage has frames to work on
    return ctx.frame_from_lines(np.zeros(0, LINE_DTYPE)).set_points(x, d)
q = frame(1)
kfs = [frame(100 + i) for i in range(npairs)]
ids = np.arange(npairs, dtype=np.int32)
times = []
for rep in range(5):
    ctx.match_pair_batch([q] * npairs, kfs, np.full(npairs, 1000, np.int32), ids, ids.astype(np.uint32) + 1)
    times.append(ctx.kernel_times().get("match_points_kernel", 0.0))
ev, full, rows = ctx.match_tc_stats()
print(f"LSL_MATCH_TC={os.environ.get('LSL_MATCH_TC', '1')} npairs={npairs} match_points ms (5 runs) {[round(t, 3) for t in times]} "
      f"exact evals/row {ev / max(rows, 1):.1f} of 600, rows rescanned {full} of {rows}, "
      f"GFLOP {2 * 600 * 600 * 128 * npairs / 1e9:.2f}")
ctx.close()
