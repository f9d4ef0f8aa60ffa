import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lineslam_b200 import api, synth
imgs, deps, _ = synth.make_stream(1, scene_seed=2001, W=320, H=240)
ctx = api.Context(max_batch=1, max_w=320, max_h=240, debug=True)
fr = ctx.extract_batch(imgs, deps, synth.camera_K(320, 240), seeds=[1])
print("ok", fr[0].num_lines)
