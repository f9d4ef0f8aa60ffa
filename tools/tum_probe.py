"""Probe (not a bench): wall time and png_unfilter_kernel time of lsl_tum_decode_batch on n VGA frames
(PNG files written by Pillow from the synthetic stream, so zlib picks the filters a real TUM file would have)."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from lineslam_b200 import api, synth, tum

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
imgs, deps, _ = synth.make_stream(8, scene_seed=2000)
rgb, dep = [], []
for i in range(8):
    b = io.BytesIO(); Image.fromarray(np.ascontiguousarray(imgs[i][:, :, ::-1])).save(b, format="PNG"); rgb.append(b.getvalue())
    b = io.BytesIO(); Image.fromarray(np.rint(np.nan_to_num(deps[i].astype(np.float64)) * 5000).astype(np.uint16)).save(b, format="PNG"); dep.append(b.getvalue())
rgb = (rgb * (n // 8 + 1))[:n]; dep = (dep * (n // 8 + 1))[:n]
print("file bytes per frame: rgb", len(rgb[0]), "depth", len(dep[0]))
ctx = api.Context(max_batch=8, max_w=640, max_h=480)
d_bgr = torch.zeros((n, 480, 640, 3), dtype=torch.uint8, device="cuda")
d_dep = torch.zeros((n, 480, 640), dtype=torch.float32, device="cuda")
for it in range(3):
    t0 = time.time()
    tum.decode_batch(ctx, rgb, dep, 640, 480, d_bgr.data_ptr(), d_dep.data_ptr())
    torch.cuda.synchronize()
    t1 = time.time()
    kt = ctx.kernel_times()
    k = kt.get("png_unfilter_kernel", 0.0) + kt.get("png_unfilter_kernel(depth)", 0.0)
    ki = kt.get("png_inflate_kernel", 0.0) + kt.get("png_inflate_kernel(depth)", 0.0)
    byt = n * (480 * (1 + 640 * 3) + 480 * (1 + 640 * 2) + 640 * 480 * 3 + 640 * 480 * 4)
    print(f"n={n}: wall {1e3*(t1-t0):.1f} ms ({n/(t1-t0):.0f} frames/s), png_inflate_kernel {ki:.2f} ms, png_unfilter_kernel {k:.2f} ms -> {byt/k/1e6:.0f} GB/s of algorithmic traffic")
tum.release(ctx); ctx.close()
