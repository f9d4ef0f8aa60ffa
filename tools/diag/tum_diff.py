"""Diagnostic: where does detect3DLines on the reference's TUM frame differ between the device and the oracle?"""
import os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lineslam_b200 import api
from oracle import pyoracle as po
tum = cv2.imread(os.path.join(ROOT, "tests/golden/ref_tum_frame.png"), cv2.IMREAD_COLOR)
H, W = tum.shape[:2]
yy, xx = np.mgrid[0:H, 0:W]
rng = np.random.default_rng(21)
dep = (1.2 + 0.002 * xx + 0.0015 * yy).astype(np.float32)
dep = (np.round(dep * 5000) / 5000).astype(np.float32)
dep[rng.random(dep.shape) < 0.05] = np.nan
K = np.array([[525., 0, 319.5], [0, 525., 239.5], [0, 0, 1]])
ctx = api.Context(max_batch=1, max_w=W, max_h=H, debug=True)
fr = ctx.extract_batch(tum[None], dep[None], K, seeds=[9])[0]
ref, dref = po.detect3DLines(tum, dep, K, seed=9, debug=True)
got, dg = fr.lines(), fr.debug()
print("lines", len(got), len(ref), "segs equal", np.array_equal(fr.segments(), dref["segs"]))
n = min(len(got), len(ref))
print("seg_of_line equal", np.array_equal(dg["seg_of_line"][:n], dref["seg_of_line"][:n]))
bad = [i for i in range(n) if dg["n_inl"][i] != dref["n_inl"][i] or not np.array_equal(dg["inl_idx"][i, :dref["n_inl"][i]], dref["inl_idx"][i, :dref["n_inl"][i]])]
print("lines with different inlier sets:", bad[:20], len(bad))
for i in bad[:3]:
    print(i, "seg", dref["seg_of_line"][i], "n_inl", dg["n_inl"][i], dref["n_inl"][i], "p", ref["p"][i], "q", ref["q"][i])
    print(" got", dg["inl_idx"][i, :dg["n_inl"][i]])
    print(" ref", dref["inl_idx"][i, :dref["n_inl"][i]])
if len(got) != len(ref):
    a = set(dg["seg_of_line"].tolist()); b = set(dref["seg_of_line"].tolist())
    print("only got", sorted(a - b), "only ref", sorted(b - a))
for name in ["p", "q", "lineEq2d", "r", "des", "lid", "haveDepth", "A", "B", "covA", "covB"]:
    g, r = got[name], ref[name]
    neq = ~((g == r) | (np.isnan(g) & np.isnan(r)))
    rows = np.unique(np.nonzero(neq)[0])
    print(name, "rows differing", rows[:10], len(rows), "nan rows", int(np.isnan(r).reshape(len(r), -1).any(1).sum()))
    for i in rows[:2]:
        print("  ", i, "got", np.ravel(g[i])[:6], "ref", np.ravel(r[i])[:6], "p", ref["p"][i], "q", ref["q"][i])
