"""Writes the golden vectors of tests/golden (run in the build container, where /root/reference is mounted and
oracle/_ref holds the unmodified upstream builds):
  lsd_upstream_scene2000_f{0,1}.npy   segment lists of the UNMODIFIED external/lsd/lsd-1.5/lsd.c on the first two
                                      frames of the cfg-2 synthetic stream (gray conversion: OpenCV 2.4 formula)
  ref_tum_frame.png, ref_chairs.png   the reference's own test images (external/lsd/lsd-1.5/1305031453.359684.png copied
                                      byte for byte; chairs.pgm re-encoded as a lossless PNG) — category-b fixtures
  lsd_upstream_ref_{tum,chairs}.npy   segment lists of the UNMODIFIED upstream lsd.c on those two images (gray of the TUM
                                      frame: OpenCV 2.4 formula on the BGR planes as cv::imread delivers them)
  oracle_digest.json                  sha256 of the oracle's line records / matches / pose for frames 0,1 (regression
                                      guard of the restated stages that have no reference-run counterpart)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from refimpl import lsd_ref  # noqa: E402
from lineslam_b200 import synth  # noqa: E402
from oracle import pyoracle as o  # noqa: E402

imgs, deps, poses = synth.make_stream(2, scene_seed=2000)
K = synth.camera_K()
for i in range(2):
    np.save(os.path.join(HERE, f"lsd_upstream_scene2000_f{i}.npy"), lsd_ref(o.gray(imgs[i])))
L = [o.detect3DLines(imgs[i], deps[i], K, seed=i + 1) for i in range(2)]
m = o.lineMatching(L[1], L[0], True)
rec, inl, rinl, _ = o.pose_ransac(L[0], L[1], m, id_train=0, id_query=1, seed=1)
dig = {"lines0": hashlib.sha256(L[0].tobytes()).hexdigest(), "lines1": hashlib.sha256(L[1].tobytes()).hexdigest(),
       "matches": hashlib.sha256(m.tobytes()).hexdigest(), "ransac_inliers": hashlib.sha256(rinl.tobytes()).hexdigest(),
       "n_lines": [len(L[0]), len(L[1])], "n_matches": len(m), "n_ransac_inliers": len(rinl), "n_inliers": len(inl),
       "tf": [float(v) for v in rec["tf"]], "input_sha": hashlib.sha256(imgs.tobytes() + deps.tobytes()).hexdigest()}
json.dump(dig, open(os.path.join(HERE, "oracle_digest.json"), "w"), indent=1)
print(dig)

# ---- the reference's own images (only where /root/reference is mounted)
REF_IMG = "/root/reference/external/lsd/lsd-1.5"
if os.path.isdir(REF_IMG):
    import shutil
    import cv2
    shutil.copyfile(os.path.join(REF_IMG, "1305031453.359684.png"), os.path.join(HERE, "ref_tum_frame.png"))
    chairs = cv2.imread(os.path.join(REF_IMG, "chairs.pgm"), cv2.IMREAD_GRAYSCALE)
    cv2.imwrite(os.path.join(HERE, "ref_chairs.png"), chairs, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert np.array_equal(cv2.imread(os.path.join(HERE, "ref_chairs.png"), cv2.IMREAD_GRAYSCALE), chairs)
    tum = cv2.imread(os.path.join(HERE, "ref_tum_frame.png"), cv2.IMREAD_COLOR)
    np.save(os.path.join(HERE, "lsd_upstream_ref_tum.npy"), lsd_ref(o.gray(tum)))
    np.save(os.path.join(HERE, "lsd_upstream_ref_chairs.npy"), lsd_ref(chairs))
