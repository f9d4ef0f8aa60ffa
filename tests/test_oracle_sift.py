"""SURVEY.md §8f row 3 — the oracle of the point-feature half of Node::Node (oracle/oracle_sift.py: OpenCV's SIFT
restated in numpy) pinned against cv2.SIFT_create of the OpenCV in this image: on the reference's own TUM frame every
cv2 keypoint is reproduced (position < 0.01 px, size, orientation, response), descriptors agree to +-1 of 255 per
element (float summation order), and the extra keypoints are exactly the duplicates cv2 removes."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle_sift as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _match(kps, ck):
    from scipy.spatial import cKDTree
    A = np.array([[k["x"], k["y"], k["size"], k["angle"], k["response"]] for k in kps])
    B = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ck])
    d, idx = cKDTree(A[:, :2]).query(B[:, :2], k=6)
    pairs = []
    for i in range(len(B)):
        for dd, j in zip(d[i], idx[i]):
            da = abs(A[j, 3] - B[i, 3])
            if dd < 0.01 and abs(A[j, 2] - B[i, 2]) < 1e-3 * B[i, 2] and min(da, 360 - da) < 0.05:
                pairs.append((i, j)); break
    return A, B, pairs


def test_sift_restatement_reproduces_cv2_on_the_reference_tum_frame():
    tum = cv2.imread(os.path.join(GOLD, "ref_tum_frame.png"), cv2.IMREAD_COLOR)
    g = cv2.cvtColor(tum, cv2.COLOR_BGR2GRAY)
    pyr = S.build_pyramid(g)
    kps = S.detect(g, pyr)
    ck, cd = cv2.SIFT_create().detectAndCompute(g, None)
    assert len(ck) > 600
    A, B, pairs = _match(kps, ck)
    assert len(pairs) == len(ck)                                   # every cv2 keypoint
    assert len(S.remove_duplicates_and_retain_best(kps, 0)) == len(ck)   # the surplus are cv2's removed duplicates
    assert max(abs(A[j, 4] - B[i, 4]) / B[i, 4] for i, j in pairs) < 1e-4
    worst, exact = 0.0, 0
    for i, j in pairs[::3]:
        e = np.abs(S.describe(pyr[0], kps[j]) - cd[i]).max()
        worst = max(worst, e); exact += e == 0
    assert worst <= 1.0 and exact > 0.8 * len(pairs[::3])
    # retainBest(600) as Node::Node applies it
    best = S.remove_duplicates_and_retain_best(kps, 600)
    cbest = cv2.SIFT_create(nfeatures=600).detect(g, None)
    assert abs(len(best) - len(cbest)) <= 2
    rb = sorted(k["response"] for k in best)[:5]; rc = sorted(k.response for k in cbest)[:5]
    assert np.allclose(rb, rc, rtol=1e-4)


def test_sift_restatement_on_a_rendered_frame_and_projection():
    from lineslam_b200 import synth
    imgs, deps, _ = synth.make_stream(1, scene_seed=2000)
    g = cv2.cvtColor(imgs[0], cv2.COLOR_BGR2GRAY)
    kps = S.detect(g)
    ck = cv2.SIFT_create().detect(g, None)
    _, _, pairs = _match(kps, ck)
    assert len(pairs) >= len(ck) - 3 and len(ck) > 100            # the two coarsest-octave keypoints may differ (4-row images)
    K = synth.camera_K()
    keep, xyz = S.project_to_3d(kps, deps[0], K)
    assert 0 < len(keep) <= 600 and np.all(np.isfinite(xyz))
    k0 = kps[keep[0]]
    z = deps[0][int(np.rint(k0["y"])), int(np.rint(k0["x"]))]
    assert xyz[0, 2] == z and abs(xyz[0, 0] - (k0["x"] - K[0, 2]) * z / K[0, 0]) < 1e-5
