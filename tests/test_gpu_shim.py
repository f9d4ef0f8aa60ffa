"""The compiled drop-in (shim/lsl_adapter.cpp + shim/shim_driver.cpp: Node::detect3DLines on two frames, Node::lineMatching,
getTransform_PtsLines_ransac, exactly the calls Node::Node / Node::matchNodePair make) gives the records of the ctypes
binding — and therefore of the oracle — bit for bit."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_signatures_through_the_adapter(api, oracle, small_frames, tmp_path):
    imgs, deps, poses, K = small_frames
    H, W = deps.shape[1:]
    shim = os.path.join(ROOT, "shim")
    subprocess.check_call(["make", "-C", shim, "all"], stdout=subprocess.DEVNULL)
    grays = [oracle.gray(im) for im in imgs]
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([W, H, 2], np.int32).tobytes())
        f.write(np.ascontiguousarray(K, np.float64).tobytes())
        for g, d in zip(grays, deps):
            f.write(g.tobytes()); f.write(np.ascontiguousarray(d, np.float32).tobytes())
    r = subprocess.run([os.path.join(shim, "shim_driver"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    off = 0
    p = api.default_params(); p.min_feature_matches = 10
    ref = [oracle.detect3DLines(grays[i], deps[i], K, seed=1 + i, params=p) for i in range(2)]
    for i in range(2):
        n = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
        assert n == len(ref[i]) > 0
        rows = np.frombuffer(buf, np.float64, n * 93, off).reshape(n, 93); off += n * 93 * 8
        assert np.array_equal(rows[:, 0:2], ref[i]["p"]) and np.array_equal(rows[:, 2:4], ref[i]["q"])
        assert np.array_equal(rows[:, 4:6], ref[i]["r"])
        assert np.array_equal(rows[:, 6:9], ref[i]["A"]) and np.array_equal(rows[:, 9:12], ref[i]["B"])
        assert np.array_equal(rows[:, 12:21], ref[i]["DU_A"].reshape(n, 9))
        assert np.array_equal(rows[:, 21:], ref[i]["des"], equal_nan=True)
    nm = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    m_ref = oracle.lineMatching(ref[1], ref[0], True)
    assert nm == len(m_ref) > 0
    mt = np.frombuffer(buf, np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]), nm, off); off += 12 * nm
    assert np.array_equal(mt["q"], m_ref["queryIdx"]) and np.array_equal(mt["t"], m_ref["trainIdx"]) and np.array_equal(mt["d"], m_ref["distance"])
    found, n_inl = np.frombuffer(buf, np.int32, 2, off); off += 8
    rmse = np.frombuffer(buf, np.float32, 1, off)[0]; off += 4
    tf = np.frombuffer(buf, np.float32, 16, off); off += 64
    inl = np.frombuffer(buf, np.int32, 2 * int(n_inl), off).reshape(-1, 2)
    rec, inl_ref, _, _ = oracle.pose_ransac(ref[0], ref[1], m_ref, id_train=0, id_query=1, seed=2, params=p)
    assert int(found) == int(rec["found"]) and int(n_inl) == len(inl_ref)
    assert np.array_equal(inl[:, 0], inl_ref["queryIdx"]) and np.array_equal(inl[:, 1], inl_ref["trainIdx"])
    assert np.allclose(tf[:12], rec["tf"][:12], atol=1e-5) and abs(float(rmse) - float(rec["rmse"])) <= 1e-6 * max(1.0, float(rec["rmse"]))
