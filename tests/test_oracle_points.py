"""CPU checks of the point / hybrid restatements in oracle/ (SURVEY.md §8 rows a21, a22, a24-a26).

OpenCV 2.4, Eigen, PCL and g2o are not in /root/reference, so these stages are restated from the published
algorithms (parity with the real libraries UNPINNED); here they are pinned against independent numpy
formulations and against the synthetic ground truth.
"""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def hybrid_pair(oracle):
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(2, scene_seed=2000, traj=synth.trajectory_orbit, stride=3)
    K = synth.camera_K()
    lines = [oracle.detect3DLines(imgs[i], deps[i], K, seed=i + 1) for i in range(2)]
    pts = [synth.make_points(2000, 3 * i, deps[i], *poses[i], K) for i in range(2)]
    return lines, pts, poses


def test_ldlt_and_error_function2(oracle):
    rng = np.random.default_rng(3)
    for _ in range(50):
        M = rng.normal(size=(3, 3)); A = M @ M.T + np.diag(rng.uniform(1e-6, 1e-2, 3))
        b = rng.normal(size=3)
        assert np.allclose(oracle.ldlt3_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    # errorFunction2 against the textbook formula (src/misc.cpp:699-786)
    c = 3 * np.tan(58.0 / 180 * np.pi / 640); rcx = c * c
    c = 3 * np.tan(45.0 / 180 * np.pi / 480); rcy = c * c
    for k in range(50):
        ang = rng.uniform(-0.2, 0.2, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
        T = np.eye(4, dtype=np.float32); T[:3, :3] = (Rx @ Rz).astype(np.float32); T[:3, 3] = rng.uniform(-0.1, 0.1, 3)
        x1 = np.array([*rng.uniform(-1, 1, 2), rng.uniform(0.8, 4), 1], np.float32)
        x2 = (T.astype(np.float64) @ x1.astype(np.float64)).astype(np.float32)
        x2[:3] += rng.normal(0, 0.004, 3).astype(np.float32)
        got = oracle.error_function2(x1, x2, T)
        Td = T.astype(np.float64); a = x1.astype(np.float64); b = x2.astype(np.float64)
        dmu = (Td @ a)[:3] - b[:3]
        dc = lambda z: (0.01 * z * z) ** 2
        if dmu @ dmu > 2 * (max(rcx, dc(a[2])) + max(rcx, dc(b[2]))):
            assert got == np.finfo(np.float64).max
            continue
        S = Td[:3, :3].T @ np.diag([rcx * a[2], rcy * a[2], dc(a[2])]) @ Td[:3, :3] + np.diag([rcx * b[2], rcy * b[2], dc(b[2])])
        assert np.isclose(got, dmu @ np.linalg.solve(S, dmu), rtol=1e-9)
    assert oracle.error_function2([0, 0, np.nan, 1], [0, 0, 1, 1], np.eye(4)) == np.finfo(np.float64).max


def test_kabsch_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    for n in (3, 3, 3, 5, 20):
        ang = rng.uniform(-0.5, 0.5)
        R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
        t = rng.uniform(-0.3, 0.3, 3)
        A = rng.uniform(-1, 1, (n, 3)) + [0, 0, 2.5]
        B = A @ R.T + t
        w = (1 / (np.abs(A[:, 2]) + np.abs(B[:, 2]))).astype(np.float32)
        tf = oracle.kabsch(A, B, w)
        assert np.allclose(tf[:3, :3], R, atol=2e-5) and np.allclose(tf[:3, 3], t, atol=5e-5)
        assert np.isclose(np.linalg.det(tf[:3, :3].astype(np.float64)), 1.0, atol=1e-5)


def test_featureMatching_bruteforce(oracle, hybrid_pair):
    lines, pts, poses = hybrid_pair
    (x0, d0, id0), (x1, d1, id1) = pts
    m = oracle.featureMatching(d1, d0, nn_ratio=0.5, seed=1)
    assert len(m) > 150
    # independent knn in float64: the accepted pairs are true landmark correspondences, train indices unique
    D = np.sqrt(((d1[:, None, :].astype(np.float64) - d0[None].astype(np.float64)) ** 2).sum(-1))
    nn = D.argmin(1)
    assert np.array_equal(m["trainIdx"], nn[m["queryIdx"]])
    assert len(np.unique(m["trainIdx"])) == len(m)
    assert (id1[m["queryIdx"]] == id0[m["trainIdx"]]).mean() > 0.98
    srt = np.sort(D, 1)
    ratio = (srt[:, 0] / srt[:, 1])[m["queryIdx"]]
    jit = oracle.rand(1, len(m)).astype(np.float64) / (1000.0 * 2147483647)
    assert np.allclose(m["distance"], ratio + jit, atol=1e-6)
    assert np.array_equal(np.sort(m["queryIdx"]), m["queryIdx"])
    # rootsift: rows non-negative with unit L2 norm
    r = oracle.rootsift(np.random.default_rng(0).normal(size=(5, 64)).astype(np.float32))
    assert (r >= 0).all() and np.allclose((r.astype(np.float64) ** 2).sum(1), 1, atol=1e-5)


def test_hybrid_ransac_recovers_ground_truth(oracle, hybrid_pair):
    from lineslam_b200 import synth
    lines, pts, poses = hybrid_pair
    (x0, d0, _), (x1, d1, _) = pts
    pm = oracle.featureMatching(d1, d0, 0.5, seed=9)
    lm = oracle.lineMatching(lines[1], lines[0], True)
    T = synth.relative_pose_q2t(*poses[1], *poses[0])
    out = oracle.pose_ransac_hybrid(lines[0], lines[1], x0, x1, pm, lm, id_train=0, id_query=1, seed=9, skip_draws=len(pm))
    rec = out["rec"]
    assert rec["found"] == 1 and rec["pad"][0] == len(pm)
    tf = rec["tf"].reshape(4, 4)
    assert np.abs(tf[:3, 3] - T[:3, 3]).max() < 0.02 and np.abs(tf[:3, :3] - T[:3, :3]).max() < 0.01
    assert len(out["pt_inliers"]) > 0.5 * len(pm) and len(out["ln_inliers"]) > 0.4 * len(lm)
    # points only / lines only go through the same function
    only_p = oracle.pose_ransac_hybrid(lines[0], lines[1], x0, x1, pm, lm[:0], seed=4)
    assert only_p["rec"]["found"] == 1 and np.abs(only_p["rec"]["tf"].reshape(4, 4)[:3, 3] - T[:3, 3]).max() < 0.03
    only_l = oracle.pose_ransac_hybrid(lines[0], lines[1], x0[:0], x1[:0], pm[:0], lm, seed=4)
    rec_l, inl_l, rinl_l, _ = oracle.pose_ransac(lines[0], lines[1], lm, seed=4)
    assert np.array_equal(only_l["rec"]["tf"], rec_l["tf"]) and np.array_equal(only_l["ln_inliers"], inl_l)


def test_relmotion_ransac_oracle(oracle, hybrid_pair):
    """Row a27 on the CPU: the line-only RANSAC + levmar refinement recovers the synthetic motion."""
    from lineslam_b200 import synth
    lines, pts, poses = hybrid_pair
    lm = oracle.lineMatching(lines[1], lines[0], True)
    r = oracle.relmotion_ransac(lines[1][lm["queryIdx"]], lines[0][lm["trainIdx"]], seed=3)
    T = synth.relative_pose_q2t(*poses[1], *poses[0])
    assert r["have"] and len(r["conset"]) > 0.5 * len(lm) and r["lm_calls"] >= 1
    assert np.abs(r["R"] - T[:3, :3]).max() < 0.01 and np.abs(r["t"] - T[:3, 3]).max() < 0.03
    assert abs(np.linalg.det(r["R"]) - 1) < 1e-9
    xs = np.linspace(-1, 1, 41)
    assert max(abs(oracle.lib().orc_m_acos(x) - np.arccos(x)) for x in xs) < 1e-15
