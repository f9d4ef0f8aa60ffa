"""GPU parity of the pair stage through the C ABI against the oracle.

Tier-E: line-match index pairs (+ distances), best RANSAC hypothesis index and its inlier set.
Tier-T (north star): final pose within 1e-5 rad / 1e-4 m, refined inlier set reported. The native LM on the
device follows the oracle's operation order, so these are also compared bit for bit first.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    """Rotation angle between two (float32-rounded, hence not exactly orthonormal) rotation matrices."""
    M = Ra.T @ Rb
    v = 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    return float(np.arctan2(np.linalg.norm(v), (np.trace(M) - 1) / 2))


def _pose_close(tf_a, tf_b):
    A, B = np.asarray(tf_a, np.float64).reshape(4, 4), np.asarray(tf_b, np.float64).reshape(4, 4)
    return rot_angle(A[:3, :3], B[:3, :3]) < 1e-5 and np.abs(A[:3, 3] - B[:3, 3]).max() < 1e-4


@pytest.fixture(scope="module")
def extracted(api, oracle, stream4):
    imgs, deps, poses, K = stream4
    ctx = api.Context(max_batch=4, max_w=640, max_h=480, debug=True)
    frames = ctx.extract_batch(imgs, deps, K, seeds=[1, 2, 3, 4])
    lines = [f.lines() for f in frames]
    yield ctx, frames, lines, poses
    ctx.close()


def test_lineMatching(api, oracle, extracted):
    ctx, frames, lines, poses = extracted
    for (q, t) in [(1, 0), (2, 1), (3, 0), (0, 3)]:
        for adjacent in (True, False):
            got = ctx.match_lines(frames[q], frames[t], adjacent)
            ref = oracle.lineMatching(lines[q], lines[t], adjacent)
            assert len(ref) > 20
            assert np.array_equal(got, ref), (q, t, adjacent)


def test_pose_ransac(api, oracle, extracted):
    ctx, frames, lines, poses = extracted
    from lineslam_b200 import synth
    for (q, t, seed) in [(1, 0, 1), (2, 1, 5), (3, 1, 9)]:
        m = oracle.lineMatching(lines[q], lines[t], True)
        rec_o, inl_o, rinl_o, tfr = oracle.pose_ransac(lines[t], lines[q], m, id_train=t, id_query=q, seed=seed)
        rec_g, inl_g, rinl_g = ctx.pose_ransac(frames[t], frames[q], m, id_train=t, id_query=q, seed=seed)
        assert rec_o["found"] == 1
        # Tier-E
        assert rec_g["best_iter"] == rec_o["best_iter"]
        assert np.array_equal(rinl_g, rinl_o)
        assert rec_g["n_ransac_inliers"] == rec_o["n_ransac_inliers"]
        # Tier-T
        assert _pose_close(rec_g["tf"], rec_o["tf"])
        assert rec_g["found"] == rec_o["found"]
        # ground truth sanity: the recovered motion is the synthetic one (cm level)
        T = synth.relative_pose_q2t(*poses[q], *poses[t])
        assert np.abs(rec_g["tf"].reshape(4, 4)[:3, 3] - T[:3, 3]).max() < 0.03
        # bit-level agreement of the refinement (same operation order on both sides)
        assert np.array_equal(inl_g, inl_o)
        assert rec_g["rmse"] == rec_o["rmse"]
        assert np.array_equal(rec_g["tf"], rec_o["tf"])


def test_match_pair_batch_and_node_api(api, oracle, extracted):
    ctx, frames, lines, poses = extracted
    qs, ts = [1, 2, 3, 3], [0, 1, 2, 0]
    seeds = [11, 12, 13, 14]
    recs = ctx.match_pair_batch([frames[q] for q in qs], [frames[t] for t in ts], qs, ts, seeds)
    for k, (q, t) in enumerate(zip(qs, ts)):
        adjacent = abs(q - t) <= 3
        m = oracle.lineMatching(lines[q], lines[t], adjacent)
        rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(lines[t], lines[q], m, id_train=t, id_query=q, seed=seeds[k])
        assert np.array_equal(ctx.pair_matches(k, 0), m)
        assert np.array_equal(ctx.pair_matches(k, 2), rinl_o)
        assert np.array_equal(ctx.pair_matches(k, 1), inl_o)
        for name in ("id_train", "id_query", "found", "n_line_matches", "n_ransac_inliers", "n_inliers", "best_iter"):
            assert recs[k][name] == rec_o[name], (k, name)
        assert _pose_close(recs[k]["tf"], rec_o["tf"])
    # reference-shaped interface
    n0 = api.Node(ctx, None, None, None, node_id=0, seed=1, frame=frames[0])
    n1 = api.Node(ctx, None, None, None, node_id=1, seed=1, frame=frames[1])
    mr = n1.matchNodePair(n0, seed=11)
    assert mr.found and mr.id1 == 0 and mr.id2 == 1
    assert np.array_equal(mr.final_trafo.ravel(), recs[0]["tf"])
    assert np.isclose(mr.informationMatrix[0, 0], len(mr.inlier_line_matches) / mr.rmse ** 2)


def test_too_few_matches_and_empty(api, oracle, extracted):
    ctx, frames, lines, poses = extracted
    m = oracle.lineMatching(lines[1], lines[0], True)[:12]          # < min_matches (20): motion.cpp:621-624
    rec_g, inl_g, rinl_g = ctx.pose_ransac(frames[0], frames[1], m)
    rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(lines[0], lines[1], m)
    assert rec_g["found"] == 0 == rec_o["found"] and rec_g["rmse"] == rec_o["rmse"] == np.float32(1e9)
    assert len(inl_g) == 0 and len(rinl_g) == 0
    empty = ctx.frame_from_lines(np.zeros(0, lines[0].dtype))
    assert len(ctx.match_lines(frames[0], empty, True)) == 0
    assert len(ctx.match_lines(empty, frames[0], True)) == 0
    recs = ctx.match_pair_batch([frames[1], empty], [empty, frames[0]], [1, 1], [0, 0], [1, 1])
    assert recs["found"].tolist() == [0, 0]
    # frames rebuilt from host records behave like extracted ones (cached features of the older frame)
    f0 = ctx.frame_from_lines(lines[0])
    assert np.array_equal(ctx.match_lines(frames[1], f0, True), oracle.lineMatching(lines[1], lines[0], True))


def test_loop_closure_ids(api, oracle, extracted):
    """|id difference| > 50 switches to min_matches_loopclose and the non-adjacent thresholds."""
    ctx, frames, lines, poses = extracted
    recs = ctx.match_pair_batch([frames[2]], [frames[0]], [100], [0], [3])
    m = oracle.lineMatching(lines[2], lines[0], False)
    rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(lines[0], lines[2], m, id_train=0, id_query=100, seed=3)
    assert np.array_equal(ctx.pair_matches(0, 0), m)
    assert recs[0]["found"] == rec_o["found"] and recs[0]["n_inliers"] == rec_o["n_inliers"]
    assert np.array_equal(ctx.pair_matches(0, 2), rinl_o)
    assert _pose_close(recs[0]["tf"], rec_o["tf"])
