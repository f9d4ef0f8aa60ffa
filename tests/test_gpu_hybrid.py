"""GPU parity of the point + line pair stage through the C ABI.

Tier-E: point-match index pairs and distances (Node::featureMatching), best RANSAC hypothesis, point and line
inlier sets of the best hypothesis. Tier-T (north star 1e-5 rad / 1e-4 m): refined pose; the device follows the
oracle's operation order, so the refined quantities are also compared bit for bit.
"""
import numpy as np
import pytest

from test_gpu_pair import _pose_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hyb(api, oracle):
    from lineslam_b200 import synth
    n = 3
    imgs, deps, poses = synth.make_stream(n, scene_seed=2000, traj=synth.trajectory_orbit, stride=3)
    K = synth.camera_K()
    ctx = api.Context(max_batch=n, max_w=640, max_h=480, debug=True)
    frames = ctx.extract_batch(imgs, deps, K, seeds=[1, 2, 3])
    lines = [f.lines() for f in frames]
    pts = [synth.make_points(2000, 3 * i, deps[i], *poses[i], K) for i in range(n)]
    # a few features without depth (NaN z) as the reference keeps them (src/misc.cpp:716-722)
    for x, d, _ in pts:
        x[5::37, 2] = np.nan
    for f, (x, d, _) in zip(frames, pts):
        f.set_points(x, d)
    yield ctx, frames, lines, pts, poses
    ctx.close()


def test_featureMatching(api, oracle, hyb):
    ctx, frames, lines, pts, poses = hyb
    for (q, t, seed) in [(1, 0, 1), (2, 1, 7), (0, 2, 3)]:
        got = ctx.match_points(frames[q], frames[t], seed)
        ref = oracle.featureMatching(pts[q][1], pts[t][1], ctx.params.nn_distance_ratio, seed)
        assert len(ref) > 100
        assert np.array_equal(got, ref), (q, t)
    # ragged / degenerate: fewer than two train rows -> no matches (knnMatch k = 2)
    one = ctx.frame_from_lines(lines[0][:0]).set_points(pts[0][0][:1], pts[0][1][:1])
    assert len(ctx.match_points(frames[1], one, 1)) == 0
    assert len(ctx.match_points(one, frames[1], 1)) == len(oracle.featureMatching(pts[0][1][:1], pts[1][1], 0.5, 1))


def _check_hybrid(ctx, rec_g, inl_g, rinl_g, out):
    rec_o = out["rec"]
    assert rec_g["best_iter"] == rec_o["best_iter"]
    assert np.array_equal(rinl_g, out["ln_ransac_inliers"])                  # Tier-E
    assert np.array_equal(ctx.pair_matches(0, 5), out["pt_ransac_inliers"])  # Tier-E
    assert _pose_close(rec_g["tf"], rec_o["tf"])                             # Tier-T
    assert rec_g["found"] == rec_o["found"]
    assert np.array_equal(inl_g, out["ln_inliers"])
    assert np.array_equal(ctx.pair_matches(0, 4), out["pt_inliers"])
    assert list(rec_g["pad"][:3]) == list(rec_o["pad"][:3])
    assert rec_g["rmse"] == rec_o["rmse"]
    assert np.array_equal(rec_g["tf"], rec_o["tf"])


def test_pose_ransac_hybrid(api, oracle, hyb):
    from lineslam_b200 import synth
    ctx, frames, lines, pts, poses = hyb
    for (q, t, seed) in [(1, 0, 1), (2, 1, 5), (2, 0, 9)]:
        pm = oracle.featureMatching(pts[q][1], pts[t][1], 0.5, seed)
        lm = oracle.lineMatching(lines[q], lines[t], True)
        out = oracle.pose_ransac_hybrid(lines[t], lines[q], pts[t][0], pts[q][0], pm, lm, id_train=t, id_query=q, seed=seed,
                                        skip_draws=len(pm))
        rec_g, inl_g, rinl_g = ctx.pose_ransac(frames[t], frames[q], lm, id_train=t, id_query=q, seed=seed, pt_matches=pm)
        assert out["rec"]["found"] == 1 and len(out["pt_inliers"]) > 50
        _check_hybrid(ctx, rec_g, inl_g, rinl_g, out)
        T = synth.relative_pose_q2t(*poses[q], *poses[t])
        assert np.abs(rec_g["tf"].reshape(4, 4)[:3, 3] - T[:3, 3]).max() < 0.03
    # points only (no line matches): every sample goes through the Kabsch solver
    pm = oracle.featureMatching(pts[1][1], pts[0][1], 0.5, 2)
    out = oracle.pose_ransac_hybrid(lines[0], lines[1], pts[0][0], pts[1][0], pm, lines[0][:0].view(np.uint8)[:0].view(pm.dtype),
                                    seed=2, skip_draws=len(pm))
    rec_g, inl_g, rinl_g = ctx.pose_ransac(frames[0], frames[1], None, seed=2, pt_matches=pm)
    _check_hybrid(ctx, rec_g, inl_g, rinl_g, out)


def test_matchNodePair_hybrid_batch(api, oracle, hyb):
    ctx, frames, lines, pts, poses = hyb
    qs, ts, seeds = [1, 2, 2], [0, 1, 0], [21, 22, 23]
    recs = ctx.match_pair_batch([frames[q] for q in qs], [frames[t] for t in ts], qs, ts, seeds)
    for k, (q, t) in enumerate(zip(qs, ts)):
        pm = oracle.featureMatching(pts[q][1], pts[t][1], 0.5, seeds[k])
        lm = oracle.lineMatching(lines[q], lines[t], True)
        out = oracle.pose_ransac_hybrid(lines[t], lines[q], pts[t][0], pts[q][0], pm, lm, id_train=t, id_query=q,
                                        seed=seeds[k], skip_draws=len(pm))
        assert np.array_equal(ctx.pair_matches(k, 3), pm)
        assert np.array_equal(ctx.pair_matches(k, 0), lm)
        assert np.array_equal(ctx.pair_matches(k, 5), out["pt_ransac_inliers"])
        assert np.array_equal(ctx.pair_matches(k, 2), out["ln_ransac_inliers"])
        assert np.array_equal(ctx.pair_matches(k, 4), out["pt_inliers"])
        assert np.array_equal(ctx.pair_matches(k, 1), out["ln_inliers"])
        for name in ("found", "n_line_matches", "n_ransac_inliers", "n_inliers", "best_iter"):
            assert recs[k][name] == out["rec"][name], (k, name)
        assert _pose_close(recs[k]["tf"], out["rec"]["tf"])
    n0 = api.Node(ctx, None, None, None, node_id=0, seed=1, frame=frames[0])
    n1 = api.Node(ctx, None, None, None, node_id=1, seed=1, frame=frames[1])
    mr = n1.matchNodePair(n0, seed=21)
    assert mr.found and len(mr.inlier_matches) > 50 and len(mr.inlier_line_matches) > 10
    assert np.array_equal(mr.final_trafo.ravel(), recs[0]["tf"])
    w = len(mr.inlier_matches) + len(mr.inlier_line_matches)
    assert np.isclose(mr.informationMatrix[0, 0], w / mr.rmse ** 2)


def test_hybrid_kernel_without_point_matches_equals_line_kernel(api, oracle, hyb):
    """Frames that carry points whose descriptors match nothing: the point + line kernel must reproduce the
    line-only path exactly (same rand() stream, same refinement)."""
    ctx, frames, lines, pts, poses = hyb
    rng = np.random.default_rng(0)
    junk = [ctx.frame_from_lines(lines[i]).set_points(pts[i][0][:40], np.abs(rng.normal(size=(40, 128))).astype(np.float32) + 5)
            for i in range(2)]
    recs = ctx.match_pair_batch([junk[1]], [junk[0]], [1], [0], [5])
    assert len(ctx.pair_matches(0, 3)) == 0
    lm = oracle.lineMatching(lines[1], lines[0], True)
    rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(lines[0], lines[1], lm, id_train=0, id_query=1, seed=5)
    assert np.array_equal(ctx.pair_matches(0, 2), rinl_o) and np.array_equal(ctx.pair_matches(0, 1), inl_o)
    assert np.array_equal(recs[0]["tf"], rec_o["tf"]) and recs[0]["rmse"] == rec_o["rmse"]


def test_relmotion_ransac_levmar(api, oracle, hyb):
    """computeRelativeMotion_Ransac + optimizeRelmotion (src/line/motion.cpp:367-526, 98-139):
    consensus index set bit-exact, refined R, t bit-exact (same levmar operation order), and within the
    north-star tolerance as the formal bar."""
    ctx, frames, lines, pts, poses = hyb
    for (q, t, seed) in [(1, 0, 3), (2, 1, 4), (2, 0, 8)]:
        lm = oracle.lineMatching(lines[q], lines[t], True)
        ref = oracle.relmotion_ransac(lines[q][lm["queryIdx"]], lines[t][lm["trainIdx"]], seed=seed)
        got = ctx.relmotion_ransac(frames[t], frames[q], lm, seed=seed)
        assert ref["have"] and ref["lm_calls"] >= 1 and len(ref["conset"]) > 20
        assert got["have"] == ref["have"] and got["lm_calls"] == ref["lm_calls"]
        assert np.array_equal(got["conset"], ref["conset"])
        assert np.abs(got["R"] - ref["R"]).max() < 1e-5 and np.abs(got["t"] - ref["t"]).max() < 1e-4
        assert np.array_equal(got["R"], ref["R"]) and np.array_equal(got["t"], ref["t"])
    # fewer than three pairs -> empty consensus; three pairs -> no refinement (motion.cpp:370-374, 476-478)
    lm = oracle.lineMatching(lines[1], lines[0], True)
    assert len(ctx.relmotion_ransac(frames[0], frames[1], lm[:2])["conset"]) == 0
    ref = oracle.relmotion_ransac(lines[1][lm["queryIdx"][:3]], lines[0][lm["trainIdx"][:3]], seed=1)
    got = ctx.relmotion_ransac(frames[0], frames[1], lm[:3], seed=1)
    assert np.array_equal(got["conset"], ref["conset"]) and got["lm_calls"] == ref["lm_calls"] == 0


def test_orb_hamming_branch_and_device_rootsift(api, oracle, hyb):
    """a21 remainder: ORB rows through the BruteForce-HammingLUT matcher, and squareroot_descriptor_space applied on
    the device (src/node.cpp:606-641, 304-310, 1823-1837)."""
    ctx, frames, lines, pts, poses = hyb
    rng = np.random.default_rng(31)
    q = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (570, 32), dtype=np.uint8)
    sel = rng.permutation(600)[:400]
    t[:400] = q[sel]
    t[:400] ^= ((rng.random((400, 32)) < 0.04) * rng.integers(0, 256, (400, 32))).astype(np.uint8)
    t[9] = t[8]
    xq = np.ones((600, 4), np.float32); xt = np.ones((570, 4), np.float32)
    fq = ctx.frame_from_lines(lines[0][:0]).set_points(xq, q)
    ft = ctx.frame_from_lines(lines[0][:0]).set_points(xt, t)
    for seed in (1, 5):
        got = ctx.match_points(fq, ft, seed)
        ref = oracle.featureMatching_hamming(q, t, ctx.params.nn_distance_ratio, seed)
        assert len(ref) > 200 and np.array_equal(got, ref)
    assert np.array_equal(fq.descriptors(), q)
    # mixing ORB and float rows in one pair is refused, not silently matched
    with pytest.raises(Exception):
        ctx.match_points(fq, frames[0], 1)
    # RootSIFT on the device == the oracle's conditioning of the raw rows, and the matches that follow are identical
    raw = [(rng.normal(size=(500, 128)) * rng.choice([0.1, 1.0, 30.0], size=(500, 1))).astype(np.float32) for _ in range(2)]
    raw[1][:300] = raw[0][rng.permutation(500)[:300]] + rng.normal(0, 0.01, (300, 128)).astype(np.float32)
    raw[0][3] = 0
    x = np.ones((500, 4), np.float32)
    fa = ctx.frame_from_lines(lines[0][:0]).set_points(x, raw[0], root_sift=True)
    fb = ctx.frame_from_lines(lines[0][:0]).set_points(x, raw[1], root_sift=True)
    ra, rb = oracle.rootsift(raw[0]), oracle.rootsift(raw[1])
    assert np.array_equal(fa.descriptors(), ra) and np.array_equal(fb.descriptors(), rb)
    got = ctx.match_points(fa, fb, 2)
    ref = oracle.featureMatching(ra, rb, ctx.params.nn_distance_ratio, 2)
    assert len(ref) > 100 and np.array_equal(got, ref)
    for f in (fq, ft, fa, fb):
        f.free()


def test_tensor_core_prefilter_gives_the_exact_matches(api, oracle):
    """N2 (north star: tensor cores for the descriptor-distance matrix): featureMatching runs G = Q T^T on tcgen05 (tf32, TMEM
    accumulator) as a pre-filter and re-evaluates the surviving candidates in OpenCV's f32 summation order — the match lists
    (indices, ratio + jitter distances) must equal the oracle's exhaustive scan, including on data built to sit inside the
    tf32 error band: near-duplicate rows, exact ties, big norms, zero rows, ragged sizes."""
    p = api.default_params()
    p.nn_distance_ratio = 0.97          # accept nearly everything: exposes any error in the SECOND nearest distance too
    ctx = api.Context(params=p, max_batch=1, max_w=64, max_h=64)
    empty = np.zeros(0, api.LINE_DTYPE) if hasattr(api, "LINE_DTYPE") else None
    rng = np.random.default_rng(77)

    def frames(q, t):
        from lineslam_b200.records import LINE_DTYPE
        fq = ctx.frame_from_lines(np.zeros(0, LINE_DTYPE)).set_points(np.ones((len(q), 4), np.float32), q)
        ft = ctx.frame_from_lines(np.zeros(0, LINE_DTYPE)).set_points(np.ones((len(t), 4), np.float32), t)
        return fq, ft

    def unit(x):
        x = np.abs(x).astype(np.float32)
        return np.sqrt(x / x.sum(1, keepdims=True)).astype(np.float32)

    cases = []
    for dim in (128, 64, 72, 8):
        q = unit(rng.random((600, dim)) ** 3); t = unit(rng.random((600, dim)) ** 3)
        t[:350] = unit(q[rng.permutation(600)[:350]] ** 2 + rng.normal(0, 0.003, (350, dim)) ** 2)
        cases.append((f"rootsift{dim}", q, t))
    # near duplicates far inside the tf32 band: 1e-6 .. 1e-4 perturbations of the same row, plus exact ties
    base = unit(rng.random((1, 128)))
    t = np.repeat(base, 300, 0) + (rng.normal(0, 1, (300, 128)) * np.logspace(-7, -4, 300)[:, None]).astype(np.float32)
    t = np.abs(t).astype(np.float32); t[17] = t[4]; t[250] = t[3]
    q = np.concatenate([base + np.float32(1e-5), unit(rng.random((130, 128)))]).astype(np.float32)
    cases.append(("near_duplicates", q, t))
    # big norms, mixed scales, zero rows, ragged sizes
    q = (rng.normal(0, 1, (129, 128)) * rng.choice([0.01, 1.0, 100.0], (129, 1))).astype(np.float32); q[5] = 0
    t = (rng.normal(0, 1, (131, 128)) * rng.choice([0.01, 1.0, 100.0], (131, 1))).astype(np.float32); t[7] = 0
    t[:60] = q[:60] * np.float32(1.001)
    cases.append(("mixed_scales", q, t))
    cases.append(("tiny", unit(rng.random((1, 64))), unit(rng.random((2, 64)))))
    cases.append(("one_train_row", unit(rng.random((5, 64))), unit(rng.random((1, 64)))))
    ctx.match_tc_stats(reset=True)
    total_pairs = 0
    for name, q, t in cases:
        fq, ft = frames(q, t)
        for seed in (1, 9):
            got = ctx.match_points(fq, ft, seed)
            ref = oracle.featureMatching(q, t, p.nn_distance_ratio, seed)
            assert len(got) == len(ref), (name, len(got), len(ref))
            assert np.array_equal(got, ref), name
        total_pairs += 2 * len(q) * len(t)
        fq.free(); ft.free()
    evals, full_rows, rows = ctx.match_tc_stats()
    assert rows > 0 and evals < 0.5 * total_pairs        # the pre-filter removed most exact evaluations even on this adversarial mix
    ctx.close()


def test_batched_entry_points_equal_the_per_call_ones(api, oracle, hyb):
    """lsl_frames_set_points_batch == per-frame lsl_frame_set_points_ex, lsl_relmotion_batch == per-pair lsl_relmotion_ransac
    (which test_relmotion_ransac_levmar pins to the oracle): same bits, one launch / one device block instead of one per call."""
    ctx, frames, lines, pts, poses = hyb
    from lineslam_b200.records import LINE_DTYPE
    fb = [ctx.frame_from_lines(lines[i]) for i in range(3)]
    ctx.set_points_batch(fb, [p[0] for p in pts], [p[1] for p in pts])
    for (q, t, seed) in [(1, 0, 1), (2, 1, 7)]:
        assert np.array_equal(ctx.match_points(fb[q], fb[t], seed), ctx.match_points(frames[q], frames[t], seed))
    assert np.array_equal(fb[2].descriptors(), pts[2][1])
    # relmotion: batch of two pairs vs the single-pair entry on the same line matches and seeds
    recs = ctx.match_pair_batch([frames[1], frames[2]], [frames[0], frames[1]], [1, 2], [0, 1], [5, 6])
    ms = [ctx.pair_matches(k, 0) for k in range(2)]
    R, t, info = ctx.relmotion_batch(2)
    for k, (q, tr) in enumerate([(1, 0), (2, 1)]):
        one = ctx.relmotion_ransac(frames[tr], frames[q], ms[k], seed=5 + k)
        assert bool(info[k, 2]) == one["have"] and int(info[k, 0]) == len(one["conset"]) and int(info[k, 1]) == one["lm_calls"]
        if one["have"]:
            assert np.array_equal(R[k], one["R"]) and np.array_equal(t[k], one["t"])
    for f in fb:
        f.free()


def test_computeInliersAndError(api, oracle, hyb):
    """Node::computeInliersAndError (src/node.cpp:1019-1080): inlier list in match order and the rms Mahalanobis distance,
    with zero-depth points (skipped), NaN-depth points (errorFunction2 = max: outliers), wrong matches and a loose / tight gate."""
    ctx, frames, lines, pts, poses = hyb
    from lineslam_b200 import synth
    rng = np.random.default_rng(5)
    xq, xt = pts[1][0].copy(), pts[0][0].copy()
    xq[3::41, 2] = 0.0; xt[7::53, 2] = 0.0                     # zero depth: the match is skipped before the distance
    fq = ctx.frame_from_lines(lines[1][:0]); ft = ctx.frame_from_lines(lines[0][:0])
    fq.set_points(xq, pts[1][1]); ft.set_points(xt, pts[0][1])
    pm = ctx.match_points(frames[1], frames[0], seed=3)
    m = np.concatenate([pm, pm[:40]])
    m["trainIdx"][-40:] = rng.integers(0, len(xt), 40)        # wrong correspondences
    T = synth.relative_pose_q2t(*poses[1], *poses[0]).astype(np.float32)
    for gate in (9.0, 0.5, 1e-6):
        got, rmse = ctx.compute_inliers_and_error(fq, ft, m, T, gate)
        keep, rmse_o = oracle.compute_inliers_and_error(m, xq, xt, T, gate)
        assert np.array_equal(got, m[keep]) and rmse == rmse_o, (gate, len(got), len(keep), rmse, rmse_o)
    assert len(keep) < 3 and rmse == 1e9                        # nothing passes the tightest gate
    got, rmse = ctx.compute_inliers_and_error(fq, ft, m, T, 9.0)
    assert 20 < len(got) < len(m)
    bad = m[:3].copy(); bad["queryIdx"][1] = len(xq)
    with pytest.raises(api.LslError):
        ctx.compute_inliers_and_error(fq, ft, bad, T, 9.0)
    fq.free(); ft.free()
