"""BASELINE.json configs 4 and 5 as parity cases: the 1 x 256 loop-closure batch (line matching +
RANSAC on cached keyframe features, sharded block-wise across ranks) and a 1280x960 frame pair."""
import numpy as np
import pytest

from test_gpu_extract import _compare
from test_gpu_pair import _pose_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orbit8(api, oracle):
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(9, scene_seed=2000, traj=synth.trajectory_orbit, stride=4)
    K = synth.camera_K()
    ctx = api.Context(max_batch=9, max_w=640, max_h=480)
    frames = ctx.extract_batch(imgs, deps, K, seeds=list(range(1, 10)))
    lines = [f.lines() for f in frames]
    yield ctx, frames, lines
    ctx.close()


def test_loop_closure_batch_1x256(api, oracle, orbit8):
    """cfg 4: one query against 256 keyframes (8 distinct keyframes x 32 seeds; ids 100.. so that |id diff| > 50
    selects min_matches_loopclose and the non-adjacent matching thresholds). Every record is compared with the
    oracle; the two blocks a 2-rank shard would own reproduce the unsharded batch."""
    from lineslam_b200 import shard
    ctx, frames, lines = orbit8
    query, ql = frames[8], lines[8]
    trains = [frames[k % 8] for k in range(256)]
    ids_t = np.arange(256, dtype=np.int32)
    ids_q = np.full(256, 400, np.int32)
    seeds = np.arange(1, 257, dtype=np.uint32)
    recs = ctx.match_pair_batch([query] * 256, trains, ids_q, ids_t, seeds)
    assert recs["found"].sum() >= 32
    lm = [oracle.lineMatching(ql, lines[k], False) for k in range(8)]
    for k in range(256):
        rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(lines[k % 8], ql, lm[k % 8], id_train=int(ids_t[k]), id_query=400, seed=int(seeds[k]))
        assert recs[k]["n_line_matches"] == len(lm[k % 8])
        for name in ("found", "n_ransac_inliers", "n_inliers", "best_iter"):
            assert recs[k][name] == rec_o[name], (k, name)
        assert _pose_close(recs[k]["tf"], rec_o["tf"]) and recs[k]["rmse"] == rec_o["rmse"]
    assert np.array_equal(ctx.pair_matches(255, 0), lm[7])
    # sharded execution (world 2): each rank's block equals the same rows of the unsharded batch
    for rank in range(2):
        lo, hi, per = shard.shard_pairs(256, 2, rank)
        part = ctx.match_pair_batch([query] * (hi - lo), trains[lo:hi], ids_q[lo:hi], ids_t[lo:hi], seeds[lo:hi])
        assert part.tobytes() == recs[lo:hi].tobytes()


def test_ragged_batch_mixed_sizes(api, oracle, orbit8):
    """Pairs of very different sizes in one launch: full frames, a 5-line frame, an empty frame."""
    ctx, frames, lines = orbit8
    small = ctx.frame_from_lines(lines[1][:5])
    empty = ctx.frame_from_lines(lines[1][:0])
    qs = [frames[1], small, frames[2], empty, frames[3]]
    ts = [frames[0], frames[0], small, frames[1], frames[2]]
    ql = [lines[1], lines[1][:5], lines[2], lines[1][:0], lines[3]]
    tl = [lines[0], lines[0], lines[1][:5], lines[1], lines[2]]
    recs = ctx.match_pair_batch(qs, ts, [1, 1, 2, 3, 3], [0, 0, 1, 1, 2], [1, 2, 3, 4, 5])
    for k in range(5):
        m = oracle.lineMatching(ql[k], tl[k], True)
        assert np.array_equal(ctx.pair_matches(k, 0), m), k
        rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(tl[k], ql[k], m, id_train=[0, 0, 1, 1, 2][k], id_query=[1, 1, 2, 3, 3][k], seed=k + 1)
        assert recs[k]["found"] == rec_o["found"] and np.array_equal(ctx.pair_matches(k, 1), inl_o)
        assert np.array_equal(recs[k]["tf"], rec_o["tf"])


def test_1280x960_pair(api, oracle):
    """cfg 5 shape: 1280x960 frames (K = [1050, 1050, 639.5, 479.5], MSLD sub-region size 8)."""
    from lineslam_b200 import synth
    W, H = 1280, 960
    imgs, deps, poses = synth.make_stream(2, scene_seed=2002, W=W, H=H, traj=synth.trajectory_orbit, stride=2)
    K = synth.camera_K(W, H)
    ctx = api.Context(max_batch=2, max_w=W, max_h=H, debug=True)
    frames = ctx.extract_batch(imgs, deps, K, seeds=[5, 6])
    refs = []
    for i in range(2):
        ref, dref = oracle.detect3DLines(imgs[i], deps[i], K, seed=5 + i, debug=True)
        assert np.array_equal(frames[i].segments(), dref["segs"])       # LSD end points, Tier-E
        _compare(frames[i].lines(), ref, frames[i].debug(), dref, f"xga{i}")
        refs.append(ref)
    assert len(refs[0]) > 80
    recs = ctx.match_pair_batch([frames[1]], [frames[0]], [1], [0], [9])
    m = oracle.lineMatching(refs[1], refs[0], True)
    rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(refs[0], refs[1], m, id_train=0, id_query=1, seed=9)
    assert np.array_equal(ctx.pair_matches(0, 0), m) and np.array_equal(ctx.pair_matches(0, 2), rinl_o)
    assert recs[0]["found"] == rec_o["found"] == 1 and _pose_close(recs[0]["tf"], rec_o["tf"])
    T = synth.relative_pose_q2t(*poses[1], *poses[0])
    assert np.abs(recs[0]["tf"].reshape(4, 4)[:3, 3] - T[:3, 3]).max() < 0.03
    ctx.close()


@pytest.mark.parametrize("W,H", [(324, 242), (332, 250), (200, 152)])
def test_odd_sizes(api, oracle, W, H):
    """Sizes whose 0.8x scaled width is odd or not a multiple of 32/128 (partial tiles everywhere, the y pass falls
    back from the TMA kernel when the scaled width is odd): every stage and the final records stay bit-exact."""
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(2, scene_seed=2003, W=W, H=H)
    K = synth.camera_K(W, H)
    p = api.default_params()
    p.min_feature_matches = 10
    ctx = api.Context(params=p, max_batch=2, max_w=W, max_h=H, debug=True)
    frames = ctx.extract_batch(imgs, deps, K, seeds=[3, 4])
    sw, sh = int(np.floor(W * 0.8)), int(np.floor(H * 0.8))
    segs, dbg = oracle.lsd(oracle.gray(imgs[0]), debug=True)
    assert np.array_equal(ctx.debug_read(1, np.float64, sw * sh).reshape(sh, sw), dbg["scaled"])
    assert np.array_equal(ctx.debug_read(2, np.float64, sw * sh).reshape(sh, sw), dbg["angles"])
    assert np.array_equal(ctx.debug_read(4, np.int32, sw * sh), dbg["seeds"])
    for i in range(2):
        ref, dref = oracle.detect3DLines(imgs[i], deps[i], K, seed=3 + i, params=p, debug=True)
        assert np.array_equal(frames[i].segments(), dref["segs"])
        _compare(frames[i].lines(), ref, frames[i].debug(), dref, f"odd{W}x{H}_{i}")
    ctx.close()
