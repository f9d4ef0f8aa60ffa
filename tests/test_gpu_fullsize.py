"""Size-independent properties at the bench's full batch (592 VGA frames, BASELINE config 2): a frame's records do not
depend on its position in the batch or on the batch size, the host-buffer path equals the small-batch path, and a
pair's 128-byte record depends only on the two frames, the ids and the seed — checked bit for bit over the whole batch
(the oracle comparison of the same frames at small sizes is tests/test_gpu_extract.py / test_gpu_pair.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batch_592_position_and_size_invariance(api):
    from lineslam_b200 import synth
    U, B = 6, 592
    imgs, deps, poses = synth.make_stream(U, scene_seed=2000)
    K = synth.camera_K()
    small = api.Context(max_batch=U, max_w=640, max_h=480)
    ref = small.extract_batch(imgs, deps, K, seeds=list(range(1, U + 1)))
    ref_lines = [f.lines().tobytes() for f in ref]
    assert all(f.num_lines > 50 for f in ref)
    order = np.array([k % U for k in range(B)])
    big = api.Context(max_batch=B, max_w=640, max_h=480)
    frames = big.extract_batch(imgs[order], deps[order], K, seeds=[int(o) + 1 for o in order])
    assert len(frames) == B
    for k in range(B):
        assert frames[k].lines().tobytes() == ref_lines[order[k]], k
    # pairs (k, k-1): the record depends only on (frame k % U, frame (k-1) % U, ids, seed)
    q, t = frames[1:], frames[:-1]
    idq = [int(order[k]) for k in range(1, B)]
    idt = [int(order[k - 1]) for k in range(1, B)]
    seeds = [7 + idq[k] for k in range(B - 1)]
    recs = big.match_pair_batch(q, t, idq, idt, seeds)
    assert int(recs["found"].sum()) >= (B - 1) * (U - 1) // U           # all but the wrap-around pairs register
    first = {}
    for k in range(B - 1):
        key = (idq[k], idt[k])
        if key not in first:
            first[key] = recs[k].tobytes()
        assert recs[k].tobytes() == first[key], k
    assert len(first) == U
    for (a, b), blob in first.items():                                   # and equals a single-pair call on the small context
        one = small.match_pair_batch([ref[a]], [ref[b]], [a], [b], [7 + a])[0]
        assert one.tobytes() == blob, (a, b)
    big.close()
    small.close()


def test_forty_distinct_stream_frames_against_the_oracle(api, oracle):
    """The bench's workload itself against the oracle: 40 DISTINCT consecutive frames of the fr1/xyz-shape stream (the frames
    bench.py renders, with the 16-bit TUM depth conversion it uses) — every line record of every frame, and every odometry
    pair (k, k-1): line-match lists, inlier sets of the best RANSAC hypothesis, refined inlier sets, the 128-byte record."""
    import multiprocessing as mp
    import bench
    U = 40
    imgs, deps, K = bench.make_unique_frames(U, 0)
    raw16, deps = bench.tum_depth_planes(deps)
    p = api.default_params()
    ctx = api.Context(params=p, max_batch=U, max_w=640, max_h=480)
    seeds = np.arange(1, U + 1, dtype=np.uint32)
    frames = ctx.extract_batch(imgs, raw16, K, seeds)                 # the e2e entry: 16-bit depth converted on the device

    def ref_lines(i):
        return oracle.detect3DLines(imgs[i], deps[i], K, seed=int(seeds[i]), params=p)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:                                  # the oracle releases the GIL inside ctypes
        ref = list(ex.map(ref_lines, range(U)))
    counts = []
    for i in range(U):
        got = frames[i].lines()
        assert len(got) == len(ref[i]) > 50, i
        assert got.tobytes() == ref[i].tobytes() or np.array_equal(got["des"], ref[i]["des"], equal_nan=True), i
        for name in ("p", "q", "r", "A", "B", "covA", "covB", "DU_A", "DU_B"):
            assert np.array_equal(got[name], ref[i][name]), (i, name)
        counts.append(len(got))
    assert max(counts) - min(counts) >= 10                             # the frames really differ
    ids = np.arange(1, U, dtype=np.int32)
    recs = ctx.match_pair_batch(frames[1:], frames[:-1], ids, ids - 1, seeds[1:])
    for k in range(U - 1):
        m = oracle.lineMatching(ref[k + 1], ref[k], True)
        assert np.array_equal(ctx.pair_matches(k, 0), m), k
        rec_o, inl_o, rinl_o, _ = oracle.pose_ransac(ref[k], ref[k + 1], m, id_train=k, id_query=k + 1, seed=int(seeds[k + 1]), params=p)
        assert np.array_equal(ctx.pair_matches(k, 2), rinl_o), k       # Tier-E: inliers of the best hypothesis
        assert np.array_equal(ctx.pair_matches(k, 1), inl_o), k
        assert recs[k]["found"] == rec_o["found"] and recs[k]["best_iter"] == rec_o["best_iter"], k
        assert np.allclose(recs[k]["tf"], rec_o["tf"], atol=1e-5, rtol=0), k
        assert np.array_equal(recs[k]["tf"], rec_o["tf"]) and recs[k]["rmse"] == rec_o["rmse"], k
    ctx.close()
