"""Boundary contract (include/lsl.h "Threading"): two host threads on ONE context — the reference's
QtConcurrent worker running detect3DLines while pool threads run matchNodePair (src/node.cpp:214,
src/graph_manager.cpp:555) — must give the results of the serial run bit for bit."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_extract_and_pair_calls_from_two_threads(api, stream4):
    imgs, deps, poses, K = stream4
    p = api.default_params()
    p.min_feature_matches = 10
    ctx = api.Context(params=p, max_batch=2, max_w=640, max_h=480)
    f01 = ctx.extract_batch(imgs[:2], deps[:2], K, seeds=[1, 2])
    # serial reference run
    serial_frames = ctx.extract_batch(imgs[2:4], deps[2:4], K, seeds=[3, 4])
    serial_lines = [f.lines().copy() for f in serial_frames]
    serial_rec = ctx.match_pair_batch([f01[1]], [f01[0]], [1], [0], [7]).copy()
    serial_m = ctx.pair_matches(0, 0).copy()
    for f in serial_frames:
        f.free()

    out, errs = {"lines": [], "recs": [], "m": []}, []

    def extractor():
        try:
            for _ in range(6):
                fr = ctx.extract_batch(imgs[2:4], deps[2:4], K, seeds=[3, 4])
                out["lines"].append([f.lines().copy() for f in fr])
                for f in fr:
                    f.free()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    def matcher():
        try:
            for _ in range(40):
                out["recs"].append(ctx.match_pair_batch([f01[1]], [f01[0]], [1], [0], [7]).copy())
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ta, tb = threading.Thread(target=extractor), threading.Thread(target=matcher)
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not errs, errs
    assert len(out["lines"]) == 6 and len(out["recs"]) == 40
    for got in out["lines"]:
        for g, s in zip(got, serial_lines):
            assert g.tobytes() == s.tobytes()
    for r in out["recs"]:
        assert r.tobytes() == serial_rec.tobytes()
    assert np.array_equal(ctx.match_pair_batch([f01[1]], [f01[0]], [1], [0], [7]), serial_rec)
    assert np.array_equal(ctx.pair_matches(0, 0), serial_m)
    ctx.close()


def test_two_contexts_run_concurrently(api, small_frames):
    """One context per host thread is the concurrent mode: both threads make progress and agree with each other."""
    imgs, deps, poses, K = small_frames
    res = [None, None]

    def work(i):
        ctx = api.Context(max_batch=2, max_w=320, max_h=240)
        fr = ctx.extract_batch(imgs, deps, K, seeds=[1, 2])
        res[i] = [f.lines().copy() for f in fr]
        ctx.close()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert res[0] is not None and res[1] is not None
    for a, b in zip(res[0], res[1]):
        assert a.tobytes() == b.tobytes() and len(a) > 0
