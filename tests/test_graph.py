"""SURVEY.md §8f row 2 — candidate selection and edge bookkeeping: the C ABI of include/lsl_graph.h against
oracle/oracle_graph.py (real glibc rand(), the reference's control flow restated line by line) on scripted
registration results. No device call is made here: the pose records are synthetic, identical for both sides."""
import math

import numpy as np
import pytest

from lineslam_b200 import graph as G
from lineslam_b200.records import POSE_DTYPE
from oracle import oracle_graph as OG

FIELDS = ("min_translation_meter min_rotation_degree max_translation_meter max_rotation_degree predecessor_candidates "
          "neighbor_candidates min_sampled_candidates geodesic_depth min_matches keep_all_nodes keep_good_nodes "
          "clear_non_keyframes clear_past_point_cloud largest_loop").split()


def _params(**kw):
    po, pp = OG.GraphParams(**kw), G.default_graph_params()
    for f in FIELDS:
        setattr(pp, f, type(getattr(pp, f))(getattr(po, f)))
    return po, pp


def _script(seed, n_nodes, p_found=0.75, motion=0.05, p_few=0.05, p_jump=0.05):
    """Deterministic registration results: the record depends only on (seed, stream position of the new frame, old id)."""
    cur = [0]

    def rec(new, old):
        rng = np.random.default_rng([seed, cur[0], old])
        r = np.zeros(1, POSE_DTYPE)[0]
        r["id_train"], r["id_query"] = old, new
        if rng.random() >= p_found or new == old:
            return r
        r["found"] = 1
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = abs(rng.normal()) * motion * (abs(new - old) ** 0.5)
        if rng.random() < p_jump:
            ang *= 40
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = rng.normal(size=3) * motion * abs(new - old) * (30 if rng.random() < p_jump else 1)
        r["tf"] = T.astype(np.float32).reshape(16)
        r["n_inliers"] = rng.integers(5, 60)
        r["pad"][2] = rng.integers(0, 80)
        r["rmse"] = np.float32(0.01 + rng.random() * 0.05)
        return r
    rng = np.random.default_rng(seed)
    stamps = np.cumsum(rng.uniform(0.02, 0.05, n_nodes)) + 1305031453.0
    feats = [int(rng.integers(0, 15)) if rng.random() < p_few else int(rng.integers(25, 400)) for _ in range(n_nodes)]
    rec.cur = cur
    return rec, stamps, feats


def _edge_of(r):
    if not r["found"]:
        return OG.Edge()
    q = np.float32(np.float32(int(r["pad"][2]) + int(r["n_inliers"]) * 1) / (r["rmse"] * r["rmse"]))
    return OG.Edge(id1=int(r["id_train"]), id2=int(r["id_query"]), transform=[float(x) for x in r["tf"]], info=float(q),
                   n_inliers=int(r["pad"][2]))


def _run_oracle(po, seed, rec, stamps, feats):
    gm = OG.GraphManager(po, seed)
    found = []
    for i in range(len(stamps)):
        rec.cur[0] = i
        found.append(bool(gm.add_node(float(stamps[i]), feats[i], feats[i], lambda a, b: _edge_of(rec(a, b)))))
    return gm, found


def _run_product(pp, seed, rec, stamps, feats):
    gm = G.GraphManager(pp, seed)
    found, cand_log = [], []
    for i in range(len(stamps)):
        rec.cur[0] = i
        action, nid, cmp_ = gm.node_begin(float(stamps[i]), feats[i], feats[i])
        if action == G.FIRST:
            found.append(True); continue
        if action == G.SKIPPED:
            found.append(False); continue
        r0 = rec(nid, cmp_) if action == G.COMPARE_PREDECESSOR else None
        action, ids, res = gm.node_predecessor(r0)
        if action == G.DROPPED:
            found.append(bool(res.found_match)); continue
        cand_log.append((nid, [int(x) for x in ids]))
        recs = np.array([rec(nid, int(c)) for c in ids], POSE_DTYPE) if len(ids) else np.zeros(0, POSE_DTYPE)
        res = gm.node_commit(recs)
        found.append(bool(res.found_match))
    return gm, found, cand_log


CASES = {
    "lineslam_launch": dict(kw=dict(min_translation_meter=0.01, min_rotation_degree=0.1, predecessor_candidates=1,
                                    neighbor_candidates=0, min_sampled_candidates=0, keep_all_nodes=True,
                                    clear_non_keyframes=True), n=80),
    "defaults_2_2_2": dict(kw=dict(), n=120),
    "octomap_5_5_5": dict(kw=dict(min_translation_meter=0.05, min_rotation_degree=1.0, predecessor_candidates=5,
                                  neighbor_candidates=5, min_sampled_candidates=5, max_translation_meter=20.0,
                                  max_rotation_degree=300, keep_good_nodes=True), n=150, motion=0.15),
    "largest_loop_depth2": dict(kw=dict(predecessor_candidates=3, neighbor_candidates=4, min_sampled_candidates=3,
                                        geodesic_depth=2, largest_loop=True, clear_non_keyframes=True,
                                        min_translation_meter=0.02), n=150),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("seed", [1, 7])
def test_graph_flow_matches_oracle(name, seed, tmp_path):
    case = CASES[name]
    po, pp = _params(**case["kw"])
    rec, stamps, feats = _script(100 * seed + len(name), case["n"], motion=case.get("motion", 0.05))
    ogm, ofound = _run_oracle(po, seed, rec, stamps, feats)
    pgm, pfound, pcands = _run_product(pp, seed, rec, stamps, feats)
    assert pfound == ofound
    assert pcands == ogm.log                                   # candidate lists, in order, for every inserted frame
    assert [int(k) for k in pgm.keyframe_ids()] == ogm.keyframe_ids
    nodes = pgm.nodes()
    assert len(nodes) == len(ogm.graph)
    for nd in nodes:
        o = ogm.graph[int(nd["id"])]
        assert (nd["seq_id"], nd["vertex_id"], bool(nd["matchable"]), bool(nd["valid_tf_estimate"]), bool(nd["has_lines"])) == \
               (o.seq_id, o.vertex_id, o.matchable, o.valid_tf_estimate, o.has_lines)
        assert nd["stamp"] == o.stamp
        assert nd["estimate"].tolist() == [float(x) for x in ogm.vertices[o.vertex_id]]      # bit-exact chaining
    edges = pgm.edges()
    assert len(edges) == len(ogm.edges)
    for e, o in zip(edges, ogm.edges):
        assert (e["id1"], e["id2"], e["n_inliers"]) == (o.id1, o.id2, o.n_inliers)
        assert e["transform"].tolist() == [float(x) for x in o.transform]
        assert e["info"] == o.info
    assert len(ofound) == case["n"] and case["n"] // 3 < sum(ofound) < case["n"]   # the script exercises both outcomes
    assert len(ogm.keyframe_ids) > 2 and len(ogm.edges) > len(ogm.graph) // 2
    # TUM trajectory file (write_poses_2file)
    fn = tmp_path / "poses.txt"
    pgm.write_poses_2file(str(fn))
    rows = [ln.split("\t") for ln in fn.read_text().splitlines()]
    want = [["%.16g" % v for v in row] for row in ogm.trajectory()]
    assert rows == want
    pgm.close()


def test_trafo_gates_match_oracle():
    rng = np.random.default_rng(5)
    po, pp = _params(min_translation_meter=0.05, min_rotation_degree=1.0, max_translation_meter=2.0, max_rotation_degree=90)
    for k in range(400):
        ang = float(rng.choice([0.0, 1e-9, 0.0174, 0.0175, 0.5, 3.1]))
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        T = np.eye(4); T[:3, :3] = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
        T[:3, 3] = rng.normal(size=3) * float(rng.choice([0.0, 0.03, 1.0]))
        if k % 7 == 0:
            T[:3, :3] *= 1.0 + 1e-7            # trace slightly above 3: acos -> NaN, every comparison false
        T = T.astype(np.float32).astype(np.float64)
        flat = [float(x) for x in T.reshape(16)]
        for dt in (0.0, -1.0, 0.033, 2.0):
            assert G.isSmallTrafo(T, dt, pp) == OG.is_small_trafo(flat, dt, po)
        assert G.isBigTrafo(T, pp) == OG.is_big_trafo(flat, po)


def test_standalone_candidate_query_and_errors():
    po, pp = _params(predecessor_candidates=3, neighbor_candidates=3, min_sampled_candidates=2)
    rec, stamps, feats = _script(42, 60, p_few=0.0)
    ogm, _ = _run_oracle(po, 3, rec, stamps, feats)
    want = [ogm.potential_edge_targets(2, 3, 2, -1, True), ogm.potential_edge_targets(1, 0, 4, 30, False)]
    pgm, _, _ = _run_product(pp, 3, rec, stamps, feats)
    got = [pgm.getPotentialEdgeTargetsWithDijkstra(2, 3, 2, -1, True), pgm.getPotentialEdgeTargetsWithDijkstra(1, 0, 4, 30, False)]
    assert [list(map(int, g)) for g in got] == want
    with pytest.raises(Exception):
        pgm.node_commit(np.zeros(0, POSE_DTYPE))     # no node in flight
    with pytest.raises(Exception):
        pgm.getPotentialEdgeTargetsWithDijkstra(1, 1, 1, 10_000, False)
    pgm.close()


def test_random_parameter_sweep_matches_oracle():
    """Forty random parameter sets (candidate counts 0..6, geodesic depth 1..4, every flag combination, motion gates
    on and off) x 50 frames: decisions, candidate lists, keyframes, edges and estimates equal the oracle's."""
    master = np.random.default_rng(2026)
    for trial in range(40):
        kw = dict(predecessor_candidates=int(master.integers(1, 7)), neighbor_candidates=int(master.integers(0, 7)),
                  min_sampled_candidates=int(master.integers(0, 7)), geodesic_depth=int(master.integers(1, 5)),
                  keep_all_nodes=bool(master.integers(0, 2)), keep_good_nodes=bool(master.integers(0, 2)),
                  clear_non_keyframes=bool(master.integers(0, 2)), clear_past_point_cloud=bool(master.integers(0, 2)),
                  largest_loop=bool(master.integers(0, 2)),
                  min_translation_meter=float(master.choice([0.0, 0.01, 0.05])), min_rotation_degree=float(master.choice([0.0, 0.5])),
                  max_translation_meter=float(master.choice([1e10, 5.0])), max_rotation_degree=int(master.choice([360, 120])),
                  min_matches=int(master.choice([20, 60])))
        po, pp = _params(**kw)
        rec, stamps, feats = _script(7000 + trial, 50, p_found=float(master.choice([0.5, 0.8, 1.0])), motion=float(master.choice([0.03, 0.1])),
                                     p_few=float(master.choice([0.0, 0.15])))
        seed = int(master.integers(1, 1000))
        ogm, ofound = _run_oracle(po, seed, rec, stamps, feats)
        pgm, pfound, pcands = _run_product(pp, seed, rec, stamps, feats)
        assert pfound == ofound, (trial, kw)
        assert pcands == ogm.log, (trial, kw)
        assert [int(k) for k in pgm.keyframe_ids()] == ogm.keyframe_ids, (trial, kw)
        e = pgm.edges()
        assert [(int(x["id1"]), int(x["id2"])) for x in e] == [(o.id1, o.id2) for o in ogm.edges], (trial, kw)
        for nd in pgm.nodes():
            o = ogm.graph[int(nd["id"])]
            assert nd["estimate"].tolist() == [float(x) for x in ogm.vertices[o.vertex_id]], (trial, kw)
            assert (bool(nd["matchable"]), bool(nd["valid_tf_estimate"]), bool(nd["has_lines"])) == (o.matchable, o.valid_tf_estimate, o.has_lines)
        pgm.close()


def test_new_initial_node_branch():
    """graph_manager.cpp:816-823: with a single node in the graph and a failed comparison, a frame with more features
    replaces the initial node (resetGraph + firstNode); later failures leave the graph alone."""
    po, pp = _params(min_translation_meter=0.01, predecessor_candidates=2)
    rec, stamps, feats = _script(11, 30, p_found=0.0)          # no pair ever registers
    feats = [30] + [200] * 29
    ogm, ofound = _run_oracle(po, 3, rec, stamps, feats)
    pgm, pfound, pcands = _run_product(pp, 3, rec, stamps, feats)
    assert ofound == pfound == [True, True] + [False] * 28
    assert pcands == ogm.log
    assert pgm.num_nodes() == len(ogm.graph) == 1
    nd = pgm.nodes()[0]
    assert (int(nd["n_feat2d"]), int(nd["seq_id"]), nd["stamp"]) == (200, ogm.graph[0].seq_id, ogm.graph[0].stamp)
    assert [int(k) for k in pgm.keyframe_ids()] == ogm.keyframe_ids == [0]
    pgm.close()


def test_bad_record_ids_are_refused_before_any_mutation():
    """ids in caller-supplied (or all-gathered) pose records are array indices inside the library: a found record whose
    id_train is out of range, whose id_query is not the node in flight, or (commit) whose id_train is not the candidate
    it answers must come back as LSL_ERR_ARG with the graph untouched — the reference asserts n->id_ == edge.id."""
    import ctypes as C
    from lineslam_b200.api import lib
    po, pp = _params(predecessor_candidates=3, neighbor_candidates=2, min_sampled_candidates=2, min_translation_meter=0.01)
    rec, stamps, feats = _script(5, 30, p_few=0.0, p_found=1.0)
    gm, _, _ = _run_product(pp, 3, rec, stamps[:20], feats[:20])
    n_nodes, n_edges = gm.num_nodes(), len(gm.edges())
    rec.cur[0] = 20
    action, nid, cmp_ = gm.node_begin(float(stamps[20]), feats[20], feats[20])
    assert action == G.COMPARE_PREDECESSOR
    for bad_train, bad_query in ((10_000, nid), (-7, nid), (cmp_, nid + 3)):
        r = rec(nid, cmp_).copy()
        r["found"] = 1; r["id_train"] = bad_train; r["id_query"] = bad_query
        with pytest.raises(Exception):
            gm.node_predecessor(r)
        assert gm.num_nodes() == n_nodes and len(gm.edges()) == n_edges
    # a too-small id buffer reports the capacity error; the retry (rec = NULL) returns the cached candidates
    L = lib()
    good = np.ascontiguousarray(rec(nid, cmp_), POSE_DTYPE).reshape(1)
    a, n = C.c_int(0), C.c_int(0)
    res = G.NodeResult()
    ids = np.zeros(64, np.int32)
    rc = L.lsl_graph_node_predecessor(gm._h, good.ctypes.data_as(C.c_void_p), C.byref(a), ids.ctypes.data_as(C.c_void_p), 0, C.byref(n), C.byref(res))
    if a.value == G.CANDIDATES and n.value > 0:
        assert rc == -3
        want = n.value
        rc = L.lsl_graph_node_predecessor(gm._h, None, C.byref(a), ids.ctypes.data_as(C.c_void_p), 64, C.byref(n), C.byref(res))
        assert rc == 0 and n.value == want
        cands = ids[:want].copy()
        recs = np.array([rec(nid, int(c)) for c in cands], POSE_DTYPE)
        bad = recs.copy(); bad["found"] = 1; bad["id_train"] = 9999
        with pytest.raises(Exception):
            gm.node_commit(bad)
        swapped = recs.copy(); swapped["found"] = 1
        if want >= 2:
            swapped["id_train"] = swapped["id_train"][::-1].copy()
            if not np.array_equal(swapped["id_train"], recs["id_train"]):
                with pytest.raises(Exception):
                    gm.node_commit(swapped)
        gm.node_commit(recs)       # the valid records still go through afterwards
    gm.close()
