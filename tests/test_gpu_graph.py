"""SURVEY.md §8f row 2 on the device: GraphManager.addNode (lsl_graph_add_frame -> lsl_match_pair_batch, one call per
phase) over a synthetic stream, against the oracle's control flow (oracle/oracle_graph.py) replayed with pose records
registered pair by pair through the same C ABI on a second set of frames. Checks the native driver: which pairs it
registers, with which seeds, in which order, the clear_past_point_cloud line release, and the resulting graph."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 10


@pytest.fixture(scope="module")
def two_sets(api):
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(N, scene_seed=2000)
    K = synth.camera_K()
    ctx = api.Context(max_batch=N, max_w=640, max_h=480)
    a = ctx.extract_batch(imgs, deps, K, seeds=list(range(1, N + 1)))
    b = ctx.extract_batch(imgs, deps, K, seeds=list(range(1, N + 1)))
    yield ctx, a, b, poses
    ctx.close()


@pytest.mark.parametrize("clear_past", [False, True])
def test_add_frame_flow(api, two_sets, clear_past, tmp_path):
    from lineslam_b200 import graph as G
    from oracle import oracle_graph as OG
    ctx, fa, fb, poses = two_sets
    if clear_past:   # the sweep releases lines for good: work on fresh copies of the module's frames
        fa = [ctx.frame_from_lines(f.lines()) for f in fa]
        fb = [ctx.frame_from_lines(f.lines()) for f in fb]
    kw = dict(min_translation_meter=0.001, min_rotation_degree=0.01, predecessor_candidates=3, neighbor_candidates=2,
              min_sampled_candidates=2, keep_all_nodes=True, clear_past_point_cloud=clear_past)
    po = OG.GraphParams(**kw)
    pp = G.default_graph_params()
    for k, v in kw.items():
        setattr(pp, k, type(getattr(pp, k))(v))
    stamps = [1305031453.0 + i / 30.0 for i in range(N)]

    # product: native driver
    gm = G.GraphManager(pp, seed=11, ctx=ctx)
    nodes = [api.Node(ctx, None, None, None, node_id=i, frame=fa[i]) for i in range(N)]
    results = [gm.addNode(nodes[i], stamps[i], n_feat2d=100, n_feat3d=100, seed=1000 * (i + 1)) for i in range(N)]

    # oracle control flow; registrations pair by pair on the second frame set
    calls = []
    state = dict(i=0, k=0)

    def match(new, old):
        seed = 1000 * (state["i"] + 1) + state["k"]
        state["k"] += 1
        rec = ctx.match_pair_batch([fb[state["i"]]], [fb_by_id[old]], [new], [old], [seed])[0]
        calls.append((new, old, seed, int(rec["found"])))
        if not rec["found"]:
            return OG.Edge()
        q = np.float32(np.float32(int(rec["pad"][2]) + int(rec["n_inliers"]) * 1) / (rec["rmse"] * rec["rmse"]))
        return OG.Edge(id1=old, id2=new, transform=[float(x) for x in rec["tf"]], info=float(q), n_inliers=int(rec["pad"][2]))

    ogm = OG.GraphManager(po, seed=11)
    fb_by_id = {}
    ofound = []
    for i in range(N):
        state["i"], state["k"] = i, 0
        if po.min_translation_meter <= 0 and po.min_rotation_degree <= 0:
            state["k"] = 1
        fb_by_id[len(ogm.graph)] = fb[i]       # the id the frame gets if it is inserted (it can be its own candidate)
        ofound.append(bool(ogm.add_node(stamps[i], 100, 100, match)))
        if clear_past:
            for nid, nd in ogm.graph.items():
                if not nd.has_lines and fb_by_id[nid].num_lines > 0:   # same release as the product's sweep
                    fb_by_id[nid].clear_lines()

    assert [bool(r.found_match) for r in results] == ofound
    assert sum(ofound) == N                                   # every frame of the smooth stream is inserted
    assert [int(k) for k in gm.keyframe_ids()] == ogm.keyframe_ids
    got_e, want_e = gm.edges(), ogm.edges
    assert [(int(e["id1"]), int(e["id2"]), int(e["n_inliers"])) for e in got_e] == [(e.id1, e.id2, e.n_inliers) for e in want_e]
    for e, o in zip(got_e, want_e):
        assert e["transform"].tolist() == [float(x) for x in o.transform]
        assert e["info"] == o.info
    for nd in gm.nodes():
        o = ogm.graph[int(nd["id"])]
        assert nd["estimate"].tolist() == [float(x) for x in ogm.vertices[o.vertex_id]]
        assert bool(nd["has_lines"]) == o.has_lines and bool(nd["valid_tf_estimate"]) == o.valid_tf_estimate
    if clear_past:
        # the reference's quirk: only the newest node keeps its lines, so nothing but the predecessor (and the frame
        # itself, when the Dijkstra neighbourhood hands it back as its own candidate) ever registers
        assert all(nodes[i].frame.num_lines == 0 for i in range(N - 1)) and nodes[N - 1].frame.num_lines > 0
        assert all(abs(e.id1 - e.id2) <= 1 for e in want_e if e.info >= 0)
    else:
        assert len(want_e) > 2 * N                            # predecessors + geodesic / sampled candidates register
        # chained estimates follow the synthetic ground truth at the centimetre level
        from lineslam_b200 import synth
        for nd in gm.nodes():
            gt = synth.relative_pose_q2t(*poses[int(nd["id"])], *poses[0])     # node -> first node coordinates
            assert np.abs(nd["estimate"].reshape(4, 4)[:3, 3] - gt[:3, 3]).max() < 0.05
    fn = tmp_path / "traj.txt"
    gm.write_poses_2file(str(fn))
    assert len(fn.read_text().splitlines()) == sum(o.valid_tf_estimate for o in ogm.graph.values())
    gm.close()
