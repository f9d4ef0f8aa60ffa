"""Host-side mirror of the reference's GraphManager front half (include/lsl_graph.h -> liblsl_b200.so):
candidate selection, the trafo gates and the edge / keyframe bookkeeping around Node::matchNodePair
(src/graph_manager.cpp:204-319, 419-708, 730-860, 901-1006). Names follow the reference.
All decisions are taken inside the C library; this file only marshals. No CPU re-implementation lives here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import _check, lib
from .records import POSE_DTYPE, ptr


class GraphParams(C.Structure):
    _fields_ = [("min_translation_meter", C.c_double), ("min_rotation_degree", C.c_double),
                ("max_translation_meter", C.c_double)] + [(n, C.c_int32) for n in (
                    "max_rotation_degree predecessor_candidates neighbor_candidates min_sampled_candidates geodesic_depth "
                    "min_matches keep_all_nodes keep_good_nodes clear_non_keyframes clear_past_point_cloud largest_loop "
                    "line_match_number_weight").split()]


GRAPH_NODE_DTYPE = np.dtype([(n, "<i4") for n in
                             "id seq_id vertex_id matchable valid_tf_estimate has_lines n_feat2d n_feat3d".split()] +
                            [("stamp", "<f8"), ("estimate", "<f8", 16)])
assert GRAPH_NODE_DTYPE.itemsize == 168
GRAPH_EDGE_DTYPE = np.dtype([("id1", "<i4"), ("id2", "<i4"), ("n_inliers", "<i4"), ("pad", "<i4"), ("info", "<f8"),
                             ("transform", "<f8", 16)])
assert GRAPH_EDGE_DTYPE.itemsize == 152


class NodeResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                "found_match node_id in_graph edges_added keyframe_added best_id1 replaced_first n_candidates".split()]


FIRST, SKIPPED, COMPARE_PREDECESSOR, CANDIDATES, DROPPED = range(5)


def default_graph_params() -> GraphParams:
    p = GraphParams()
    lib().lsl_graph_params_default(C.byref(p))
    return p


def lineslam_launch_params() -> GraphParams:
    p = GraphParams()
    lib().lsl_graph_params_lineslam_launch(C.byref(p))
    return p


def isBigTrafo(T, params: GraphParams) -> bool:
    t = np.ascontiguousarray(T, np.float64).reshape(16)
    return bool(_check(lib().lsl_is_big_trafo(ptr(t), C.byref(params))))


def isSmallTrafo(T, seconds: float, params: GraphParams) -> bool:
    t = np.ascontiguousarray(T, np.float64).reshape(16)
    return bool(_check(lib().lsl_is_small_trafo(ptr(t), C.c_double(seconds), C.byref(params))))


class GraphManager:
    """addNode(node) registers the node against the candidates the reference would pick — one
    lsl_match_pair_batch per phase — and keeps nodes, chained vertex estimates, edges and keyframes."""

    def __init__(self, params: GraphParams | None = None, seed: int = 1, ctx=None):
        self.params = params if params is not None else default_graph_params()
        self.ctx = ctx
        h = C.c_void_p()
        L = lib()
        L.lsl_graph_destroy.argtypes = [C.c_void_p]
        L.lsl_graph_destroy.restype = None
        L.lsl_graph_num_nodes.argtypes = [C.c_void_p]
        _check(L.lsl_graph_create(C.byref(h), C.byref(self.params), C.c_uint32(seed)))
        self._h = h
        self.graph_ = {}       # id -> api.Node (keeps the frames alive)

    def close(self):
        if self._h:
            lib().lsl_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the three phases (usable without a device: the caller supplies the pose records) ----
    def node_begin(self, stamp: float, n_feat2d: int, n_feat3d: int):
        a, nid, cmp_ = C.c_int(0), C.c_int(-1), C.c_int(-1)
        _check(lib().lsl_graph_node_begin(self._h, C.c_double(stamp), n_feat2d, n_feat3d, C.byref(a), C.byref(nid), C.byref(cmp_)))
        return a.value, nid.value, cmp_.value

    def node_predecessor(self, rec):
        a, n = C.c_int(0), C.c_int(0)
        ids = np.zeros(self.num_nodes() + 8, np.int32)
        res = NodeResult()
        r = None if rec is None else np.ascontiguousarray(rec, POSE_DTYPE).reshape(1)
        _check(lib().lsl_graph_node_predecessor(self._h, ptr(r), C.byref(a), ptr(ids), len(ids), C.byref(n), C.byref(res)))
        return a.value, ids[:n.value].copy(), res

    def node_commit(self, recs):
        r = np.ascontiguousarray(recs, POSE_DTYPE)
        res = NodeResult()
        _check(lib().lsl_graph_node_commit(self._h, ptr(r) if len(r) else None, len(r), C.byref(res)))
        return res

    def getPotentialEdgeTargetsWithDijkstra(self, sequential_targets, geodesic_targets, sampled_targets, predecessor_id=-1,
                                            include_predecessor=False):
        ids = np.zeros(self.num_nodes() + 8, np.int32)
        n = C.c_int(0)
        _check(lib().lsl_graph_potential_edge_targets(self._h, sequential_targets, geodesic_targets, sampled_targets,
                                                      predecessor_id, int(include_predecessor), ptr(ids), len(ids), C.byref(n)))
        return ids[:n.value].copy()

    # ---- addNode against the device (src/graph_manager.cpp:730) ----
    def addNode(self, node, stamp: float, n_feat2d: int | None = None, n_feat3d: int | None = None, seed: int = 1) -> NodeResult:
        npts = node.frame.num_points
        res = NodeResult()
        _check(lib().lsl_graph_add_frame(self._h, self.ctx._h, node.frame._h, C.c_double(stamp),
                                         npts if n_feat2d is None else n_feat2d, npts if n_feat3d is None else n_feat3d,
                                         C.c_uint32(seed), C.byref(res)), self.ctx._h)
        if res.replaced_first:
            self.graph_.clear()
        if res.in_graph:
            node.id_ = res.node_id
            self.graph_[res.node_id] = node
        return res

    # ---- state ----
    def num_nodes(self) -> int:
        return _check(lib().lsl_graph_num_nodes(self._h))

    def nodes(self) -> np.ndarray:
        out = np.zeros(max(self.num_nodes(), 1), GRAPH_NODE_DTYPE)
        n = C.c_int(0)
        _check(lib().lsl_graph_nodes(self._h, ptr(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def edges(self) -> np.ndarray:
        n = C.c_int(0)
        lib().lsl_graph_edges(self._h, None, 0, C.byref(n))
        out = np.zeros(max(n.value, 1), GRAPH_EDGE_DTYPE)
        _check(lib().lsl_graph_edges(self._h, ptr(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def keyframe_ids(self) -> np.ndarray:
        out = np.zeros(self.num_nodes() + 1, np.int32)
        n = C.c_int(0)
        _check(lib().lsl_graph_keyframes(self._h, ptr(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def write_poses_2file(self, filename: str):
        """TUM trajectory format (src/graph_manager.cpp:864-884)."""
        _check(lib().lsl_graph_write_poses(self._h, filename.encode()))
