"""TEST INFRASTRUCTURE — CPU restatement of the reference's candidate selection and edge bookkeeping
(SURVEY.md §8f row 2). Only tests/ may import this; the product (lineslam_b200/csrc/lsl_graph.cu) never does.

Restates, function by function (all paths into /root/reference/):
  trafoSize / isBigTrafo / isSmallTrafo            src/misc.cpp:254-296 (Eigen::Isometry3d overloads)
  GraphManager::getPotentialEdgeTargetsWithDijkstra src/graph_manager.cpp:204-319
  GraphManager::firstNode                           src/graph_manager.cpp:358-400
  GraphManager::nodeComparisons                     src/graph_manager.cpp:419-708 (concurrent_edge_construction branch)
  GraphManager::addNode                             src/graph_manager.cpp:730-860 (mapping branch, incl. the
                                                    clear_past_point_cloud sweep :845-857)
  GraphManager::addKeyframe                         src/graph_manager.cpp:901-926
  GraphManager::addEdgeToG2O                        src/graph_manager.cpp:928-1006 (vertex creation / estimate chaining)

Third-party piece absent from the reference tree: g2o::HyperDijkstra::shortestPaths with g2o::UniformCostFunction
(g2o core/hyper_dijkstra.cpp, the ROS-packaged libg2o the reference links; not vendored). Its published algorithm
is a Dijkstra over the hyper-graph in which a vertex z is relaxed when dist(u) + 1 < dist(z) AND dist(u) + 1 <
maxDistance (strict), so `visited()` = the source plus every vertex at hop distance < geodesic_depth. UNVERIFIED
against g2o itself (parity unpinned for this one function); everything else follows the reference's own lines.

rand() is the process's real glibc rand() (ctypes), seeded with srand(seed) like src/main.cpp:168 — this also pins
the product's replay of the TYPE_3 generator (csrc/shared/lsl_rand.h).

Author quirks kept on purpose:
  * `MatchingResult mr; int prev_best = mr.edge.id1;` (graph_manager.cpp:457-458) reads a fresh result, so prev_best is
    always -1 and the "reuse best matched node" append (:533-535) never fires.
  * earliest_loop_closure_node_ only moves for pose_relative_to == "largest_loop" (:998-1001); with the default
    "first" every node without an edge to a keyframe makes its predecessor a keyframe (:795-797).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field

_libc = ctypes.CDLL("libc.so.6")
_libc.rand.restype = ctypes.c_int


@dataclass
class GraphParams:  # src/parameter_server.cpp:82-149,196 defaults
    min_translation_meter: float = 0.0
    min_rotation_degree: float = 0.0
    max_translation_meter: float = 1e10
    max_rotation_degree: int = 360
    predecessor_candidates: int = 2
    neighbor_candidates: int = 2
    min_sampled_candidates: int = 2
    geodesic_depth: int = 3
    min_matches: int = 20
    keep_all_nodes: bool = False
    keep_good_nodes: bool = False
    clear_non_keyframes: bool = False
    clear_past_point_cloud: bool = True
    largest_loop: bool = False  # pose_relative_to == "largest_loop"


def lineslam_launch_params() -> GraphParams:
    """launch/lineslam.launch:15-36."""
    return GraphParams(min_translation_meter=0.01, min_rotation_degree=0.1, predecessor_candidates=1,
                       neighbor_candidates=0, min_sampled_candidates=0, keep_all_nodes=True, clear_non_keyframes=True)


@dataclass
class Edge:
    id1: int = -1
    id2: int = -1
    transform: list = field(default_factory=lambda: [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0])  # row-major 4x4
    info: float = 1.0            # informationMatrix = info * I6 (node.cpp:1531-1532); -1: the constant-position edge
    n_inliers: int = 0           # mr.inlier_matches.size() (point inliers)


@dataclass
class NodeState:
    id: int
    seq_id: int
    stamp: float
    n2d: int
    n3d: int
    vertex_id: int = -1
    matchable: bool = True
    valid_tf_estimate: bool = True
    has_lines: bool = True


def trafo_size(T):
    """misc.cpp:254-258. Isometry3d::rotation() is the linear part for Mode == Isometry."""
    tr = T[0] + T[5] + T[10]
    c = (tr - 1) / 2
    angle = (math.acos(c) if -1.0 <= c <= 1.0 else float("nan")) * 180.0 / math.pi   # libm acos: NaN outside [-1, 1]
    dist = math.sqrt(T[3] * T[3] + T[7] * T[7] + T[11] * T[11])
    return angle, dist


def is_big_trafo(T, P: GraphParams):
    angle, dist = trafo_size(T)
    return dist > P.min_translation_meter or angle > P.min_rotation_degree


def is_small_trafo(T, seconds, P: GraphParams):
    if seconds <= 0.0:
        return True
    angle, dist = trafo_size(T)
    return dist / seconds < P.max_translation_meter and angle / seconds < P.max_rotation_degree


def iso_mul(A, B):
    """Isometry3d * Isometry3d: linear = A.lin * B.lin, translation = A.lin * B.t + A.t, last row 0 0 0 1."""
    C = [0.0] * 16
    for r in range(3):
        for c in range(3):
            C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c]
        C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3]
    C[15] = 1.0
    return C


def iso_inv(A):
    """Isometry inverse: R^T, -R^T t."""
    C = [0.0] * 16
    for r in range(3):
        for c in range(3):
            C[4 * r + c] = A[4 * c + r]
    for r in range(3):
        C[4 * r + 3] = -(C[4 * r] * A[3] + C[4 * r + 1] * A[7] + C[4 * r + 2] * A[11])
    C[15] = 1.0
    return C


class GraphManager:
    def __init__(self, params: GraphParams, seed: int = 1):
        self.P = params
        _libc.srand(ctypes.c_uint(seed))
        self.graph: dict[int, NodeState] = {}
        self.vertices: dict[int, list] = {}      # vertex id -> estimate (row-major 4x4)
        self.adj: dict[int, set] = {}            # vertex id -> neighbouring vertex ids (cam_cam_edges_)
        self.edges: list[Edge] = []
        self.keyframe_ids: list[int] = []
        self.next_seq_id = 0
        self.next_vertex_id = 0
        self.earliest_loop_closure_node = 0
        self.curr_best = Edge()
        self.loop_closures_edges = 0
        self.sequential_edges = 0
        self.log: list = []                      # (node id, candidate list) per nodeComparisons call

    @staticmethod
    def _rand():
        return _libc.rand()

    # ---- graph_manager.cpp:204-319
    def potential_edge_targets(self, sequential_targets, geodesic_targets, sampled_targets, predecessor_id=-1,
                               include_predecessor=False):
        ids = []
        n = len(self.graph)
        if predecessor_id < 0:
            predecessor_id = n - 1
        if len(self.vertices) <= sequential_targets + geodesic_targets + sampled_targets or len(self.vertices) <= 1:
            sequential_targets = sequential_targets + geodesic_targets + sampled_targets
            geodesic_targets = 0
            sampled_targets = 0
            predecessor_id = n - 1
        if sequential_targets > 0:
            i = 1
            while i < sequential_targets + 1 and predecessor_id - i >= 0:
                ids.append(predecessor_id - i)
                i += 1
        if geodesic_targets > 0:
            src = self.graph[predecessor_id].vertex_id
            dist = {src: 0}
            frontier = [src]
            while frontier:  # uniform costs: breadth-first == Dijkstra
                nxt = []
                for u in frontier:
                    for z in sorted(self.adj.get(u, ())):
                        if z not in dist and dist[u] + 1 < self.P.geodesic_depth:
                            dist[z] = dist[u] + 1
                            nxt.append(z)
                frontier = nxt
            v2n = {nd.vertex_id: nd.id for nd in self.graph.values()}
            neigh = {}
            sum_w = 0
            for vid in dist:
                nid = v2n[vid]
                if not self.graph[nid].matchable:
                    continue
                if nid < predecessor_id - sequential_targets or (predecessor_id < nid <= n - 1):
                    w = abs(predecessor_id - nid)
                    neigh[nid] = w
                    sum_w += w
            while len(ids) < sequential_targets + geodesic_targets and len(neigh) != 0:
                pick = self._rand() % sum_w
                acc = 0
                for nid in sorted(neigh):        # std::map order
                    acc += neigh[nid]
                    if acc > pick:
                        ids.insert(0, nid)
                        sum_w -= neigh[nid]
                        del neigh[nid]
                        break
        if sampled_targets > 0:
            non = [k for k in self.keyframe_ids if k not in ids and self.graph[k].matchable]
            while len(ids) < geodesic_targets + sampled_targets + sequential_targets and len(non) != 0:
                j = self._rand() % len(non)
                sid = non[j]
                non[j] = non[-1]
                non.pop()
                ids.insert(0, sid)
        if include_predecessor:
            ids.append(predecessor_id)
        return ids

    # ---- graph_manager.cpp:901-926
    def add_keyframe(self, nid):
        if self.P.clear_non_keyframes and len(self.keyframe_ids) >= 2:
            most, second = self.keyframe_ids[-1], self.keyframe_ids[-2]
            for nd in self.graph.values():
                if second < nd.id < most:
                    nd.matchable = False
        self.keyframe_ids.append(nid)

    # ---- graph_manager.cpp:928-1006
    def add_edge(self, e: Edge, large_edge: bool, set_estimate: bool):
        n1, n2 = self.graph.get(e.id1), self._pending if e.id2 == self._pending.id else self.graph.get(e.id2)
        v1 = n1.vertex_id if n1.vertex_id in self.vertices else -1
        v2 = n2.vertex_id if n2.vertex_id in self.vertices else -1
        if (v1 < 0 or v2 < 0) and not large_edge:
            return False
        if v1 < 0 and v2 < 0:
            return False
        if v1 < 0:
            v1 = self.next_vertex_id
            self.next_vertex_id += 1
            n1.vertex_id = v1
            self.vertices[v1] = iso_mul(self.vertices[v2], iso_inv(e.transform))
        elif v2 < 0:
            v2 = self.next_vertex_id
            self.next_vertex_id += 1
            n2.vertex_id = v2
            self.vertices[v2] = iso_mul(self.vertices[v1], e.transform)
        elif set_estimate:
            self.vertices[v2] = iso_mul(self.vertices[v1], e.transform)
        self.adj.setdefault(v1, set()).add(v2)
        self.adj.setdefault(v2, set()).add(v1)
        self.edges.append(e)
        if abs(e.id1 - e.id2) > self.P.predecessor_candidates:
            self.loop_closures_edges += 1
        else:
            self.sequential_edges += 1
        if self.P.largest_loop:
            self.earliest_loop_closure_node = min(self.earliest_loop_closure_node, e.id1, e.id2)
        return True

    # ---- graph_manager.cpp:358-400
    def first_node(self, nd: NodeState):
        nd.id = len(self.graph)
        nd.seq_id = self.next_seq_id
        self.next_seq_id += 1
        nd.vertex_id = self.next_vertex_id
        self.next_vertex_id += 1
        self.graph[nd.id] = nd
        self.vertices[nd.vertex_id] = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]
        self.add_keyframe(nd.id)

    # ---- graph_manager.cpp:419-708; match(new_id, old_id) -> Edge with id1 < 0 when no transformation was found
    def node_comparisons(self, nd: NodeState, match):
        P = self.P
        if nd.n2d < P.min_matches and not P.keep_all_nodes:
            return False, False
        nd.id = len(self.graph)
        nd.seq_id = self.next_seq_id
        self.next_seq_id += 1
        self._pending = nd
        self.earliest_loop_closure_node = nd.id
        edges_before = len(self.edges)
        edge_to_keyframe = False
        seq_prev = max(self.graph)
        self.curr_best = Edge()
        predecessor_matched = False
        if P.min_translation_meter > 0.0 or P.min_rotation_degree > 0.0:
            prev = self.graph[len(self.graph) - 1]
            mr = match(nd.id, prev.id)
            if mr.id1 >= 0 and mr.id2 >= 0:
                dt = nd.stamp - prev.stamp
                if not is_big_trafo(mr.transform, P) or not is_small_trafo(mr.transform, dt, P):
                    self.curr_best = mr
                    return False, False
                if self.add_edge(mr, True, True):
                    self.graph[nd.id] = nd
                    if mr.id1 in self.keyframe_ids:
                        edge_to_keyframe = True
                    self.graph[mr.id1].valid_tf_estimate = True
                    self.curr_best = mr
                else:
                    return False, False
                predecessor_matched = True
        seq_cand = P.predecessor_candidates - 1
        if predecessor_matched:
            cands = self.potential_edge_targets(seq_cand, P.neighbor_candidates, P.min_sampled_candidates, self.curr_best.id1)
        else:
            cands = self.potential_edge_targets(seq_cand, P.neighbor_candidates, P.min_sampled_candidates, seq_prev, True)
        self.log.append((nd.id, list(cands)))
        results = [match(nd.id, c) for c in cands]   # QtConcurrent::blockingMapped, results in list order
        for mr in results:
            if mr.id1 >= 0:
                dt = nd.stamp - self.graph[mr.id1].stamp
                if is_small_trafo(mr.transform, dt, P) and self.add_edge(
                        mr, is_big_trafo(mr.transform, P), mr.n_inliers > self.curr_best.n_inliers):
                    self.graph[nd.id] = nd
                    self.graph[mr.id1].valid_tf_estimate = True
                    if mr.n_inliers > self.curr_best.n_inliers:
                        self.curr_best = mr
                    if mr.id1 in self.keyframe_ids:
                        edge_to_keyframe = True
        found_trafo = len(self.edges) != edges_before
        keep_anyway = P.keep_all_nodes or (nd.n3d > P.min_matches and P.keep_good_nodes)
        # odom_frame_name is empty in every launch file of the reference: invalid_odometry is always true (:630-632)
        if not found_trafo and keep_anyway:
            e = Edge(id1=seq_prev, id2=nd.id, info=-1.0)
            self.add_edge(e, True, True)
            self.graph[nd.id] = nd
            nd.valid_tf_estimate = False
            self.curr_best = e
        return len(self.edges) > edges_before, edge_to_keyframe

    # ---- graph_manager.cpp:730-860 (mapping branch)
    def add_node(self, stamp, n2d, n3d, match):
        nd = NodeState(id=-1, seq_id=-1, stamp=stamp, n2d=n2d, n3d=n3d)
        if len(self.graph) == 0:
            self.first_node(nd)
            return True
        found, edge_to_kf = self.node_comparisons(nd, match)
        if found:
            self.graph[nd.id] = nd
            if not edge_to_kf and self.earliest_loop_closure_node > self.keyframe_ids[-1]:
                self.add_keyframe(nd.id - 1)
        elif len(self.graph) == 1 and n2d > self.graph[0].n2d:   # "choosing new initial node" (:816-823)
            self._reset()
            nd2 = NodeState(id=-1, seq_id=-1, stamp=stamp, n2d=n2d, n3d=n3d)
            self.first_node(nd2)
            return True
        if self.P.clear_past_point_cloud:
            for x in self.graph.values():
                if x.id < len(self.graph) - 1:
                    x.has_lines = False
        return found

    def _reset(self):
        """resetGraph (graph_manager.cpp:326-356); rand() is not reseeded."""
        self.graph.clear(); self.vertices.clear(); self.adj.clear(); self.edges.clear(); self.keyframe_ids.clear()
        self.next_seq_id = self.next_vertex_id = 0
        self.curr_best = Edge()
        self.loop_closures_edges = self.sequential_edges = 0

    def trajectory(self):
        """write_poses_2file rows (graph_manager.cpp:864-884): (stamp, tx, ty, tz, qx, qy, qz, qw) of valid nodes."""
        out = []
        for nid in sorted(self.graph):
            nd = self.graph[nid]
            if not nd.valid_tf_estimate:
                continue
            T = self.vertices[nd.vertex_id]
            out.append((nd.stamp, T[3], T[7], T[11]) + r2q_xyzw(T))
        return out


def r2q_xyzw(T):
    """r2q (src/line/utils.cpp:1709-1720) in the output order qx qy qz qw of write_poses_2file."""
    t = T[0] + T[5] + T[10]
    r = math.sqrt(1 + t)
    s = 0.5 / r
    w = 0.5 * r
    return ((T[9] - T[6]) * s, (T[2] - T[8]) * s, (T[4] - T[1]) * s, w)
