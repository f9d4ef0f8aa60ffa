// lsl_graph.cu — host side of include/lsl_graph.h: which earlier frames a new frame is registered against, and
// what happens to the 128-byte pose records that come back. No kernels in this file: the
// registrations themselves are lsl_match_pair_batch (k_pair.cu / k_hybrid.cu), called once per phase for all
// candidates of a frame instead of once per candidate (src/graph_manager.cpp:555 maps matchNodePair over a QList).
//
// Layout: nodes, vertices and edges are flat vectors indexed by id (ids are dense: graph_.size() at insertion);
// the pose-graph adjacency is one small vector per vertex. Nothing here links or includes oracle/.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include <algorithm>
#include "../../include/lsl_graph.h"
#include "shared/lsl_rand.h"

namespace {

struct Iso { double m[16]; };

Iso iso_identity() { Iso r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return r; }

// Isometry3d product: linear = A.lin B.lin, translation = A.lin B.t + A.t (Eigen Transform * Transform, Isometry mode)
Iso iso_mul(const Iso& A, const Iso& B) {
  Iso C;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) C.m[4 * r + c] = A.m[4 * r] * B.m[c] + A.m[4 * r + 1] * B.m[4 + c] + A.m[4 * r + 2] * B.m[8 + c];
    C.m[4 * r + 3] = A.m[4 * r] * B.m[3] + A.m[4 * r + 1] * B.m[7] + A.m[4 * r + 2] * B.m[11] + A.m[4 * r + 3];
  }
  C.m[12] = C.m[13] = C.m[14] = 0.0; C.m[15] = 1.0;
  return C;
}

Iso iso_inv(const Iso& A) {
  Iso C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C.m[4 * r + c] = A.m[4 * c + r];
  for (int r = 0; r < 3; ++r) C.m[4 * r + 3] = -(C.m[4 * r] * A.m[3] + C.m[4 * r + 1] * A.m[7] + C.m[4 * r + 2] * A.m[11]);
  C.m[12] = C.m[13] = C.m[14] = 0.0; C.m[15] = 1.0;
  return C;
}

// trafoSize (misc.cpp:254-258); Isometry3d::rotation() of an Isometry-mode transform is its linear part
void trafo_size(const double* T, double* angle, double* dist) {
  *angle = std::acos((T[0] + T[5] + T[10] - 1) / 2) * 180.0 / M_PI;
  *dist = std::sqrt(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
}

struct Node {
  int id = -1, seq_id = -1, vertex_id = -1;
  bool matchable = true, valid_tf = true, has_lines = true;
  int n2d = 0, n3d = 0;
  double stamp = 0;
  lsl_frame* frame = nullptr;   // borrowed (lsl_graph_add_frame only)
};

struct EdgeRec {
  int id1 = -1, id2 = -1, n_inliers = 0;
  double info = 1.0;
  Iso tf = iso_identity();
};

}  // namespace

struct lsl_graph {
  lsl_graph_params P;
  lslm::GRand rng;
  std::vector<Node> nodes;              // graph_ (id == index)
  std::vector<Iso> vest;                // vertex id -> estimate
  std::vector<std::vector<int>> adj;    // vertex id -> adjacent vertex ids (cam_cam_edges_)
  std::vector<EdgeRec> edges;
  std::vector<int> keyframes;
  int next_seq_id = 0, next_vertex_id = 0, earliest_loop_closure_node = 0;
  int loop_closure_edges = 0, sequential_edges = 0;
  EdgeRec curr_best;
  // the node between lsl_graph_node_begin and the end of its insertion
  Node pend;
  bool pend_active = false, pend_in_graph = false, pend_edge_to_kf = false, pend_pred_matched = false;
  size_t pend_edges_before = 0;
  int pend_seq_prev = -1;
  int phase = 0;                        // 0 idle, 1 waiting for the predecessor record, 2 waiting for commit
  std::vector<int> pend_cands;

  int rnd() { return lslm::grand_next(&rng); }
  Node& node_ref(int id) { return (pend_active && !pend_in_graph && id == pend.id) ? pend : nodes[(size_t)id]; }
  bool is_keyframe(int id) const { return std::find(keyframes.begin(), keyframes.end(), id) != keyframes.end(); }
  void put_pending_in_graph() {
    if (pend_in_graph) return;
    nodes.push_back(pend);               // graph_[new_node->id_] = new_node; id == graph_.size()
    pend_in_graph = true;
  }
  Node& pending() { return pend_in_graph ? nodes[(size_t)pend.id] : pend; }
};

namespace {

bool big_trafo(const double* T, const lsl_graph_params& P) {
  double a, d; trafo_size(T, &a, &d);
  return d > P.min_translation_meter || a > P.min_rotation_degree;
}
bool small_trafo(const double* T, double seconds, const lsl_graph_params& P) {
  if (seconds <= 0.0) return true;
  double a, d; trafo_size(T, &a, &d);
  return d / seconds < P.max_translation_meter && a / seconds < P.max_rotation_degree;
}

// graph_manager.cpp:901-926
void add_keyframe(lsl_graph* g, int id) {
  if (g->P.clear_non_keyframes && g->keyframes.size() >= 2) {
    const int most = g->keyframes.back(), second = g->keyframes[g->keyframes.size() - 2];
    for (Node& n : g->nodes)
      if (n.id > second && n.id < most) n.matchable = false;   // clearFeatureInformation (node.cpp:1709)
  }
  g->keyframes.push_back(id);
}

// graph_manager.cpp:928-1006
bool add_edge(lsl_graph* g, const EdgeRec& e, bool large_edge, bool set_estimate) {
  Node& n1 = g->node_ref(e.id1);
  Node& n2 = g->node_ref(e.id2);
  const bool has1 = n1.vertex_id >= 0, has2 = n2.vertex_id >= 0;
  if ((!has1 || !has2) && !large_edge) return false;
  if (!has1 && !has2) return false;
  auto new_vertex = [&](const Iso& est) { g->vest.push_back(est); g->adj.emplace_back(); return g->next_vertex_id++; };
  if (!has1) n1.vertex_id = new_vertex(iso_mul(g->vest[(size_t)n2.vertex_id], iso_inv(e.tf)));
  else if (!has2) n2.vertex_id = new_vertex(iso_mul(g->vest[(size_t)n1.vertex_id], e.tf));
  else if (set_estimate) g->vest[(size_t)n2.vertex_id] = iso_mul(g->vest[(size_t)n1.vertex_id], e.tf);
  const int v1 = n1.vertex_id, v2 = n2.vertex_id;
  auto link = [&](int a, int b) { auto& v = g->adj[(size_t)a]; if (std::find(v.begin(), v.end(), b) == v.end()) v.push_back(b); };
  link(v1, v2); link(v2, v1);
  g->edges.push_back(e);
  if (std::abs(e.id1 - e.id2) > g->P.predecessor_candidates) g->loop_closure_edges++; else g->sequential_edges++;
  if (g->P.largest_loop) g->earliest_loop_closure_node = std::min(g->earliest_loop_closure_node, std::min(e.id1, e.id2));
  return true;
}

// graph_manager.cpp:204-319. HyperDijkstra with UniformCostFunction up to geodesic_depth (g2o, not in the reference
// tree): every vertex whose hop distance from the predecessor's vertex is < geodesic_depth.
void edge_targets(lsl_graph* g, int seq_t, int geo_t, int samp_t, int pred, bool include_pred, std::vector<int>* out) {
  std::vector<int>& ids = *out;
  ids.clear();
  const int gsize = (int)g->nodes.size();
  const int nvert = (int)g->vest.size();
  if (pred < 0) pred = gsize - 1;
  if (nvert <= seq_t + geo_t + samp_t || nvert <= 1) { seq_t = seq_t + geo_t + samp_t; geo_t = 0; samp_t = 0; pred = gsize - 1; }
  if (seq_t > 0)
    for (int i = 1; i < seq_t + 1 && pred - i >= 0; ++i) ids.push_back(pred - i);
  if (geo_t > 0) {
    std::vector<int> dist((size_t)nvert, -1), frontier, next;
    const int src = g->nodes[(size_t)pred].vertex_id;
    dist[(size_t)src] = 0; frontier.push_back(src);
    while (!frontier.empty()) {
      next.clear();
      for (int u : frontier)
        for (int z : g->adj[(size_t)u])
          if (dist[(size_t)z] < 0 && dist[(size_t)u] + 1 < g->P.geodesic_depth) { dist[(size_t)z] = dist[(size_t)u] + 1; next.push_back(z); }
      frontier.swap(next);
    }
    std::vector<int> v2n((size_t)nvert, -1);
    for (const Node& n : g->nodes) if (n.vertex_id >= 0) v2n[(size_t)n.vertex_id] = n.id;
    std::map<int, int> neigh;   // node id -> weight, visited in id order like the reference's std::map
    int sum_w = 0;
    for (int v = 0; v < nvert; ++v) {
      if (dist[(size_t)v] < 0) continue;
      const int id = v2n[(size_t)v];
      if (id < 0 || !g->nodes[(size_t)id].matchable) continue;
      if (id < pred - seq_t || (id > pred && id <= gsize - 1)) { const int w = std::abs(pred - id); neigh[id] = w; sum_w += w; }
    }
    while ((int)ids.size() < seq_t + geo_t && !neigh.empty()) {
      const int pick = g->rnd() % sum_w;
      int acc = 0;
      for (auto it = neigh.begin(); it != neigh.end(); ++it) {
        acc += it->second;
        if (acc > pick) { ids.insert(ids.begin(), it->first); sum_w -= it->second; neigh.erase(it); break; }
      }
    }
  }
  if (samp_t > 0) {
    std::vector<int> non;
    for (int k : g->keyframes)
      if (std::find(ids.begin(), ids.end(), k) == ids.end() && g->nodes[(size_t)k].matchable) non.push_back(k);
    while ((int)ids.size() < geo_t + samp_t + seq_t && !non.empty()) {
      const int j = g->rnd() % (int)non.size();
      const int sid = non[(size_t)j];
      non[(size_t)j] = non.back(); non.pop_back();
      ids.insert(ids.begin(), sid);
    }
  }
  if (include_pred) ids.push_back(pred);
}

// MatchingResult.edge of one record (node.cpp:1524-1536)
EdgeRec edge_of(const lsl_pose_rec& r, const lsl_graph_params& P) {
  EdgeRec e;
  if (!r.found) return e;
  e.id1 = r.id_train; e.id2 = r.id_query;
  e.n_inliers = r.pad[2];
  for (int i = 0; i < 16; ++i) e.tf.m[i] = (double)r.tf[i];
  // size_t / (float * float): the quotient is formed in float, then scales the double identity (node.cpp:1531-1532)
  const float q = (float)((size_t)r.pad[2] + (size_t)r.n_inliers * (size_t)P.line_match_number_weight) / (r.rmse * r.rmse);
  e.info = (double)q;
  return e;
}

// A found record must name a node that exists (or the pending node itself, which can be drawn as its own geodesic
// candidate, :262) as the train frame and the pending node as the query frame (the reference asserts n->id_ == edge.id).
bool rec_ids_ok(const lsl_graph* g, const lsl_pose_rec& r, int expect_train) {
  if (!r.found) return true;
  const int nn = (int)g->nodes.size();
  const bool train_ok = (r.id_train >= 0 && r.id_train < nn) || (g->pend_active && !g->pend_in_graph && r.id_train == g->pend.id);
  if (!train_ok || r.id_query != g->pend.id) return false;
  return expect_train < 0 || r.id_train == expect_train;
}

void first_node(lsl_graph* g, Node nd) {   // graph_manager.cpp:358-400
  nd.id = (int)g->nodes.size();
  nd.seq_id = g->next_seq_id++;
  nd.vertex_id = g->next_vertex_id++;
  g->vest.push_back(iso_identity()); g->adj.emplace_back();
  g->nodes.push_back(nd);
  add_keyframe(g, nd.id);
}

void reset_graph(lsl_graph* g) {           // graph_manager.cpp:326-356 (rand() is not reseeded)
  g->nodes.clear(); g->vest.clear(); g->adj.clear(); g->edges.clear(); g->keyframes.clear();
  g->next_seq_id = g->next_vertex_id = 0; g->loop_closure_edges = g->sequential_edges = 0;
  g->curr_best = EdgeRec();
}

// tail of addNode after nodeComparisons returned `found` (graph_manager.cpp:748-860)
void finish_node(lsl_graph* g, bool found, lsl_graph_node_result* res) {
  lsl_graph_node_result r; std::memset(&r, 0, sizeof r);
  r.keyframe_added = -1;
  r.node_id = g->pend.id;
  r.edges_added = (int)(g->edges.size() - g->pend_edges_before);
  r.n_candidates = (int)g->pend_cands.size();
  if (found) {
    g->put_pending_in_graph();
    if (!g->pend_edge_to_kf && g->earliest_loop_closure_node > g->keyframes.back()) {
      add_keyframe(g, g->pend.id - 1);
      r.keyframe_added = g->pend.id - 1;
    }
  } else if (g->nodes.size() == 1 && g->pend.n2d > g->nodes[0].n2d) {   // "choosing new initial node" (:816-823)
    Node nd = g->pend;
    reset_graph(g);
    nd.id = nd.seq_id = nd.vertex_id = -1;
    first_node(g, nd);
    g->pend_in_graph = true; g->pend.id = 0;
    r.replaced_first = 1; r.node_id = 0; found = true; r.edges_added = 0;
  }
  r.found_match = found ? 1 : 0;
  r.in_graph = g->pend_in_graph ? 1 : 0;
  r.best_id1 = g->curr_best.id1;
  if (g->P.clear_past_point_cloud)                                        // :845-857
    for (Node& n : g->nodes)
      if (n.id < (int)g->nodes.size() - 1) n.has_lines = false;
  g->pend_active = false; g->phase = 0;
  if (res) *res = r;
}

// second half of nodeComparisons: the main loop over the candidates' records and the keep_anyway edge (:553-686)
void main_loop(lsl_graph* g, const lsl_pose_rec* recs, int n, lsl_graph_node_result* res) {
  const lsl_graph_params& P = g->P;
  for (int i = 0; i < n; ++i) {
    EdgeRec mr = edge_of(recs[i], P);
    if (mr.id1 < 0) continue;
    const double dt = g->pend.stamp - g->node_ref(mr.id1).stamp;
    if (small_trafo(mr.tf.m, dt, P) && add_edge(g, mr, big_trafo(mr.tf.m, P), mr.n_inliers > g->curr_best.n_inliers)) {
      g->pend.vertex_id = g->pending().vertex_id;
      g->put_pending_in_graph();
      g->node_ref(mr.id1).valid_tf = true;
      if (mr.n_inliers > g->curr_best.n_inliers) g->curr_best = mr;
      if (g->is_keyframe(mr.id1)) g->pend_edge_to_kf = true;
    }
  }
  const bool found_trafo = g->edges.size() != g->pend_edges_before;
  const bool keep_anyway = P.keep_all_nodes || (g->pend.n3d > P.min_matches && P.keep_good_nodes);
  // odom_frame_name is empty in every launch file: invalid_odometry is always true (:630-632)
  if (!found_trafo && keep_anyway) {
    EdgeRec e; e.id1 = g->pend_seq_prev; e.id2 = g->pend.id; e.info = -1.0;
    add_edge(g, e, true, true);
    g->pend.vertex_id = g->pending().vertex_id;
    g->put_pending_in_graph();
    g->nodes[(size_t)g->pend.id].valid_tf = false;
    g->curr_best = e;
  }
  finish_node(g, g->edges.size() > g->pend_edges_before, res);
}

}  // namespace

extern "C" void lsl_graph_params_default(lsl_graph_params* p) {
  if (!p) return;
  p->min_translation_meter = 0.0; p->min_rotation_degree = 0.0; p->max_translation_meter = 1e10; p->max_rotation_degree = 360;
  p->predecessor_candidates = 2; p->neighbor_candidates = 2; p->min_sampled_candidates = 2; p->geodesic_depth = 3;
  p->min_matches = 20; p->keep_all_nodes = 0; p->keep_good_nodes = 0; p->clear_non_keyframes = 0; p->clear_past_point_cloud = 1;
  p->largest_loop = 0; p->line_match_number_weight = 1;
}
extern "C" void lsl_graph_params_lineslam_launch(lsl_graph_params* p) {
  if (!p) return;
  lsl_graph_params_default(p);
  p->min_translation_meter = 0.01; p->min_rotation_degree = 0.1; p->predecessor_candidates = 1; p->neighbor_candidates = 0;
  p->min_sampled_candidates = 0; p->keep_all_nodes = 1; p->clear_non_keyframes = 1;
}

extern "C" int lsl_graph_create(lsl_graph** out, const lsl_graph_params* p, uint32_t seed) {
  if (!out) return LSL_ERR_ARG;
  lsl_graph* g = new lsl_graph();
  if (p) g->P = *p; else lsl_graph_params_default(&g->P);
  lslm::grand_seed(&g->rng, seed);
  *out = g;
  return LSL_OK;
}
extern "C" void lsl_graph_destroy(lsl_graph* g) { delete g; }

extern "C" int lsl_is_big_trafo(const double T[16], const lsl_graph_params* p) { return (T && p) ? (big_trafo(T, *p) ? 1 : 0) : LSL_ERR_ARG; }
extern "C" int lsl_is_small_trafo(const double T[16], double seconds, const lsl_graph_params* p) {
  return (T && p) ? (small_trafo(T, seconds, *p) ? 1 : 0) : LSL_ERR_ARG;
}

extern "C" int lsl_graph_potential_edge_targets(lsl_graph* g, int seq_t, int geo_t, int samp_t, int pred, int include_pred,
                                                int32_t* ids, int cap, int* n) {
  if (!g || !n || g->nodes.empty()) return LSL_ERR_ARG;
  if (pred >= (int)g->nodes.size()) return LSL_ERR_ARG;
  std::vector<int> v;
  edge_targets(g, seq_t, geo_t, samp_t, pred, include_pred != 0, &v);
  *n = (int)v.size();
  if (cap < *n || (!ids && *n)) return LSL_ERR_CAPACITY;
  for (int i = 0; i < *n; ++i) ids[i] = v[(size_t)i];
  return LSL_OK;
}

extern "C" int lsl_graph_node_begin(lsl_graph* g, double stamp, int n2d, int n3d, int* action, int* node_id, int* compare_with) {
  if (!g || !action || g->phase != 0) return LSL_ERR_ARG;
  Node nd; nd.stamp = stamp; nd.n2d = n2d; nd.n3d = n3d;
  if (node_id) *node_id = -1;
  if (compare_with) *compare_with = -1;
  if (g->nodes.empty()) {
    first_node(g, nd);
    if (node_id) *node_id = 0;
    *action = LSL_GRAPH_FIRST;
    return LSL_OK;
  }
  g->pend = nd; g->pend_active = true; g->pend_in_graph = false; g->pend_edge_to_kf = false; g->pend_pred_matched = false;
  g->pend_edges_before = g->edges.size();
  g->pend_cands.clear();
  if (n2d < g->P.min_matches && !g->P.keep_all_nodes) {     // nodeComparisons :428-434, then the tail of addNode
    lsl_graph_node_result r;
    finish_node(g, false, &r);
    *action = r.replaced_first ? LSL_GRAPH_FIRST : LSL_GRAPH_SKIPPED;
    if (node_id && r.replaced_first) *node_id = 0;
    return LSL_OK;
  }
  g->pend.id = (int)g->nodes.size();
  g->pend.seq_id = g->next_seq_id++;
  g->earliest_loop_closure_node = g->pend.id;
  g->pend_seq_prev = g->nodes.back().id;
  g->curr_best = EdgeRec();
  if (node_id) *node_id = g->pend.id;
  if (g->P.min_translation_meter > 0.0 || g->P.min_rotation_degree > 0.0) {
    if (compare_with) *compare_with = (int)g->nodes.size() - 1;
    *action = LSL_GRAPH_COMPARE_PREDECESSOR;
  } else *action = LSL_GRAPH_CANDIDATES;
  g->phase = 1;
  return LSL_OK;
}

extern "C" int lsl_graph_node_predecessor(lsl_graph* g, const lsl_pose_rec* rec, int* action, int32_t* ids, int cap, int* n,
                                          lsl_graph_node_result* res) {
  if (!g || !action || !n) return LSL_ERR_ARG;
  if (g->phase == 2 && !rec) {   // retry after LSL_ERR_CAPACITY: the candidates are cached, nothing is recomputed (no second rand() draw)
    *n = (int)g->pend_cands.size();
    *action = LSL_GRAPH_CANDIDATES;
    if (cap < *n || (!ids && *n)) return LSL_ERR_CAPACITY;
    for (int i = 0; i < *n; ++i) ids[i] = g->pend_cands[(size_t)i];
    return LSL_OK;
  }
  if (g->phase != 1) return LSL_ERR_ARG;
  const lsl_graph_params& P = g->P;
  *n = 0;
  if (rec && !rec_ids_ok(g, *rec, -1)) return LSL_ERR_ARG;   // nothing has been touched yet
  if (rec) {                                                  // initial comparison (:462-519)
    EdgeRec mr = edge_of(*rec, P);
    if (mr.id1 >= 0 && mr.id2 >= 0) {
      const Node& prev = g->nodes[(size_t)mr.id1];
      const double dt = g->pend.stamp - prev.stamp;
      if (!big_trafo(mr.tf.m, P) || !small_trafo(mr.tf.m, dt, P)) {
        g->curr_best = mr;
        finish_node(g, false, res);
        *action = LSL_GRAPH_DROPPED;
        return LSL_OK;
      }
      if (add_edge(g, mr, true, true)) {
        g->pend.vertex_id = g->pending().vertex_id;
        g->put_pending_in_graph();
        if (g->is_keyframe(mr.id1)) g->pend_edge_to_kf = true;
        g->nodes[(size_t)mr.id1].valid_tf = true;
        g->curr_best = mr;
      } else {
        finish_node(g, false, res);
        *action = LSL_GRAPH_DROPPED;
        return LSL_OK;
      }
      g->pend_pred_matched = true;
    }
  }
  const int seq_cand = P.predecessor_candidates - 1;
  if (g->pend_pred_matched) edge_targets(g, seq_cand, P.neighbor_candidates, P.min_sampled_candidates, g->curr_best.id1, false, &g->pend_cands);
  else edge_targets(g, seq_cand, P.neighbor_candidates, P.min_sampled_candidates, g->pend_seq_prev, true, &g->pend_cands);
  // (prev_best of :457-458 is read from a fresh MatchingResult: always -1, so the append at :533-535 never fires)
  *n = (int)g->pend_cands.size();
  g->phase = 2;
  *action = LSL_GRAPH_CANDIDATES;
  if (cap < *n || (!ids && *n)) return LSL_ERR_CAPACITY;
  for (int i = 0; i < *n; ++i) ids[i] = g->pend_cands[(size_t)i];
  return LSL_OK;
}

extern "C" int lsl_graph_node_commit(lsl_graph* g, const lsl_pose_rec* recs, int n, lsl_graph_node_result* res) {
  if (!g || g->phase != 2 || n != (int)g->pend_cands.size() || (n && !recs)) return LSL_ERR_ARG;
  for (int i = 0; i < n; ++i)   // validate every record before any state changes
    if (!rec_ids_ok(g, recs[i], g->pend_cands[(size_t)i])) return LSL_ERR_ARG;
  main_loop(g, recs, n, res);
  return LSL_OK;
}

extern "C" int lsl_graph_add_frame(lsl_graph* g, lsl_ctx* ctx, lsl_frame* frame, double stamp, int n2d, int n3d, uint32_t seed,
                                   lsl_graph_node_result* res) {
  if (!g || !ctx || !frame) return LSL_ERR_ARG;
  lsl_graph_node_result r; std::memset(&r, 0, sizeof r); r.node_id = -1; r.keyframe_added = -1; r.best_id1 = -1;
  int action = 0, nid = -1, cmp = -1, rc;
  auto release_lines = [&]() {   // the clear_past_point_cloud sweep frees the line vectors of all but the newest node
    for (Node& nd : g->nodes)
      if (!nd.has_lines && nd.frame && lsl_frame_num_lines(nd.frame) > 0) lsl_frame_clear_lines(nd.frame);
  };
  if ((rc = lsl_graph_node_begin(g, stamp, n2d, n3d, &action, &nid, &cmp)) != LSL_OK) return rc;
  if (action == LSL_GRAPH_FIRST) {
    g->nodes[0].frame = frame;
    r.found_match = 1; r.node_id = 0; r.in_graph = 1; r.keyframe_added = 0;
    if (res) *res = r;
    return LSL_OK;
  }
  if (action == LSL_GRAPH_SKIPPED) { if (res) *res = r; release_lines(); return LSL_OK; }
  g->pend.frame = frame;
  lsl_pose_rec prec;
  const lsl_pose_rec* pp = nullptr;
  if (action == LSL_GRAPH_COMPARE_PREDECESSOR) {
    const lsl_frame* q = frame; const lsl_frame* t = g->nodes[(size_t)cmp].frame;
    if (!t) { g->phase = 0; g->pend_active = false; return LSL_ERR_ARG; }
    int32_t iq = nid, it = cmp;
    if ((rc = lsl_match_pair_batch(ctx, 1, &q, &t, &iq, &it, &seed, &prec)) != LSL_OK) { g->phase = 0; g->pend_active = false; return rc; }
    pp = &prec;
  }
  std::vector<int32_t> ids(g->nodes.size() + 8);
  int n = 0;
  if ((rc = lsl_graph_node_predecessor(g, pp, &action, ids.data(), (int)ids.size(), &n, &r)) != LSL_OK) return rc;
  if (action == LSL_GRAPH_DROPPED) { if (res) *res = r; release_lines(); return LSL_OK; }
  std::vector<lsl_pose_rec> recs((size_t)n);
  int fail_rc = LSL_OK;
  if (n) {
    std::vector<const lsl_frame*> qs((size_t)n, frame), ts((size_t)n);
    std::vector<int32_t> iq((size_t)n, nid), it((size_t)n);
    std::vector<uint32_t> sd((size_t)n);
    for (int i = 0; i < n; ++i) {
      const Node& c = (ids[(size_t)i] == nid) ? g->pending() : g->nodes[(size_t)ids[(size_t)i]];
      ts[(size_t)i] = c.frame; it[(size_t)i] = ids[(size_t)i]; sd[(size_t)i] = seed + 1u + (uint32_t)i;
      if (!c.frame) {   // node inserted through the three-phase calls: no frame to register against
        fail_rc = LSL_ERR_ARG; break;
      }
    }
    if (fail_rc == LSL_OK) fail_rc = lsl_match_pair_batch(ctx, n, qs.data(), ts.data(), iq.data(), it.data(), sd.data(), recs.data());
    if (fail_rc != LSL_OK) {
      // the predecessor edge may already be in the graph: finish the node with the records obtained so far (none of
      // the candidates registered) so that keyframe / clear_past_point_cloud bookkeeping still runs, then report
      std::memset(recs.data(), 0, sizeof(lsl_pose_rec) * (size_t)n);
      lsl_graph_node_commit(g, recs.data(), n, &r);
      if (res) *res = r;
      release_lines();
      return fail_rc;
    }
  }
  rc = lsl_graph_node_commit(g, recs.data(), n, &r);
  if (res) *res = r;
  release_lines();
  return rc;
}

extern "C" int lsl_graph_num_nodes(const lsl_graph* g) { return g ? (int)g->nodes.size() : LSL_ERR_ARG; }

extern "C" int lsl_graph_nodes(const lsl_graph* g, lsl_graph_node* dst, int cap, int* n) {
  if (!g || !n) return LSL_ERR_ARG;
  *n = (int)g->nodes.size();
  if (cap < *n || (!dst && *n)) return LSL_ERR_CAPACITY;
  for (int i = 0; i < *n; ++i) {
    const Node& s = g->nodes[(size_t)i];
    lsl_graph_node& d = dst[i];
    d.id = s.id; d.seq_id = s.seq_id; d.vertex_id = s.vertex_id; d.matchable = s.matchable; d.valid_tf_estimate = s.valid_tf;
    d.has_lines = s.has_lines; d.n_feat2d = s.n2d; d.n_feat3d = s.n3d; d.stamp = s.stamp;
    const Iso e = s.vertex_id >= 0 ? g->vest[(size_t)s.vertex_id] : iso_identity();
    std::memcpy(d.estimate, e.m, sizeof e.m);
  }
  return LSL_OK;
}

extern "C" int lsl_graph_edges(const lsl_graph* g, lsl_graph_edge* dst, int cap, int* n) {
  if (!g || !n) return LSL_ERR_ARG;
  *n = (int)g->edges.size();
  if (cap < *n || (!dst && *n)) return LSL_ERR_CAPACITY;
  for (int i = 0; i < *n; ++i) {
    const EdgeRec& s = g->edges[(size_t)i];
    dst[i].id1 = s.id1; dst[i].id2 = s.id2; dst[i].n_inliers = s.n_inliers; dst[i].pad = 0; dst[i].info = s.info;
    std::memcpy(dst[i].transform, s.tf.m, sizeof s.tf.m);
  }
  return LSL_OK;
}

extern "C" int lsl_graph_keyframes(const lsl_graph* g, int32_t* dst, int cap, int* n) {
  if (!g || !n) return LSL_ERR_ARG;
  *n = (int)g->keyframes.size();
  if (cap < *n || (!dst && *n)) return LSL_ERR_CAPACITY;
  for (int i = 0; i < *n; ++i) dst[i] = g->keyframes[(size_t)i];
  return LSL_OK;
}

// write_poses_2file (graph_manager.cpp:864-884) with r2q (src/line/utils.cpp:1709-1720); ofstream precision(16) in the
// default float notation is printf's %.16g
extern "C" int lsl_graph_write_poses(const lsl_graph* g, const char* filename) {
  if (!g || !filename) return LSL_ERR_ARG;
  FILE* f = std::fopen(filename, "w");
  if (!f) return LSL_ERR_ARG;
  for (const Node& nd : g->nodes) {
    if (!nd.valid_tf || nd.vertex_id < 0) continue;
    const double* T = g->vest[(size_t)nd.vertex_id].m;
    const double t = T[0] + T[5] + T[10];
    const double r = std::sqrt(1 + t);
    const double s = 0.5 / r;
    const double w = 0.5 * r, x = (T[9] - T[6]) * s, y = (T[2] - T[8]) * s, z = (T[4] - T[1]) * s;
    std::fprintf(f, "%.16g\t%.16g\t%.16g\t%.16g\t%.16g\t%.16g\t%.16g\t%.16g\n", nd.stamp, T[3], T[7], T[11], x, y, z, w);
  }
  std::fclose(f);
  return LSL_OK;
}
