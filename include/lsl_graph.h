/* lsl_graph.h — C ABI of the caller of the hot path ("next"): candidate selection and edge
 * bookkeeping of GraphManager, feeding lsl_match_pair_batch (include/lsl.h) and consuming its 128-byte records.
 * Host-side integer / small-matrix logic (no kernels): it exists so that pair throughput becomes trajectory
 * throughput — one lsl_match_pair_batch call per inserted frame instead of one matchNodePair per candidate.
 * The g2o optimiser itself stays with the reference ("g2o back end untouched"): vertex estimates here are the
 * chained edge transforms addEdgeToG2O sets (src/graph_manager.cpp:961-975), edges are handed out for g2o.
 *
 * Reference symbols replaced (all under /root/reference/src):
 *   isBigTrafo / isSmallTrafo (Eigen::Isometry3d)        misc.cpp:260-296          lsl_is_big_trafo / lsl_is_small_trafo
 *   GraphManager::getPotentialEdgeTargetsWithDijkstra    graph_manager.cpp:204-319  lsl_graph_potential_edge_targets
 *   GraphManager::addNode / firstNode / nodeComparisons  graph_manager.cpp:730-860, 358-400, 419-708
 *                                                        lsl_graph_node_begin / _predecessor / _commit, lsl_graph_add_frame
 *   GraphManager::addKeyframe, addEdgeToG2O              graph_manager.cpp:901-926, 928-1006 (inside the calls above)
 *   GraphManager::write_poses_2file                      graph_manager.cpp:864-884  lsl_graph_write_poses (TUM trajectory)
 * rand(): replayed from the seed given to lsl_graph_create (one stream per graph, serial order of the reference).
 */
#ifndef LSL_GRAPH_H_
#define LSL_GRAPH_H_
#include "lsl.h"
#ifdef __cplusplus
extern "C" {
#endif

/* src/parameter_server.cpp:82-149,196 (defaults) */
typedef struct lsl_graph_params {
  double min_translation_meter, min_rotation_degree, max_translation_meter;
  int32_t max_rotation_degree;
  int32_t predecessor_candidates, neighbor_candidates, min_sampled_candidates, geodesic_depth;
  int32_t min_matches, keep_all_nodes, keep_good_nodes, clear_non_keyframes, clear_past_point_cloud;
  int32_t largest_loop;              /* pose_relative_to == "largest_loop" */
  int32_t line_match_number_weight;  /* sysPara.line_match_number_weight (edge information, node.cpp:1531) */
} lsl_graph_params;

typedef struct lsl_graph_node {
  int32_t id, seq_id, vertex_id, matchable, valid_tf_estimate, has_lines, n_feat2d, n_feat3d;
  double stamp;
  double estimate[16]; /* row-major Isometry3d of the node's vertex (identity if it has none) */
} lsl_graph_node;

typedef struct lsl_graph_edge {
  int32_t id1, id2, n_inliers, pad;
  double info;          /* informationMatrix = info * I6; -1: the constant-position edge (graph_manager.cpp:662-672) */
  double transform[16]; /* row-major Isometry3d, maps id2 (newer) coordinates into id1 (older) */
} lsl_graph_edge;

enum {
  LSL_GRAPH_FIRST = 0,               /* node became the first node; nothing else to call */
  LSL_GRAPH_SKIPPED = 1,             /* too few features (graph_manager.cpp:428-434); nothing else to call */
  LSL_GRAPH_COMPARE_PREDECESSOR = 2, /* register (node, *compare_with), then lsl_graph_node_predecessor(rec) */
  LSL_GRAPH_CANDIDATES = 3,          /* register the node against ids[0..n), then lsl_graph_node_commit */
  LSL_GRAPH_DROPPED = 4              /* predecessor transformation out of bounds / edge refused; node not added */
};

typedef struct lsl_graph_node_result {
  int32_t found_match;   /* return value of addNode */
  int32_t node_id;       /* id the node got (-1 if it never received one) */
  int32_t in_graph;      /* node is part of graph_ afterwards */
  int32_t edges_added, keyframe_added /* id or -1 */, best_id1 /* curr_best_result_.edge.id1 */, replaced_first;
  int32_t n_candidates;
} lsl_graph_node_result;

typedef struct lsl_graph lsl_graph;

void lsl_graph_params_default(lsl_graph_params* p);          /* parameter_server.cpp defaults */
void lsl_graph_params_lineslam_launch(lsl_graph_params* p);  /* launch/lineslam.launch:15-36 overrides */
int lsl_graph_create(lsl_graph** out, const lsl_graph_params* p, uint32_t seed);
void lsl_graph_destroy(lsl_graph* g);

int lsl_is_big_trafo(const double T[16], const lsl_graph_params* p);
int lsl_is_small_trafo(const double T[16], double seconds, const lsl_graph_params* p);

/* getPotentialEdgeTargetsWithDijkstra on the current graph (consumes rand() like the reference). */
int lsl_graph_potential_edge_targets(lsl_graph* g, int sequential_targets, int geodesic_targets, int sampled_targets,
                                     int predecessor_id, int include_predecessor, int32_t* ids, int cap, int* n);

/* addNode in three phases, so that the registrations run as ONE lsl_match_pair_batch per phase. */
int lsl_graph_node_begin(lsl_graph* g, double stamp, int n_feat2d, int n_feat3d, int* action, int* node_id,
                         int* compare_with);
int lsl_graph_node_predecessor(lsl_graph* g, const lsl_pose_rec* rec /* NULL after LSL_GRAPH_CANDIDATES */, int* action,
                               int32_t* ids, int cap, int* n, lsl_graph_node_result* res /* filled on DROPPED */);
int lsl_graph_node_commit(lsl_graph* g, const lsl_pose_rec* recs, int n, lsl_graph_node_result* res);

/* The three phases driven against the device: frames of earlier nodes are kept by the graph (borrowed pointers; the
 * caller frees them after lsl_graph_destroy). seed -> per-pair seeds seed + k in candidate order. Nodes whose lines
 * were released by the clear_past_point_cloud sweep are cleared with lsl_frame_clear_lines first, like the reference. */
int lsl_graph_add_frame(lsl_graph* g, lsl_ctx* ctx, lsl_frame* frame, double stamp, int n_feat2d, int n_feat3d,
                        uint32_t seed, lsl_graph_node_result* res);

int lsl_graph_num_nodes(const lsl_graph* g);
int lsl_graph_nodes(const lsl_graph* g, lsl_graph_node* dst, int cap, int* n);
int lsl_graph_edges(const lsl_graph* g, lsl_graph_edge* dst, int cap, int* n);
int lsl_graph_keyframes(const lsl_graph* g, int32_t* dst, int cap, int* n);
/* write_poses_2file: "ts\ttx\tty\ttz\tqx\tqy\tqz\tqw\n", precision 16, nodes without a valid estimate skipped. */
int lsl_graph_write_poses(const lsl_graph* g, const char* filename);

#ifdef __cplusplus
}
#endif
#endif
