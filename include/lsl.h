/* lsl.h — C ABI of the B200-native LineSLAM line front end (liblsl_b200.so).
 *
 * Drop-in boundary for the reference's hot path. The reference has no FFI;
 * these entry points are what a maintainer binds from the C++ symbols listed beside each
 * function (see INTEGRATION.md for the adapter code). POD only, no exceptions, no exit():
 * every call returns 0 or a negative lsl_status.
 *
 * Threading: every entry point may be called from any host thread. Calls on ONE context are serialised
 * by a per-context lock in arrival order (the reference calls detect3DLines on a QtConcurrent worker while QThreadPool
 * threads run matchNodePair, src/node.cpp:214, src/graph_manager.cpp:555 — those callers work unchanged and get the
 * results of a serial run); "last call" accessors (lsl_pair_matches, lsl_kernel_times, lsl_last_timing) refer to the
 * last call that completed on the context, so a thread that needs them keeps its pair calls on a context of its own.
 * Host threads that should overlap on the device use one context each (contexts share nothing but the device).
 *
 * RNG contract: the process-global rand() of the reference becomes an explicit
 * seed per call; the library replays glibc's TYPE_3 generator from that seed in the serial
 * order of a single-threaded reference run.
 */
#ifndef LSL_H_
#define LSL_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum lsl_status {
  LSL_OK = 0,
  LSL_ERR_ARG = -1,        /* NULL / out-of-range argument */
  LSL_ERR_CUDA = -2,       /* CUDA runtime error (lsl_last_error has the text) */
  LSL_ERR_CAPACITY = -3,   /* caller buffer or internal table too small */
  LSL_ERR_NO_DEVICE = -4,  /* no CUDA device: there is NO CPU fallback */
  LSL_ERR_NCCL = -5,
  LSL_ERR_BUSY = -6        /* a pair batch is in flight (lsl_match_pair_batch_begin): finish it with lsl_match_pair_batch_end first */
} lsl_status;

/* Parameters: src/parameter_server.cpp:160-199 + SystemParameters::init (src/line/lineslam.cpp:577-640)
 * + the LSD constants of lsd_scale() (external/lsd/lsd.cpp:2070-2091). Layout is shared with the
 * test oracle; keep doubles first, then ints. */
typedef struct lsl_params {
  double lsd_scale, lsd_sigma_scale, lsd_quant, lsd_ang_th, lsd_eps, lsd_density_th, lsd_max_grad;
  double line_2d_len_thres, msld_sample_interval, line_3d_len_thres_m, collin_pts_ratio, line_sample_interval;
  double pt2line_mahdist_extractline, ratio_support_pts_on_line, stdev_sample_pt_imgline;
  double depth_stdev_coeff_c1, depth_stdev_coeff_c2, depth_stdev_coeff_c3, depth_scaling;
  double max_mah_dist_for_inliers, g2o_line_error_weight, g2o_BA_kernel_delta;
  double pt2line3d_dist_relmotion, line3d_angle_relmotion;
  double sigma_depth, nn_distance_ratio;   /* point features: src/parameter_server.cpp:45,146 */
  int32_t lsd_n_bins, line_sample_max_num, line_sample_min_num, line3d_mle_iter_num;
  int32_t ransac_iters_extract_line, num_cells_lineseg_range;
  int32_t ransac_iters_line_motion, adjacent_linematch_window, line_match_number_weight;
  int32_t min_feature_matches, min_matches_loopclose, g2o_BA_use_kernel;
} lsl_params;

/* One 3D line of a frame: FrameLine + RandomLine3d (src/line/lineslam.h:84-151). 1040 bytes. */
typedef struct lsl_line_rec {
  double p[2], q[2], lineEq2d[3], r[2], des[72];
  double A[3], B[3], covA[9], covB[9], DU_A[9], DU_B[9], Wsqrt_A[3], Wsqrt_B[3];
  int32_t lid, haveDepth;
} lsl_line_rec;

/* cv::DMatch subset used by the path */
typedef struct lsl_match { int32_t queryIdx, trainIdx; float distance; } lsl_match;

/* Result of one pair registration; fixed 128-byte record (the unit of lsl_allgather_poses).
 * tf maps query(newer) -> train(older) coordinates, row-major Matrix4f (motion.cpp:534). */
typedef struct lsl_pose_rec {
  int32_t id_train, id_query, found, n_line_matches, n_ransac_inliers, n_inliers;
  float rmse;
  float tf[16];
  int32_t best_iter;
  int32_t pad[8];
} lsl_pose_rec;

typedef struct lsl_ctx lsl_ctx;     /* owns device, stream, scratch */
typedef struct lsl_frame lsl_frame; /* library-owned: device-resident line records + host mirror */

void lsl_params_default(lsl_params* p);
const char* lsl_strerror(int status);
const char* lsl_last_error(const lsl_ctx* ctx);

/* max_batch frames / pairs can be in flight per call; scratch is sized at creation. */
int lsl_ctx_create(lsl_ctx** out, const lsl_params* params, int cuda_device, int max_batch, int max_w, int max_h);
void lsl_ctx_destroy(lsl_ctx* ctx);
/* Run every call of this context on the caller's cudaStream_t (NULL restores the context's own stream). */
int lsl_ctx_set_stream(lsl_ctx* ctx, void* cuda_stream);
/* Debug mode keeps the LSD rows and per-line intermediates of extracted frames on the host
 * (lsl_frame_segments, lsl_frame_debug); off by default: the product path moves 8 bytes per frame. */
int lsl_ctx_set_debug(lsl_ctx* ctx, int on);

/* Node::detect3DLines (src/node.h:286-287, src/line/lineslam.cpp:200-357) for one frame.
 * img: u8, channels = 1 (gray, as detect3DLines receives it) or 3 (interleaved, memory-order
 * CV_RGB2GRAY like src/node.cpp:191-196). depth: f32 metres*depth_scaling, 0/NaN invalid.
 * K: row-major 3x3 (global K of src/node.cpp:200-206). asynch_dt: Node::asynch_time_diff_sec_. */
int lsl_extract(lsl_ctx* ctx, const uint8_t* img, int channels, const float* depth, int W, int H,
                const double K[9], double asynch_dt_s, uint32_t rand_seed, lsl_frame** out);
/* Same for n frames in one launch sequence (the Node constructors of a stream). Host buffers:
 * imgs[i] / depths[i] point to frame i. */
int lsl_extract_batch(lsl_ctx* ctx, int n, const uint8_t* const* imgs, int channels,
                      const float* const* depths, int W, int H, const double K[9],
                      double asynch_dt_s, const uint32_t* rand_seeds, lsl_frame** out);
/* Host buffers with the depth map as the sensor / the TUM PNG delivers it: 16-bit, 0 = no measurement. The conversion
 * of src/openni_listener.cpp:1233-1244 (convertTo(CV_32FC1); values < 1e-5 -> NaN; "/ depth_factor" = float multiply by
 * (float)(1 / depth_factor), 5000 for TUM) runs on the device, so a frame costs W*H*(channels + 2) bytes of PCIe instead of
 * W*H*(channels + 4). Results are identical to lsl_extract_batch on the converted planes. */
int lsl_extract_batch_u16(lsl_ctx* ctx, int n, const uint8_t* const* imgs, int channels, const uint16_t* const* depths,
                          int W, int H, const double K[9], double asynch_dt_s, const uint32_t* rand_seeds,
                          double depth_factor, lsl_frame** out);
/* Inputs already resident in device memory: d_imgs = n*H*W*channels u8, d_depths = n*H*W f32. */
int lsl_extract_batch_dev(lsl_ctx* ctx, int n, const uint8_t* d_imgs, int channels, const float* d_depths,
                          int W, int H, const double K[9], double asynch_dt_s, const uint32_t* rand_seeds,
                          lsl_frame** out);

int lsl_frame_num_lines(const lsl_frame* f);
int lsl_frame_lines(const lsl_frame* f, lsl_line_rec* dst, int cap, int* n);
/* LSD output (ntuple_list of lsd(), external/lsd/lsd.h): n x 5 doubles x1,y1,x2,y2,width.
 * Only frames extracted in debug mode keep the rows; otherwise *n is set and LSL_ERR_ARG returned. */
int lsl_frame_segments(const lsl_frame* f, double* dst, int cap, int* n);
/* Builds a frame from caller-supplied records (e.g. features cached by the host). */
int lsl_frame_from_lines(lsl_ctx* ctx, const lsl_line_rec* recs, int n, lsl_frame** out);
void lsl_frame_free(lsl_frame* f);
/* vector<FrameLine>().swap(node->lines) of the clear_past_point_cloud sweep (src/graph_manager.cpp:845-857): the
 * frame keeps its point features but has no lines afterwards (device records released). */
int lsl_frame_clear_lines(lsl_frame* f);

/* Node::lineMatching (src/node.h:288, src/node.cpp:1619-1694): query = this, train = other. */
int lsl_match_lines(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, int adjacent,
                    lsl_match* out, int cap, int* n);

/* Point features of a frame: Node::feature_locations_3d_ (Eigen::Vector4f rows x,y,z,1; z may be NaN) and
 * Node::feature_descriptors_ (CV_32F rows, after squareroot_descriptor_space, src/node.cpp:304-310, 1823-1837).
 * They are INPUTS of the path (the detectors are out of scope); copied to the device, n <= 2048. */
int lsl_frame_set_points(lsl_ctx* ctx, lsl_frame* f, const float* xyz1, const float* desc, int n, int dim);
/* Same with the two conditioning choices of the reference's Node constructor behind the ABI:
 *   desc_is_u8 = 1: binary rows (ORB, 32 bytes: src/features.cpp:174-211); featureMatching then runs the
 *                   "BruteForce-HammingLUT" matcher (src/node.cpp:609-613): popcount distances as float, same k = 2 /
 *                   ratio / unique-train / rand() jitter pass. dim = bytes per row (multiple of 4).
 *   root_sift = 1:  squareroot_descriptor_space (src/node.cpp:304-310, 1823-1837) is applied ON THE DEVICE to the f32 rows
 *                   (abs -> L1 normalise by the row's sequential float sum -> sqrt; zero rows untouched); pass the raw
 *                   SIFT / SURF rows. Ignored for u8 rows (the reference does not condition ORB). */
int lsl_frame_set_points_ex(lsl_ctx* ctx, lsl_frame* f, const float* xyz1, const void* desc, int n, int dim, int desc_is_u8,
                            int root_sift);
/* The same for n frames at once (the Node constructors of a batch): xyz1 / desc hold the frames' rows back to back,
 * counts[i] rows for frames[i]. One device block and two uploads for the whole batch. */
int lsl_frames_set_points_batch(lsl_ctx* ctx, int n, lsl_frame* const* frames, const float* xyz1, const void* desc,
                                const int32_t* counts, int dim, int desc_is_u8, int root_sift);
/* Copies the frame's (conditioned) descriptor rows back: n x dim floats, or n x dim bytes for u8 rows. */
int lsl_frame_descriptors(lsl_ctx* ctx, const lsl_frame* f, void* dst, int64_t cap_bytes);
/* The point-feature half of Node::Node on the device (src/node.cpp:219-310, 952-1018): when a detector
 * is selected, every lsl_extract* call also runs detector->detect (SIFT as OpenCV implements it), removeDepthless,
 * KeyPointsFilter::retainBest(max_keypoints), extractor->compute, projectTo3D and (root_sift) squareroot_descriptor_space on
 * the frames' gray / depth planes already in HBM and attaches the result as the frames' point features — what
 * lsl_frame_set_points would upload, without a host round trip. kind: 0 none (default), 1 SIFT. Needs 3-channel or gray
 * input through the extract call; Tier-T against cv2's SIFT (tests/test_gpu_sift.py). */
int lsl_ctx_set_point_detector(lsl_ctx* ctx, int kind, int max_keypoints, int root_sift);
/* Read-back of a frame's point features: xyz1 [n][4], desc [n][dim] (f32 rows), kp [n][6] = x, y, size, angle, response,
 * octave + 256 * layer (device-detected points only); any of the three may be NULL. */
int lsl_frame_points(lsl_ctx* ctx, const lsl_frame* f, float* xyz1, float* desc, float* kp, int cap, int* n);
/* Node::computeInliersAndError (src/node.cpp:1019-1080; compiled but not called when the reference is built with USE_LINES — the
 * hybrid RANSAC scores points itself — provided because ICP / point-only builds call it): inliers of `matches` (queryIdx into
 * `query`'s points, trainIdx into `train`'s) under the float transform `tf` (row-major 4 x 4) with errorFunction2 <=
 * squared_max_inlier_dist, in match order; *rmse = sqrt(mean squared Mahalanobis distance), 1e9 with fewer than 3 inliers. */
int lsl_compute_inliers_and_error(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, const lsl_match* matches, int n,
                                  const float tf[16], double squared_max_inlier_dist, lsl_match* inliers, int cap, int* n_inliers,
                                  double* rmse);
int lsl_frame_num_points(const lsl_frame* f);
/* fx = K(0,0) and Node::asynch_time_diff_sec_ used by the point-edge information matrices of the refinement
 * (compPt3dCov, src/transformation_estimation.cpp:243-262). Every extract call sets them from its K / dt. */
int lsl_ctx_set_camera(lsl_ctx* ctx, double fx, double asynch_dt_s);

/* Node::featureMatching, matcher_type BRUTEFORCE (src/node.h:139, src/node.cpp:606-641): L2 k = 2 nearest
 * neighbours, ratio < nn_distance_ratio, unique trainIdx, distance = ratio + rand()/(1000 RAND_MAX). */
int lsl_match_points(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, uint32_t seed, lsl_match* out,
                     int cap, int* n);

/* getTransform_PtsLines_ransac (src/line/utils.h:147-153, src/line/motion.cpp:605-849) with point and/or line
 * matches. inliers_out receives output_line_inlier_matches; ransac_inliers_out (optional) the
 * max_line_inlier_set of the best hypothesis; the point lists are read with lsl_pair_matches(ctx, 0, 3..5).
 * With point matches the rand() stream starts after npt draws (the featureMatching jitter of the same
 * matchNodePair call). rec->pad[0..2] = #point matches, #point inliers of the best hypothesis, #refined. */
int lsl_pose_ransac(lsl_ctx* ctx, const lsl_frame* train, const lsl_frame* query, int id_train, int id_query,
                    const lsl_match* pt_matches, int npt, const lsl_match* ln_matches, int nln, uint32_t seed,
                    lsl_pose_rec* rec, lsl_match* inliers_out, int cap, int* n_inl,
                    lsl_match* ransac_inliers_out, int cap2, int* n_rinl);

/* computeRelativeMotion_Ransac (src/line/utils.h, src/line/motion.cpp:367-526): line-only RANSAC on the matched
 * line pairs (Euclidean consensus: pt2line3d_dist_relmotion, line3d_angle_relmotion) followed by
 * optimizeRelmotion (motion.cpp:98-139: dlevmar_dif on quaternion + translation, 50 iterations) and consensus
 * growing. R (row-major 3x3), t: x_train = R x_query + t. conset receives the indices into ln_matches.
 * *have = 0 when no hypothesis has support (R, t untouched). */
int lsl_relmotion_ransac(lsl_ctx* ctx, const lsl_frame* train, const lsl_frame* query, const lsl_match* ln_matches, int nln,
                         uint32_t seed, double R[9], double t[3], int32_t* conset, int cap, int* n_conset, int* lm_calls,
                         int* have);

/* lsl_relmotion_ransac for every pair of the last lsl_match_pair_batch, on the line matches that call left on the device
 * ("levmar LM refine per edge", BASELINE config 5): Rt[i] = R (9, row-major) then t (3), info[i] = {consensus size,
 * optimizeRelmotion calls, have, 0}; seeds are the pairs' seeds. One launch. */
int lsl_relmotion_batch(lsl_ctx* ctx, int npairs, double* Rt, int32_t* info);

/* Node::matchNodePair (src/node.h:107, src/node.cpp:1494-1545) for npairs independent pairs, the
 * unit GraphManager::nodeComparisons maps over (src/graph_manager.cpp:555): featureMatching (when both frames
 * carry point features) + lineMatching + pose RANSAC + refinement on the device, one 128-byte record per pair back. */
int lsl_match_pair_batch(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                         const int32_t* id_query, const int32_t* id_train, const uint32_t* seeds,
                         lsl_pose_rec* out);

/* The same batch in two halves, for a stream of batches: _begin enqueues matching + registration of the pairs on the
 * context's PAIR STREAM and returns at once; _end waits for it and delivers the records. Between the two the caller may
 * run lsl_extract* of the NEXT batch on the same context: its image / LSD / 3D-line kernels (latency-bound, few warps per
 * SM) then share the GPU with the registration of the previous batch (register-bound, 16 warps per SM) instead of
 * running after it — the two stages touch disjoint workspaces. While a batch is in flight every other entry point that
 * uses the pair workspace returns LSL_ERR_BUSY, and the frames of the batch must stay alive (lsl_frame_free of any frame
 * waits for the pair stream first). lsl_match_pair_batch == _begin immediately followed by _end. */
int lsl_match_pair_batch_begin(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                               const int32_t* id_query, const int32_t* id_train, const uint32_t* seeds);
int lsl_match_pair_batch_end(lsl_ctx* ctx, lsl_pose_rec* out, int cap);
/* Match lists of pair `pair` of the last lsl_match_pair_batch / lsl_pose_ransac call: what = 0 all line
 * matches (Node::lineMatching output), 1 refined inliers (output_line_inlier_matches), 2 inliers of the best
 * RANSAC hypothesis (max_line_inlier_set, motion.cpp:714-721); 3, 4, 5 the same three lists for points
 * (Node::featureMatching output, output_point_inlier_matches, max_point_inlier_set). */
int lsl_pair_matches(lsl_ctx* ctx, int pair, int what, lsl_match* out, int cap, int* n);

/* Graph-insert-time exchange: all ranks contribute nlocal records and receive
 * nranks*nlocal. nccl_comm is an ncclComm_t created by the host; NCCL is resolved with dlopen. local_recs == NULL: the
 * first nlocal records of the last lsl_match_pair_batch are sent straight from the device buffer the pose kernel wrote. */
int lsl_allgather_poses(lsl_ctx* ctx, void* nccl_comm, int nranks, const lsl_pose_rec* local_recs, int nlocal,
                        lsl_pose_rec* all_recs);
/* Communicator owned by the library (the reference has no communication layer; this is the plumbing the
 * sharded deployment needs): rank 0 calls lsl_comm_unique_id (128-byte ncclUniqueId), the host ships it
 * to the other ranks by any means, every rank calls lsl_comm_init; lsl_allgather_poses(ctx, NULL, 0, ...)
 * then uses it. */
int lsl_comm_unique_id(void* id128);
int lsl_comm_init(lsl_ctx* ctx, const void* id128, int nranks, int rank);

/* Loop-closure batches sharded over ranks (BASELINE config 4): the keyframes' features are pre-distributed
 * block-wise, the ONE query frame's line records (~0.27 MB) are broadcast from `root` to every rank with ncclBroadcast over
 * the library's communicator (lsl_comm_init). On the root `frame` is the extracted query and *out == frame; on the other
 * ranks `frame` is ignored and *out is a new frame holding the records (free it with lsl_frame_free). Point features are
 * not carried. Device to device: the records never visit the host. */
int lsl_bcast_frame(lsl_ctx* ctx, int root, lsl_frame* frame, lsl_frame** out);
/* One stream split block-wise over the ranks (the map of src/graph_manager.cpp:555 sharded): the pair
 * (t-1, t) at the head of rank r's block needs the record of frame t-1, which rank r-1 extracted as the tail of its block.
 * Every rank sends `frame` (its tail) to rank (r + 1) % nranks and receives the tail of rank (r - 1 + nranks) % nranks in
 * *out (a new frame; free it with lsl_frame_free): ncclSend / ncclRecv in one group over the library's communicator,
 * peer to peer over NVLink, device to device. Rank 0 receives the tail of the last rank, i.e. the predecessor of the head
 * of its NEXT block. Line records only (point features are not carried). Collective: every rank must call it. */
int lsl_shift_frame(lsl_ctx* ctx, const lsl_frame* frame, lsl_frame** out);

/* Stage counters of the last call (segments, lines, matches, LM iterations, kernel launches). */
typedef struct lsl_stats {
  int64_t kernel_launches, frames, segments, lines3d, pairs, matches, h2d_bytes, d2h_bytes;
} lsl_stats;
int lsl_get_stats(const lsl_ctx* ctx, lsl_stats* out);
/* Device-time of the region-growing kernel of the last extract call in ms (CUDA events on the
 * context stream), and of the whole last call. */
int lsl_last_timing(const lsl_ctx* ctx, float* ms_total, float* ms_region_grow);
/* Device time (ms, CUDA events on the context stream) of every kernel launched by the last extract call
 * and the last pair call; lsl_kernel_name(i) names entry i. */
int lsl_kernel_times(const lsl_ctx* ctx, float* ms, int cap, int* n);
const char* lsl_kernel_name(int i);

/* Tensor-core pre-filter of lsl_match_points / the point half of lsl_match_pair_batch (f32 rows of 8..128 floats, multiple
 * of 8; LSL_MATCH_TC=0 in the environment selects the exact scalar kernel): out[0] exact distance evaluations made by the
 * refine step, out[1] query rows that overflowed their candidate list and were rescanned in full, out[2] query rows. */
int lsl_match_tc_stats(lsl_ctx* ctx, int64_t out[3], int reset);

/* Stage-wise read-back for parity tests (frame 0 of the last extract call):
 * what: 0 gray u8[H*W], 1 scaled f64[sh*sw], 2 angles f64, 3 modgrad f64, 4 seeds i32 (x|y<<16),
 *       5 gx i16[H*W], 6 gy i16[H*W]. Returns the number of elements copied (or <0). */
int64_t lsl_debug_read(lsl_ctx* ctx, int what, void* dst, int64_t cap_bytes);
/* Per-line intermediates of a frame for parity tests: 3D-RANSAC inlier counts [n], inlier sample
 * indices [n][101], LSD segment index of each kept line [n], levmar iteration counts [n]. */
int lsl_frame_debug(const lsl_frame* f, int32_t* npts, int32_t* inl_idx, int32_t* seg_of_line, int32_t* lm_iters);

#ifdef __cplusplus
}
#endif
#endif /* LSL_H_ */
