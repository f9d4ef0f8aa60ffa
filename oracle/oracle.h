// oracle.h — CPU restatement of the LineSLAM line front end (TEST INFRASTRUCTURE).
//
// This directory is the parity checker for the CUDA path. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it. The product library (lineslam_b200/csrc) never links or calls it.
//
// Each function cites the reference file:line it restates (paths relative to
// /root/reference). Arithmetic is IEEE double, built with -ffp-contract=off
// (the reference builds with -march=corei7-avx: AVX, no FMA).
// Transcendentals come from csrc/shared/lsl_math.h so that CPU and GPU agree
// bit for bit; oracle/_ref (the unmodified upstream lsd.c + levmar built with
// glibc libm) pins the restatement — see oracle/Makefile and tests/test_oracle_ref.py.
#pragma once
#include <stdint.h>
#include <vector>

namespace orc {

// Appendix B of SURVEY.md: src/parameter_server.cpp:160-199, src/line/lineslam.cpp:577-640.
struct Params {
  // LSD (external/lsd/lsd.cpp:2070-2097)
  double lsd_scale = 0.8, lsd_sigma_scale = 0.6, lsd_quant = 2.0, lsd_ang_th = 22.5,
         lsd_eps = 0.0, lsd_density_th = 0.7, lsd_max_grad = 255.0;
  int lsd_n_bins = 1024;
  // 2D / 3D line extraction
  double line_2d_len_thres = 10.0, msld_sample_interval = 1.0, line_3d_len_thres_m = 0.02,
         collin_pts_ratio = 0.6, line_sample_interval = 1.0;
  int line_sample_max_num = 100, line_sample_min_num = 10, line3d_mle_iter_num = 100;
  double pt2line_mahdist_extractline = 1.5;
  int ransac_iters_extract_line = 100, num_cells_lineseg_range = 10;
  double ratio_support_pts_on_line = 0.7, stdev_sample_pt_imgline = 3.0;
  double depth_stdev_coeff_c1 = 0.00273, depth_stdev_coeff_c2 = 0.00074, depth_stdev_coeff_c3 = -0.00058;
  double depth_scaling = 1.0;
  // pair registration
  int ransac_iters_line_motion = 500, adjacent_linematch_window = 3, line_match_number_weight = 1;
  int min_feature_matches = 20, min_matches_loopclose = 20;
  double max_mah_dist_for_inliers = 3.0, g2o_line_error_weight = 1.0, g2o_BA_kernel_delta = 10.0;
  int g2o_BA_use_kernel = 1;
  double pt2line3d_dist_relmotion = 0.05, line3d_angle_relmotion = 10.0;
  // point features (src/parameter_server.cpp:45,146)
  double sigma_depth = 0.01, nn_distance_ratio = 0.5;
};

// glibc TYPE_3 rand() restated; stdlib/random_r.c semantics.
struct GlibcRand {
  int32_t r[34];
  int f, b;  // front / rear indices into the 31-word state
  void seed(uint32_t s);
  int next();  // == rand()
};

struct Segment { double x1, y1, x2, y2, width; };

struct LsdDebug {  // optional intermediates for stage-wise parity tests
  int sw = 0, sh = 0;
  std::vector<double> scaled, angles, modgrad;
  std::vector<int32_t> seeds;  // x | y<<16 in list order
};

// a1: cvtColor(CV_RGB2GRAY) OpenCV 2.4 fixed point (src/node.cpp:191-196)
void gray_from_3ch(const uint8_t* img, int W, int H, uint8_t* gray);
// a2-a9: callLsd (src/line/utils.cpp:112-135) + lsd() (external/lsd/lsd.cpp:2094)
void lsd_detect(const uint8_t* gray, int W, int H, const Params& P, std::vector<Segment>& out,
                LsdDebug* dbg = nullptr);

struct Pt3 { double pos[3], cov[9], DU[9], W_sqrt[3]; };  // RandomPoint3d, lineslam.h:41-82

struct Line {  // FrameLine + RandomLine3d (src/line/lineslam.h:84-151), POD
  double p[2], q[2], lineEq2d[3], r[2], des[72];
  double A[3], B[3], covA[9], covB[9], DU_A[9], DU_B[9], Wsqrt_A[3], Wsqrt_B[3];
  int32_t lid, haveDepth;
};

struct ExtractDebug {
  std::vector<Segment> segs;          // all LSD segments
  std::vector<int> seg_of_line;       // LSD segment index of each kept line
  std::vector<std::vector<int>> inlier_idx;  // per kept line: RANSAC inlier sample indices
  std::vector<std::vector<double>> pts;      // per kept line: inlier points xyz
  std::vector<double> A0B0;           // per kept line: RANSAC endpoints before MLE (6)
  std::vector<int> lm_iters;          // per kept line: LM iterations
  std::vector<double> gx, gy;         // Sobel planes
  int rand_draws = 0;
};

// Node::detect3DLines (src/line/lineslam.cpp:200-357); K row-major 3x3.
void detect3DLines(const uint8_t* gray, const float* depth, int W, int H, const double K[9],
                   double asynch_dt, uint32_t seed, const Params& P, std::vector<Line>& lines,
                   ExtractDebug* dbg = nullptr, int omp_threads = 1);

struct Match { int32_t queryIdx, trainIdx; float distance; };

// Node::lineMatching (src/node.cpp:1619-1694); f1 = this (query), f2 = other (train)
void lineMatching(const std::vector<Line>& f1, const std::vector<Line>& f2, bool adjacent,
                  std::vector<Match>& matches, int omp_threads = 1);

// Point features of a frame as the pair stage sees them (inputs):
// feature_locations_3d_ (Eigen::Vector4f, w = 1, z may be NaN) and feature_descriptors_ (CV_32F rows).
struct Points { int n = 0, dim = 0; const float* xyz1 = nullptr; const float* desc = nullptr; };

// Node::featureMatching, BRUTEFORCE branch (src/node.cpp:606-641): BFMatcher L2 knnMatch k = 2, ratio test,
// unique trainIdx, distance = ratio + rand()/(1000 RAND_MAX); one rand() per accepted match.
void featureMatching(const Points& query, const Points& train, double nn_ratio, GlibcRand& rng, std::vector<Match>& out);
void featureMatching_hamming(const uint8_t* qd, int nq, const uint8_t* td, int nt, int nbytes, double nn_ratio, GlibcRand& rng,
                             std::vector<Match>& out);   // ORB rows, BruteForce-HammingLUT (src/node.cpp:609-613)
// squareroot_descriptor_space (src/node.cpp:1823-1837), in place
void rootsift(float* desc, int n, int dim);

struct PoseResult {
  std::vector<Match> pt_inliers;          // output_point_inlier_matches
  std::vector<Match> pt_ransac_inliers;   // max_point_inlier_set of the best hypothesis
  bool found = false;
  float tf[16];       // row-major Matrix4f, query -> train
  float rmse = 1e9f;
  std::vector<Match> inliers;          // refined line inliers (output_line_inlier_matches)
  std::vector<Match> ransac_inliers;   // max_line_inlier_set of the best hypothesis (Tier-E)
  float tf_ransac[16];                 // tf_best before refinement
  int best_iter = -1;
};

// getTransform_PtsLines_ransac (src/line/motion.cpp:605-849), line-only inputs (nPt = 0)
void getTransform_Lines_ransac(const std::vector<Line>& train, const std::vector<Line>& query,
                               int id_train, int id_query, const std::vector<Match>& ln_matches,
                               uint32_t seed, const Params& P, PoseResult& out);

// getTransform_PtsLines_ransac (src/line/motion.cpp:605-849) with point and line matches; fx = K(0,0) (the
// focal length compPt3dCov uses for the point-edge information, transformation_estimation.cpp:243-262).
// The rand() stream continues from `rng` (featureMatching draws first in matchNodePair).
void getTransform_PtsLines_ransac(const std::vector<Line>& train, const std::vector<Line>& query,
                                  const Points& train_pts, const Points& query_pts, int id_train, int id_query,
                                  const std::vector<Match>& pt_matches, const std::vector<Match>& ln_matches,
                                  GlibcRand& rng, double fx, double asynch_dt, const Params& P, PoseResult& out);
// getTransformFromHybridMatchesG2O (src/transformation_estimation.cpp:218-461) restated natively
void refine_pose_hybrid(const std::vector<Line>& train, const std::vector<Line>& query, const Points& train_pts,
                        const Points& query_pts, const std::vector<Match>& pt_ms, const std::vector<Match>& ln_ms,
                        float tf[16], int iterations, double fx, double asynch_dt, const Params& P);

// computeRelativeMotion_Ransac (src/line/motion.cpp:367-526): line-only RANSAC on paired 3D lines (a[i] <-> b[i],
// x_b = R x_a + t) with Euclidean consensus, then optimizeRelmotion (motion.cpp:98-139: dlevmar_dif, m = 7
// quaternion + translation, cost motion.cpp:60-96) and consensus growing. Returns the consensus index set.
struct RelMotion { double R[9], t[3]; std::vector<int> conset; int lm_calls = 0; bool have = false; };
void computeRelativeMotion_Ransac(const std::vector<Line>& a, const std::vector<Line>& b, uint32_t seed, const Params& P,
                                  RelMotion& out);
void optimizeRelmotion(const std::vector<Line>& a, const std::vector<Line>& b, double R[9], double t[3]);

// sub-pieces exposed for unit tests
void get_gradient_probe(const double* xG, const double* yG, int W, int H, const double* pq, double* r);  // FrameLine::getGradient
double cvnorm(const double* v, int len);                  // cv::norm of a CV_64F vector (OpenCV 2.4 normL2_)
double cvnorm_diff72(const double* a, const double* b);   // cv::norm(des1 - des2), 72-D
void sobel5(const uint8_t* gray, int W, int H, std::vector<double>& gx, std::vector<double>& gy);
int dlevmar_dif_restated(void (*func)(double*, double*, int, int, void*), double* p, double* x, int m,
                         int n, int itmax, const double opts[5], double info[10], void* adata);
bool relmotion_svd(const double* qA, const double* qB, const double* tA, const double* tB, int n,
                   double R[9], double t[3]);
void refine_pose_lines(const std::vector<Line>& train, const std::vector<Line>& query,
                       const std::vector<Match>& ms, float tf[16], int iterations, const Params& P);

}  // namespace orc
