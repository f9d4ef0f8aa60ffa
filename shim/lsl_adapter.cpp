// lsl_adapter.cpp — the drop-in a LineSLAM maintainer compiles into the reference (-DUSE_LSL_B200, -llsl_b200): the
// reference's own member functions, same signatures and error behaviour, re-implemented on the C ABI of include/lsl.h.
//   Node::detect3DLines             src/node.h:286-287   (body: src/line/lineslam.cpp:200-357)
//   Node::lineMatching              src/node.h:288       (src/node.cpp:1619-1694)
//   Node::featureMatching           src/node.h:139       (BRUTEFORCE branch, src/node.cpp:606-641)
//   getTransform_PtsLines_ransac    src/line/utils.h:147-153 (src/line/motion.cpp:605-849)
// In this repository it is compiled against shim/stubs/shim_stubs.h (LSL_SHIM_STUBS) because OpenCV / Eigen / ROS headers
// are not in the image; tests/test_shim.py builds it and tests/test_gpu_shim.py runs it on the device.
#include <stdexcept>
#include "../include/lsl.h"
#ifdef LSL_SHIM_STUBS
#include "stubs/shim_stubs.h"
#else
#include "node.h"
#include "line/utils.h"
#endif

static lsl_ctx* g_lsl = nullptr;   // one per process / GPU
extern "C" void lsl_shim_shutdown() { if (g_lsl) { lsl_ctx_destroy(g_lsl); g_lsl = nullptr; } }
static lsl_ctx* lsl() {
  if (!g_lsl) {
    lsl_params p;
    lsl_params_default(&p);
    p.lsd_ang_th = sysPara.lsd_angle_thres;                  // SystemParameters::init, src/line/lineslam.cpp:577-640
    p.lsd_density_th = sysPara.lsd_density_thres;
    p.line_2d_len_thres = sysPara.line_2d_len_thres;
    p.line_3d_len_thres_m = sysPara.line_3d_len_thres_m;
    p.min_feature_matches = sysPara.min_feature_matches;
    p.min_matches_loopclose = sysPara.min_matches_loopclose;
    p.line_match_number_weight = sysPara.line_match_number_weight;
    p.ransac_iters_line_motion = sysPara.ransac_iters_line_motion;
    p.max_mah_dist_for_inliers = sysPara.max_mah_dist_for_inliers;
    p.g2o_line_error_weight = sysPara.g2o_line_error_weight;
    int rc = lsl_ctx_create(&g_lsl, &p, /*cuda_device*/ 0, /*max_batch*/ 8, 1280, 960);
    if (rc != LSL_OK) throw std::runtime_error(std::string("liblsl_b200: ") + lsl_strerror(rc));   // no CPU fallback
  }
  return g_lsl;
}

Node::~Node() { if (lsl_handle_) lsl_frame_free(lsl_handle_); }

static cv::Mat mat3x3(const double* v) { cv::Mat m(3, 3, CV_64F); for (int i = 0; i < 9; ++i) m.at<double>(i / 3, i % 3) = v[i]; return m; }
static void fill_rnd(RandomPoint3d& r, const double* pos, const double* cov, const double* DU, const double* Wsqrt) {
  r.pos = cv::Point3d(pos[0], pos[1], pos[2]);
  r.cov = mat3x3(cov);
  for (int i = 0; i < 9; ++i) r.DU[i] = DU[i];
  for (int i = 0; i < 3; ++i) {
    r.W_sqrt[i] = Wsqrt[i];
    r.dux[i] = DU[3 * i] * pos[0] + DU[3 * i + 1] * pos[1] + DU[3 * i + 2] * pos[2];   // lineslam.h:77-79
  }
}

// src/line/lineslam.cpp:200-357. Threshold arguments are parameters of the context (sysPara at creation), as in the reference
// where they are read from sysPara at the call site (src/node.cpp:214-215).
void Node::detect3DLines(const cv::Mat& gray_uchar, const cv::Mat& depth_float, double, const cv::Mat& K, double, double, double, string) {
  double Kr[9];
  for (int i = 0; i < 9; ++i) Kr[i] = K.at<double>(i / 3, i % 3);
  lsl_frame* f = nullptr;
  int rc = lsl_extract(lsl(), gray_uchar.data, 1, (const float*)depth_float.data, gray_uchar.cols, gray_uchar.rows, Kr,
                       asynch_time_diff_sec_, /*seed*/ 1u + (uint32_t)id_, &f);
  if (rc != LSL_OK) throw std::runtime_error(lsl_last_error(lsl()));   // the reference exit(0)s here (lineslam.cpp:272-275)
  if (lsl_handle_) lsl_frame_free(lsl_handle_);
  lsl_handle_ = f;
  std::vector<lsl_line_rec> recs((size_t)lsl_frame_num_lines(f));
  int n = 0;
  if ((rc = lsl_frame_lines(f, recs.data(), (int)recs.size(), &n)) != LSL_OK) throw std::runtime_error(lsl_last_error(lsl()));
  lines.assign((size_t)n, FrameLine());
  for (int i = 0; i < n; ++i) {
    FrameLine& L = lines[(size_t)i];
    const lsl_line_rec& r = recs[(size_t)i];
    L.p = cv::Point2d(r.p[0], r.p[1]); L.q = cv::Point2d(r.q[0], r.q[1]);
    L.l = cv::Mat(3, 1, CV_64F);
    for (int k = 0; k < 3; ++k) { L.l.at<double>(k) = r.lineEq2d[k]; L.lineEq2d[k] = r.lineEq2d[k]; }
    L.r = cv::Point2d(r.r[0], r.r[1]);
    L.des = cv::Mat(72, 1, CV_64F);
    for (int k = 0; k < 72; ++k) L.des.at<double>(k) = r.des[k];
    L.haveDepth = r.haveDepth != 0; L.lid = r.lid;
    L.line3d.A = cv::Point3d(r.A[0], r.A[1], r.A[2]); L.line3d.B = cv::Point3d(r.B[0], r.B[1], r.B[2]);
    L.line3d.covA = mat3x3(r.covA); L.line3d.covB = mat3x3(r.covB);
    fill_rnd(L.line3d.rndA, r.A, r.covA, r.DU_A, r.Wsqrt_A);
    fill_rnd(L.line3d.rndB, r.B, r.covB, r.DU_B, r.Wsqrt_B);
  }
}

// src/node.cpp:1619-1694
unsigned int Node::lineMatching(const Node* other, const bool adjacentFrame, std::vector<cv::DMatch>* matches) const {
  std::vector<lsl_match> m(lines.size() ? lines.size() : 1);
  int n = 0;
  if (!lsl_handle_ || !other->lsl_handle_) return 0;
  if (lsl_match_lines(lsl(), lsl_handle_, other->lsl_handle_, adjacentFrame ? 1 : 0, m.data(), (int)m.size(), &n) != LSL_OK) return 0;
  for (int i = 0; i < n; ++i) matches->push_back(cv::DMatch(m[(size_t)i].queryIdx, m[(size_t)i].trainIdx, m[(size_t)i].distance));
  return (unsigned int)n;
}

// end of Node::Node (src/node.cpp:219-310): the detectors' output becomes an input of the device pair stage
void Node::uploadPointFeatures() {
  if (!lsl_handle_) return;
  lsl_frame_set_points(lsl(), lsl_handle_, feature_locations_3d_.empty() ? nullptr : feature_locations_3d_[0].data(),
                       (const float*)feature_descriptors_.data, (int)feature_locations_3d_.size(), feature_descriptors_.cols ? feature_descriptors_.cols : 1);
}

// src/node.cpp:606-641 (matcher_type BRUTEFORCE)
unsigned int Node::featureMatching(const Node* other, std::vector<cv::DMatch>* matches) const {
  std::vector<lsl_match> m(feature_locations_3d_.size() ? feature_locations_3d_.size() : 1);
  int n = 0;
  if (!lsl_handle_ || !other->lsl_handle_) return 0;
  if (lsl_match_points(lsl(), lsl_handle_, other->lsl_handle_, /*seed*/ 1u + (uint32_t)id_, m.data(), (int)m.size(), &n) != LSL_OK) return 0;
  for (int i = 0; i < n; ++i) matches->push_back(cv::DMatch(m[(size_t)i].queryIdx, m[(size_t)i].trainIdx, m[(size_t)i].distance));
  return (unsigned int)n;
}

// src/line/motion.cpp:605-849
bool getTransform_PtsLines_ransac(const Node* trainNode, const Node* queryNode, const std::vector<cv::DMatch> all_point_matches,
                                  const std::vector<cv::DMatch> all_line_matches, std::vector<cv::DMatch>& output_point_inlier_matches,
                                  std::vector<cv::DMatch>& output_line_inlier_matches, Eigen::Matrix4f& ransac_tf, float& inlier_rmse) {
  const size_t nl = all_line_matches.size(), np = all_point_matches.size();
  std::vector<lsl_match> m(nl ? nl : 1), pm(np ? np : 1), inl(nl ? nl : 1), pinl(np ? np : 1);
  for (size_t i = 0; i < nl; ++i) { m[i].queryIdx = all_line_matches[i].queryIdx; m[i].trainIdx = all_line_matches[i].trainIdx; m[i].distance = all_line_matches[i].distance; }
  for (size_t i = 0; i < np; ++i) { pm[i].queryIdx = all_point_matches[i].queryIdx; pm[i].trainIdx = all_point_matches[i].trainIdx; pm[i].distance = all_point_matches[i].distance; }
  lsl_pose_rec rec;
  int ni = 0, npi = 0;
  inlier_rmse = 1e9f;
  if (lsl_pose_ransac(lsl(), trainNode->lsl_handle_, queryNode->lsl_handle_, trainNode->id_, queryNode->id_, pm.data(), (int)np, m.data(),
                      (int)nl, /*seed*/ 1u + (uint32_t)queryNode->id_, &rec, inl.data(), (int)inl.size(), &ni, nullptr, 0, nullptr) != LSL_OK)
    return false;
  lsl_pair_matches(lsl(), 0, 4, pinl.data(), (int)pinl.size(), &npi);   // output_point_inlier_matches
  for (int i = 0; i < ni; ++i) output_line_inlier_matches.push_back(cv::DMatch(inl[(size_t)i].queryIdx, inl[(size_t)i].trainIdx, inl[(size_t)i].distance));
  for (int i = 0; i < npi; ++i) output_point_inlier_matches.push_back(cv::DMatch(pinl[(size_t)i].queryIdx, pinl[(size_t)i].trainIdx, pinl[(size_t)i].distance));
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) ransac_tf(r, c) = rec.tf[r * 4 + c];   // query (newer) -> train (older), motion.cpp:534
  inlier_rmse = rec.rmse;
  return rec.found != 0;
}
