// shim_stubs.h — ~150-line stand-ins for the reference's third-party and own headers, ONLY so that shim/lsl_adapter.cpp
// (the drop-in with the reference's signatures) can be compiled and run in this image, where OpenCV / Eigen / ROS headers
// are absent. In the reference tree the adapter is compiled with -DUSE_LSL_B200 against the real node.h / lineslam.h /
// utils.h instead (LSL_SHIM_STUBS undefined). Member names and types mirror src/line/lineslam.h:41-151 and src/node.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6

namespace cv {
struct Point2d { double x = 0, y = 0; Point2d() {} Point2d(double x_, double y_) : x(x_), y(y_) {} };
struct Point3d { double x = 0, y = 0, z = 0; Point3d() {} Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {} };
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0;
  DMatch() {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), distance(d) {}
};
// dense, continuous, ref-counted matrix view: what the adapter needs of cv::Mat (rows, cols, data, at<T>, clone)
class Mat {
 public:
  int rows = 0, cols = 0, type_ = CV_8U;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), type_(type), data((uint8_t*)ext) {}   // borrowed
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    own_ = std::shared_ptr<std::vector<uint8_t>>(new std::vector<uint8_t>((size_t)r * c * elem(type)));
    data = own_->data();
  }
  static size_t elem(int type) { return type == CV_64F ? 8 : type == CV_32F ? 4 : 1; }
  Mat clone() const { Mat m(rows, cols, type_); if (data) memcpy(m.data, data, (size_t)rows * cols * elem(type_)); return m; }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  template <class T> T& at(int r, int c = 0) { return ((T*)data)[(size_t)r * cols + c]; }
  template <class T> const T& at(int r, int c = 0) const { return ((const T*)data)[(size_t)r * cols + c]; }
 private:
  std::shared_ptr<std::vector<uint8_t>> own_;
};
}  // namespace cv

namespace Eigen {
struct Vector4f { float v[4] = {0, 0, 0, 1}; float* data() { return v; } const float* data() const { return v; } float& operator()(int i) { return v[i]; } };
struct Matrix4f {            // column-major storage like Eigen's default
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
};
}  // namespace Eigen

using std::string;
using std::vector;

// src/line/lineslam.h:41-151 (public members only; the adapter fills W_sqrt / DU from the device record instead of
// re-running cv::SVD in the constructor)
class RandomPoint3d {
 public:
  cv::Point3d pos; cv::Mat cov, U, W; double W_sqrt[3], DU[9], dux[3];
  RandomPoint3d() {}
};
class RandomLine3d {
 public:
  vector<RandomPoint3d> pts; cv::Point3d A, B; cv::Mat covA, covB; RandomPoint3d rndA, rndB; cv::Point3d u, d;
};
class FrameLine {
 public:
  cv::Point2d p, q; cv::Mat l; double lineEq2d[3]; bool haveDepth = false; RandomLine3d line3d; cv::Point2d r; cv::Mat des;
  int lid = -1, gid = -1, lid_prvKfrm = -1;
};
struct SystemParametersStub {   // the fields of SystemParameters (src/line/lineslam.h:215-275) the adapter forwards
  double lsd_angle_thres = 22.5, lsd_density_thres = 0.7, line_2d_len_thres = 10.0, line_3d_len_thres_m = 0.02;
  int min_feature_matches = 20, min_matches_loopclose = 20, line_match_number_weight = 1, ransac_iters_line_motion = 500;
  double max_mah_dist_for_inliers = 3.0, g2o_line_error_weight = 1.0;
};
extern SystemParametersStub sysPara;

struct lsl_frame;
// src/node.h: the members / methods the line front end touches
class Node {
 public:
  int id_ = 0;
  double asynch_time_diff_sec_ = 0.0;
  std::vector<FrameLine> lines;
  std::vector<Eigen::Vector4f> feature_locations_3d_;
  cv::Mat feature_descriptors_;
  lsl_frame* lsl_handle_ = nullptr;   // new member under USE_LSL_B200: the device-resident frame
  Node() {}
  ~Node();
  void detect3DLines(const cv::Mat& gray_uchar, const cv::Mat& depth_float, double line2d_len_thres, const cv::Mat& K,
                     double ratio_of_collinear_pts, double line_3d_len_thres_m, double depth_scaling, string algorithm);   // node.h:286-287
  unsigned int lineMatching(const Node* other, const bool adjacentFrame, std::vector<cv::DMatch>* matches) const;          // node.h:288
  unsigned int featureMatching(const Node* other, std::vector<cv::DMatch>* matches) const;                                 // node.h:139
  void uploadPointFeatures();
};
// src/line/utils.h:147-153
bool getTransform_PtsLines_ransac(const Node* trainNode, const Node* queryNode, const std::vector<cv::DMatch> all_point_matches,
                                  const std::vector<cv::DMatch> all_line_matches, std::vector<cv::DMatch>& output_point_inlier_matches,
                                  std::vector<cv::DMatch>& output_line_inlier_matches, Eigen::Matrix4f& ransac_tf, float& inlier_rmse);
