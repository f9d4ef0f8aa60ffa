// oracle_pair.cpp — CPU restatement of the pair-registration path (TEST INFRASTRUCTURE):
// Node::lineMatching (src/node.cpp:1619-1694), getTransform_PtsLines_ransac with line-only
// input (src/line/motion.cpp:605-849), getTransform_Line_svd / computeRelativeMotion_svd
// (motion.cpp:581-603, 315-365) and a native restatement of the g2o refinement
// getTransformFromHybridMatchesG2O (src/transformation_estimation.cpp:218-461, line edges
// src/line/edge_se3_lineendpts.cpp:146-189). g2o itself is not in the container: the LM
// follows SURVEY.md Appendix C.3 from memory -> parity UNPINNED for that stage (Tier-T).
#include "oracle.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <algorithm>
#include "../lineslam_b200/csrc/shared/lsl_math.h"
#include "../lineslam_b200/csrc/shared/lsl_linalg.h"
#include "../lineslam_b200/csrc/shared/lsl_points.h"

using namespace lslm;

namespace orc {

double cvnorm_diff72(const double* a, const double* b) {  // cv::norm(a - b), see oracle_extract.cpp:cvnorm
  double result = 0;
  for (int i = 0; i < 72; i += 4) {
    double v0 = a[i] - b[i], v1 = a[i + 1] - b[i + 1];
    result += v0 * v0 + v1 * v1;
    v0 = a[i + 2] - b[i + 2]; v1 = a[i + 3] - b[i + 3];
    result += v0 * v0 + v1 * v1;
  }
  return sqrt(result);
}
// pt_to_line_dist2d (utils.cpp:1250-1264)
static double pt_to_line_dist2d(const double p[2], const double l[3]) {
  double a = l[0], b = l[1], c = l[2], x = p[0], y = p[1];
  return fabs((a * x + b * y + c)) / sqrt(a * a + b * b);
}
// line_to_line_dist2d (utils.cpp:1265-1273)
static double line_to_line_dist2d(const Line& a, const Line& b) {
  return 0.25 * pt_to_line_dist2d(a.p, b.lineEq2d) + 0.25 * pt_to_line_dist2d(a.q, b.lineEq2d) +
         0.25 * pt_to_line_dist2d(b.p, a.lineEq2d) + 0.25 * pt_to_line_dist2d(b.q, a.lineEq2d);
}
static double norm2(double x, double y) { return sqrt(x * x + y * y); }
// projectPt2d_to_line2d (utils.cpp:1612-1618)
static double project2d(const double X[2], const double A[2], const double B[2]) {
  double BX[2] = {X[0] - B[0], X[1] - B[1]}, BA[2] = {A[0] - B[0], A[1] - B[1]};
  double n = norm2(BA[0], BA[1]);
  return (BX[0] * BA[0] + BX[1] * BA[1]) / n / n;
}
// lineSegmentOverlap (utils.cpp:1620-1638)
static double lineSegmentOverlap(const Line& a, const Line& b) {
  double la = norm2(a.p[0] - a.q[0], a.p[1] - a.q[1]), lb = norm2(b.p[0] - b.q[0], b.p[1] - b.q[1]);
  if (la < lb) {
    double lp = project2d(a.p, b.p, b.q), lq = project2d(a.q, b.p, b.q);
    if ((lp < 0 && lq < 0) || (lp > 1 && lq > 1)) return -1;
    return fabs(lp - lq) * lb;
  } else {
    double lp = project2d(b.p, a.p, a.q), lq = project2d(b.q, a.p, a.q);
    if ((lp < 0 && lq < 0) || (lp > 1 && lq > 1)) return -1;
    return fabs(lp - lq) * la;
  }
}

void lineMatching(const std::vector<Line>& f1, const std::vector<Line>& f2, bool adjacent,
                  std::vector<Match>& matches, int omp_threads) {
  const double PI_T = 3.14159265;  // lineslam.h:38
  double lineDistThresh, lineAngleThresh, descDiffThresh, lineOverlapThresh, ratio;
  if (adjacent) { lineDistThresh = 45; lineAngleThresh = 30 * PI_T / 180; descDiffThresh = 0.85; lineOverlapThresh = 0; ratio = 0.7; }
  else { lineDistThresh = 80; lineAngleThresh = 30 * PI_T / 180; descDiffThresh = 0.7; lineOverlapThresh = -1; ratio = 0.7; }
  int n1 = (int)f1.size(), n2 = (int)f2.size();
  if (n1 == 0 || n2 == 0) return;
  double cosT = lsl_cos(lineAngleThresh);
  std::vector<double> D((size_t)n1 * n2, 100.0);
#pragma omp parallel for num_threads(omp_threads) if (omp_threads > 1)
  for (int i = 0; i < n1; ++i)
    for (int j = 0; j < n2; ++j)
      if ((f1[i].r[0] * f2[j].r[0] + f1[i].r[1] * f2[j].r[1] > cosT) &&
          (line_to_line_dist2d(f1[i], f2[j]) < lineDistThresh) &&
          (lineSegmentOverlap(f1[i], f2[j]) > lineOverlapThresh))
        D[(size_t)i * n2 + j] = cvnorm_diff72(f1[i].des, f2[j].des);
  for (int i = 0; i < n1; ++i) {
    double minVal = D[(size_t)i * n2]; int minX = 0;       // cv::minMaxLoc: first minimum
    for (int j = 1; j < n2; ++j) if (D[(size_t)i * n2 + j] < minVal) { minVal = D[(size_t)i * n2 + j]; minX = j; }
    if (minVal < descDiffThresh) {
      double minV = D[minX]; int minY = 0;
      for (int j = 1; j < n1; ++j) if (D[(size_t)j * n2 + minX] < minV) { minV = D[(size_t)j * n2 + minX]; minY = j; }
      if (i == minY) {
        double rowmin2 = 100, colmin2 = 100;
        for (int j = 0; j < n2; ++j) { if (j == minX) continue; if (rowmin2 > D[(size_t)i * n2 + j]) rowmin2 = D[(size_t)i * n2 + j]; }
        for (int j = 0; j < n1; ++j) { if (j == minY) continue; if (colmin2 > D[(size_t)j * n2 + minX]) colmin2 = D[(size_t)j * n2 + minX]; }
        if (rowmin2 * ratio > minVal && colmin2 * ratio > minVal) {
          Match m; m.queryIdx = i; m.trainIdx = minX; m.distance = (float)minVal;
          if (f1[i].haveDepth && f2[minX].haveDepth) matches.push_back(m);
        }
      }
    }
  }
}

// ------------------------------------------------------- minimal solver ----
static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static void q2r(const double q[4], double R[9]) {  // utils.cpp:1659-1694
  double a = q[0], b = q[1], c = q[2], d = q[3];
  double nm = sqrt(a * a + b * b + c * c + d * d);
  a = a / nm; b = b / nm; c = c / nm; d = d / nm;
  R[0] = a * a + b * b - c * c - d * d; R[1] = 2 * b * c - 2 * a * d; R[2] = 2 * b * d + 2 * a * c;
  R[3] = 2 * b * c + 2 * a * d; R[4] = a * a - b * b + c * c - d * d; R[5] = 2 * c * d - 2 * a * b;
  R[6] = 2 * b * d - 2 * a * c; R[7] = 2 * c * d + 2 * a * b; R[8] = a * a - b * b - c * c + d * d;
}
static void skew(const double v[3], double m[9]) {  // vec2SkewMat utils.cpp:1649
  m[0] = 0; m[1] = -v[2]; m[2] = v[1]; m[3] = v[2]; m[4] = 0; m[5] = -v[0]; m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}
// computeRelativeMotion_svd (motion.cpp:315-365). a = query lines, b = train lines; x_b = R x_a + t.
bool relmotion_svd(const double* aA, const double* aB, const double* bA, const double* bB, int n, double R[9], double t[3]) {
  if (n < 2) return false;
  std::vector<double> au(3 * n), ad(3 * n), bu(3 * n), bd(3 * n);
  for (int i = 0; i < n; ++i) {
    for (int s = 0; s < 2; ++s) {
      const double* A = s ? bA + 3 * i : aA + 3 * i;
      const double* B = s ? bB + 3 * i : aB + 3 * i;
      double* u = s ? &bu[3 * i] : &au[3 * i];
      double* d = s ? &bd[3 * i] : &ad[3 * i];
      double l[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
      double m[3] = {(A[0] + B[0]) * 0.5, (A[1] + B[1]) * 0.5, (A[2] + B[2]) * 0.5};
      double inv = 1 / sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      u[0] = l[0] * inv; u[1] = l[1] * inv; u[2] = l[2] * inv;
      cross3(u, m, d);
    }
  }
  double A[16];
  for (int i = 0; i < 16; ++i) A[i] = 0;
  for (int i = 0; i < n; ++i) {
    double Ai[16];
    for (int k = 0; k < 16; ++k) Ai[k] = 0;
    double dm[3] = {au[3 * i] - bu[3 * i], au[3 * i + 1] - bu[3 * i + 1], au[3 * i + 2] - bu[3 * i + 2]};
    double dp[3] = {au[3 * i] + bu[3 * i], au[3 * i + 1] + bu[3 * i + 1], au[3 * i + 2] + bu[3 * i + 2]};
    double dn[3] = {bu[3 * i] - au[3 * i], bu[3 * i + 1] - au[3 * i + 1], bu[3 * i + 2] - au[3 * i + 2]};
    Ai[1] = dm[0]; Ai[2] = dm[1]; Ai[3] = dm[2];
    Ai[4] = dn[0]; Ai[8] = dn[1]; Ai[12] = dn[2];
    double S[9]; skew(dp, S);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ai[(r + 1) * 4 + c + 1] = S[r * 3 + c];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += Ai[k * 4 + r] * Ai[k * 4 + c];
        A[r * 4 + c] = A[r * 4 + c] + s;
      }
  }
  double w[4], V[16];
  jacobi_sym<4>(A, w, V);
  double q[4] = {V[3], V[7], V[11], V[15]};  // svd.u.col(3): smallest singular value
  q2r(q, R);
  double uu[9], udr[3] = {0, 0, 0};
  for (int i = 0; i < 9; ++i) uu[i] = 0;
  for (int i = 0; i < n; ++i) {
    double S[9]; skew(&bu[3 * i], S);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += S[r * 3 + k] * S[c * 3 + k];  // S * S^T
        uu[r * 3 + c] = uu[r * 3 + c] + s;
      }
    double Rad[3], v[3];
    for (int r = 0; r < 3; ++r) Rad[r] = R[r * 3] * ad[3 * i] + R[r * 3 + 1] * ad[3 * i + 1] + R[r * 3 + 2] * ad[3 * i + 2];
    for (int r = 0; r < 3; ++r) v[r] = bd[3 * i + r] - Rad[r];
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += S[k * 3 + r] * v[k];  // S^T * v
      udr[r] = udr[r] + s;
    }
  }
  double ui[9];
  inv3(uu, ui);
  for (int r = 0; r < 3; ++r) t[r] = ui[r * 3] * udr[0] + ui[r * 3 + 1] * udr[1] + ui[r * 3 + 2] * udr[2];
  return true;
}

// Eigen Matrix4f * Vector4f, column-wise accumulation in float, rows 0..2
static void tf_apply_f(const float tf[16], const float v[4], double out[3]) {
  for (int r = 0; r < 3; ++r) {
    float acc = tf[r * 4 + 0] * v[0];
    acc = tf[r * 4 + 1] * v[1] + acc;
    acc = tf[r * 4 + 2] * v[2] + acc;
    acc = tf[r * 4 + 3] * v[3] + acc;
    out[r] = (double)acc;
  }
}

struct Score { std::vector<int> inl; double sse; };
// the line half of the scoring loops at motion.cpp:688-699 / :800-813
static void score_lines(const std::vector<Line>& train, const std::vector<Line>& query, const std::vector<Match>& ms,
                        const float tf[16], double thr, Score& sc, bool float_sse) {
  sc.inl.clear();
  float sse_f = 0; double sse_d = 0;
  for (size_t i = 0; i < ms.size(); ++i) {
    const Line& q = query[ms[i].queryIdx];
    const Line& t = train[ms[i].trainIdx];
    float qa[4] = {(float)q.A[0], (float)q.A[1], (float)q.A[2], 1.f}, qb[4] = {(float)q.B[0], (float)q.B[1], (float)q.B[2], 1.f};
    double qA[3], qB[3];
    tf_apply_f(tf, qa, qA);
    tf_apply_f(tf, qb, qB);
    double da = mah_dist3d_pt_line(t.A, t.DU_A, qA, qB);
    double db = mah_dist3d_pt_line(t.B, t.DU_B, qA, qB);
    if (da < thr && db < thr) {
      sc.inl.push_back((int)i);
      if (float_sse) sse_f += da * da + db * db; else sse_d += da * da + db * db;
    }
  }
  sc.sse = float_sse ? (double)sse_f : sse_d;
}

// ------------------------------------------------- g2o-style refinement ----
// Unknowns: cam1 pose (VertexSE3, right-multiplied increment [t, qxyz]) and one free
// 6-vector of endpoints per line match (VertexLineEndpts, additive), initialised with the
// newer frame's endpoints. Each match carries two EdgeSE3LineEndpts (newer: identity pose,
// fixed; older: cam1). Numeric central-difference Jacobians (delta 1e-9), Huber(delta),
// Levenberg-Marquardt with g2o's lambda policy; the arrow-shaped normal equations are
// solved exactly by eliminating the per-line 6x6 blocks (same solution as the full
// Cholesky up to rounding).
struct Iso { double R[9], t[3]; };
static void iso_mul(const Iso& a, const Iso& b, Iso& c) {
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[r * 3] * b.R[k] + a.R[r * 3 + 1] * b.R[3 + k] + a.R[r * 3 + 2] * b.R[6 + k];
    c.t[r] = a.R[r * 3] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
}
static void iso_inv(const Iso& a, Iso& c) {
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[k * 3 + r];
  for (int r = 0; r < 3; ++r) c.t[r] = -(c.R[r * 3] * a.t[0] + c.R[r * 3 + 1] * a.t[1] + c.R[r * 3 + 2] * a.t[2]);
}
static void iso_oplus(const Iso& est, const double u[6], Iso& out) {  // VertexSE3::oplusImpl, fromVectorMQT
  double w2 = 1. - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  double q[4] = {w2 > 0 ? sqrt(w2) : 0.0, u[3], u[4], u[5]};
  Iso inc;
  q2r(q, inc.R);
  inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
  iso_mul(est, inc, out);
}
// EdgeSE3LineEndpts::computeError (edge_se3_lineendpts.cpp:146-189); w2n = pose^-1
static void edge_error(const Iso& w2n, const double L[6], const double meas[6], const double AffA[9], const double AffB[9], double e[6]) {
  double ptA[3], ptB[3];
  for (int r = 0; r < 3; ++r) {
    ptA[r] = w2n.R[r * 3] * L[0] + w2n.R[r * 3 + 1] * L[1] + w2n.R[r * 3 + 2] * L[2] + w2n.t[r];
    ptB[r] = w2n.R[r * 3] * L[3] + w2n.R[r * 3 + 1] * L[4] + w2n.R[r * 3 + 2] * L[5] + w2n.t[r];
  }
  for (int h = 0; h < 2; ++h) {
    const double* Af = h ? AffB : AffA;
    const double* mp = meas + 3 * h;
    double dA[3] = {ptA[0] - mp[0], ptA[1] - mp[1], ptA[2] - mp[2]}, dB[3] = {ptB[0] - mp[0], ptB[1] - mp[1], ptB[2] - mp[2]};
    double Ap[3], Bp[3], BA[3];
    for (int r = 0; r < 3; ++r) {
      Ap[r] = Af[r * 3] * dA[0] + Af[r * 3 + 1] * dA[1] + Af[r * 3 + 2] * dA[2];
      Bp[r] = Af[r * 3] * dB[0] + Af[r * 3 + 1] * dB[1] + Af[r * 3 + 2] * dB[2];
    }
    for (int r = 0; r < 3; ++r) BA[r] = Bp[r] - Ap[r];
    double tt = -(Ap[0] * BA[0] + Ap[1] * BA[1] + Ap[2] * BA[2]) / (BA[0] * BA[0] + BA[1] * BA[1] + BA[2] * BA[2]);
    for (int r = 0; r < 3; ++r) e[3 * h + r] = Ap[r] + tt * BA[r];
  }
}
// endpt_AffnMat = D^-1/2 U^T (transformation_estimation.cpp:349-372)
static void affn(const double cov[9], double Af[9]) {
  double A[9], w[3], V[9];
  for (int i = 0; i < 9; ++i) A[i] = cov[i];
  jacobi_sym<3>(A, w, V);
  for (int i = 0; i < 3; ++i) {
    double d = sqrt(1 / w[i]);
    for (int j = 0; j < 3; ++j) Af[i * 3 + j] = d * V[j * 3 + i];
  }
}
static void huber(double e2, double delta, double rho[3]) {  // g2o RobustKernelHuber::robustify
  double dsqr = delta * delta;
  if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
  else { double sqrte = sqrt(e2); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
}

// Point edge EdgeSE3PointXYZ::computeError (src/line/edge_se3_ptxyz.cpp:84-90): e = w2n * X - measurement
static void pt_edge_error(const Iso& w2n, const double X[3], const double meas[3], double e[3]) {
  for (int r = 0; r < 3; ++r)
    e[r] = (w2n.R[r * 3] * X[0] + w2n.R[r * 3 + 1] * X[1] + w2n.R[r * 3 + 2] * X[2] + w2n.t[r]) - meas[r];
}
// chi2 = e^T Omega e (Omega e first, Eigen 3x3 * vec3 order)
static double pt_chi2(const double e[3], const double Om[9], double Oe[3]) {
  for (int r = 0; r < 3; ++r) Oe[r] = (Om[r * 3] * e[0] + Om[r * 3 + 1] * e[1]) + Om[r * 3 + 2] * e[2];
  return (e[0] * Oe[0] + e[1] * Oe[1]) + e[2] * Oe[2];
}

// Unknowns: cam1, one free 3-vector per point match (VertexPointXYZ, initialised with the newer frame's
// point), one free 6-vector per line match. Edges in insertion order: per point match newer then older
// (EdgeSE3PointXYZ, information = compPt3dCov(Vector3f)^-1), then per line match newer then older.
void refine_pose_hybrid(const std::vector<Line>& train, const std::vector<Line>& query, const Points& train_pts,
                        const Points& query_pts, const std::vector<Match>& pt_ms, const std::vector<Match>& ms,
                        float tf[16], int iterations, double fx, double asynch_dt, const Params& P) {
  int n = (int)ms.size(), np = (int)pt_ms.size();
  if (n + np == 0) return;
  // tfinv = tf.inverse() (Matrix4f general inverse) -> cam1 = SE3Quat(Quaterniond(R), t). The rigid
  // inverse in double and a quaternion round trip are restated as: cam1 = [R^T, -R^T t] in double.
  Iso tfd, cam1;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tfd.R[r * 3 + c] = (double)tf[r * 4 + c]; tfd.t[r] = (double)tf[r * 4 + 3]; }
  iso_inv(tfd, cam1);
  Iso ident; for (int i = 0; i < 9; ++i) ident.R[i] = (i % 4 == 0) ? 1 : 0; ident.t[0] = ident.t[1] = ident.t[2] = 0;
  std::vector<double> L(6 * n), measN(6 * n), measO(6 * n), AfN(18 * n), AfO(18 * n);
  for (int i = 0; i < n; ++i) {
    const Line& q = query[ms[i].queryIdx];
    const Line& t = train[ms[i].trainIdx];
    for (int k = 0; k < 3; ++k) { measN[6 * i + k] = q.A[k]; measN[6 * i + 3 + k] = q.B[k]; measO[6 * i + k] = t.A[k]; measO[6 * i + 3 + k] = t.B[k]; }
    for (int k = 0; k < 6; ++k) L[6 * i + k] = measN[6 * i + k];
    affn(q.covA, &AfN[18 * i]); affn(q.covB, &AfN[18 * i + 9]);
    affn(t.covA, &AfO[18 * i]); affn(t.covB, &AfO[18 * i + 9]);
  }
  std::vector<double> X(3 * np), pmN(3 * np), pmO(3 * np), OmN(9 * np), OmO(9 * np);
  for (int i = 0; i < np; ++i) {
    const float* qn = query_pts.xyz1 + 4 * pt_ms[i].queryIdx;
    const float* to = train_pts.xyz1 + 4 * pt_ms[i].trainIdx;
    for (int k = 0; k < 3; ++k) { pmN[3 * i + k] = (double)qn[k]; pmO[3 * i + k] = (double)to[k]; X[3 * i + k] = (double)qn[k]; }
    pt_info_f(qn, fx, P.stdev_sample_pt_imgline, P.depth_stdev_coeff_c1, P.depth_stdev_coeff_c2, P.depth_stdev_coeff_c3, asynch_dt, &OmN[9 * i]);
    pt_info_f(to, fx, P.stdev_sample_pt_imgline, P.depth_stdev_coeff_c1, P.depth_stdev_coeff_c2, P.depth_stdev_coeff_c3, asynch_dt, &OmO[9 * i]);
  }
  const double w = P.g2o_line_error_weight, hdelta = P.g2o_BA_kernel_delta;
  const bool robust = P.g2o_BA_use_kernel != 0;
  auto chi2_all = [&](const Iso& c1, const std::vector<double>& Xv, const std::vector<double>& Lv) {
    Iso w2n; iso_inv(c1, w2n);
    double chi = 0;
    for (int i = 0; i < np; ++i)
      for (int side = 0; side < 2; ++side) {
        double e[3], Oe[3];
        pt_edge_error(side ? w2n : ident, &Xv[3 * i], side ? &pmO[3 * i] : &pmN[3 * i], e);
        double c2 = pt_chi2(e, side ? &OmO[9 * i] : &OmN[9 * i], Oe);
        if (robust) { double rho[3]; huber(c2, hdelta, rho); chi += rho[0]; } else chi += c2;
      }
    for (int i = 0; i < n; ++i) {
      double e[6];
      for (int side = 0; side < 2; ++side) {
        edge_error(side ? w2n : ident, &Lv[6 * i], side ? &measO[6 * i] : &measN[6 * i], side ? &AfO[18 * i] : &AfN[18 * i],
                   side ? &AfO[18 * i + 9] : &AfN[18 * i + 9], e);
        double c2 = 0; for (int k = 0; k < 6; ++k) c2 += e[k] * w * e[k];
        if (robust) { double rho[3]; huber(c2, hdelta, rho); chi += rho[0]; } else chi += c2;
      }
    }
    return chi;
  };
  double lambda = 0, ni = 2;
  const double tau = 1e-5, lowS = 1. / 3., upS = 2. / 3.;
  const double del = 1e-9, scalar = 1 / (2 * del);
  std::vector<double> Hll(36 * n), Hpl(36 * n), bl(6 * n), dl(6 * n), HllInv(36 * n);
  std::vector<double> Hxx(9 * np), Hpx(18 * np), bx(3 * np), dx(3 * np), HxxInv(9 * np);
  for (int it = 0; it < iterations; ++it) {
    double currentChi = chi2_all(cam1, X, L);
    // ---- build system
    double Hpp[36], bp[6];
    for (int i = 0; i < 36; ++i) Hpp[i] = 0;
    for (int i = 0; i < 6; ++i) bp[i] = 0;
    Iso w2n; iso_inv(cam1, w2n);
    for (int i = 0; i < np; ++i) {
      double* hxx = &Hxx[9 * i]; double* hpx = &Hpx[18 * i]; double* b = &bx[3 * i];
      for (int k = 0; k < 9; ++k) hxx[k] = 0;
      for (int k = 0; k < 18; ++k) hpx[k] = 0;
      for (int k = 0; k < 3; ++k) b[k] = 0;
      for (int side = 0; side < 2; ++side) {
        const double* meas = side ? &pmO[3 * i] : &pmN[3 * i];
        const double* Om = side ? &OmO[9 * i] : &OmN[9 * i];
        double e[3], Jl[9], Jp[18];
        for (int d = 0; d < 3; ++d) {  // numeric Jacobian wrt the point vertex (additive)
          double Xp[3], e1[3], e2[3];
          for (int k = 0; k < 3; ++k) Xp[k] = X[3 * i + k];
          Xp[d] = X[3 * i + d] + del;
          pt_edge_error(side ? w2n : ident, Xp, meas, e1);
          Xp[d] = X[3 * i + d] + (-del);
          pt_edge_error(side ? w2n : ident, Xp, meas, e2);
          for (int k = 0; k < 3; ++k) Jl[k * 3 + d] = scalar * (e1[k] - e2[k]);
        }
        if (side) {
          for (int d = 0; d < 6; ++d) {
            double u[6] = {0, 0, 0, 0, 0, 0}, e1[3], e2[3];
            Iso c, ci;
            u[d] = del; iso_oplus(cam1, u, c); iso_inv(c, ci); pt_edge_error(ci, &X[3 * i], meas, e1);
            u[d] = -del; iso_oplus(cam1, u, c); iso_inv(c, ci); pt_edge_error(ci, &X[3 * i], meas, e2);
            for (int k = 0; k < 3; ++k) Jp[k * 6 + d] = scalar * (e1[k] - e2[k]);
          }
        }
        pt_edge_error(side ? w2n : ident, &X[3 * i], meas, e);
        double Oe[3];
        double c2 = pt_chi2(e, Om, Oe);
        double r1 = 1.0;
        if (robust) { double rho[3]; huber(c2, hdelta, rho); r1 = rho[1]; }
        double WO[9], WOe[3];
        for (int k = 0; k < 9; ++k) WO[k] = r1 * Om[k];
        for (int k = 0; k < 3; ++k) WOe[k] = r1 * Oe[k];
        double AtO[9];  // Jl^T * WO
        for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) { double s = 0; for (int j = 0; j < 3; ++j) s += Jl[j * 3 + a] * WO[j * 3 + k]; AtO[a * 3 + k] = s; }
        for (int a = 0; a < 3; ++a) {
          double s = 0; for (int k = 0; k < 3; ++k) s += Jl[k * 3 + a] * WOe[k];
          b[a] -= s;
          for (int c = 0; c < 3; ++c) { double h = 0; for (int k = 0; k < 3; ++k) h += AtO[a * 3 + k] * Jl[k * 3 + c]; hxx[a * 3 + c] += h; }
        }
        if (side) {
          double PtO[18];  // Jp^T * WO (6x3)
          for (int a = 0; a < 6; ++a) for (int k = 0; k < 3; ++k) { double s = 0; for (int j = 0; j < 3; ++j) s += Jp[j * 6 + a] * WO[j * 3 + k]; PtO[a * 3 + k] = s; }
          for (int a = 0; a < 6; ++a) {
            double s = 0; for (int k = 0; k < 3; ++k) s += Jp[k * 6 + a] * WOe[k];
            bp[a] -= s;
            for (int c = 0; c < 6; ++c) { double h = 0; for (int k = 0; k < 3; ++k) h += PtO[a * 3 + k] * Jp[k * 6 + c]; Hpp[a * 6 + c] += h; }
            for (int c = 0; c < 3; ++c) { double g = 0; for (int k = 0; k < 3; ++k) g += PtO[a * 3 + k] * Jl[k * 3 + c]; hpx[a * 3 + c] += g; }
          }
        }
      }
    }
    for (int i = 0; i < n; ++i) {
      double* hll = &Hll[36 * i]; double* hpl = &Hpl[36 * i]; double* b = &bl[6 * i];
      for (int k = 0; k < 36; ++k) { hll[k] = 0; hpl[k] = 0; }
      for (int k = 0; k < 6; ++k) b[k] = 0;
      for (int side = 0; side < 2; ++side) {
        const double* meas = side ? &measO[6 * i] : &measN[6 * i];
        const double* A1 = side ? &AfO[18 * i] : &AfN[18 * i];
        const double* A2 = A1 + 9;
        double e[6], Jl[36], Jp[36];
        // numeric Jacobian wrt the line vertex (additive)
        for (int d = 0; d < 6; ++d) {
          double Lp[6], e1[6], e2[6];
          for (int k = 0; k < 6; ++k) Lp[k] = L[6 * i + k];
          Lp[d] = L[6 * i + d] + del;
          edge_error(side ? w2n : ident, Lp, meas, A1, A2, e1);
          Lp[d] = L[6 * i + d] + (-del);
          edge_error(side ? w2n : ident, Lp, meas, A1, A2, e2);
          for (int k = 0; k < 6; ++k) Jl[k * 6 + d] = scalar * (e1[k] - e2[k]);
        }
        if (side) {  // numeric Jacobian wrt cam1
          for (int d = 0; d < 6; ++d) {
            double u[6] = {0, 0, 0, 0, 0, 0}, e1[6], e2[6];
            Iso c, ci;
            u[d] = del; iso_oplus(cam1, u, c); iso_inv(c, ci); edge_error(ci, &L[6 * i], meas, A1, A2, e1);
            u[d] = -del; iso_oplus(cam1, u, c); iso_inv(c, ci); edge_error(ci, &L[6 * i], meas, A1, A2, e2);
            for (int k = 0; k < 6; ++k) Jp[k * 6 + d] = scalar * (e1[k] - e2[k]);
          }
        }
        edge_error(side ? w2n : ident, &L[6 * i], meas, A1, A2, e);
        double c2 = 0; for (int k = 0; k < 6; ++k) c2 += e[k] * w * e[k];
        double wgt = w;
        if (robust) { double rho[3]; huber(c2, hdelta, rho); wgt = rho[1] * w; }
        for (int a = 0; a < 6; ++a) {
          double s = 0; for (int k = 0; k < 6; ++k) s += Jl[k * 6 + a] * (wgt * e[k]);
          b[a] -= s;
          for (int c = 0; c < 6; ++c) { double h = 0; for (int k = 0; k < 6; ++k) h += Jl[k * 6 + a] * wgt * Jl[k * 6 + c]; hll[a * 6 + c] += h; }
        }
        if (side) {
          for (int a = 0; a < 6; ++a) {
            double s = 0; for (int k = 0; k < 6; ++k) s += Jp[k * 6 + a] * (wgt * e[k]);
            bp[a] -= s;
            for (int c = 0; c < 6; ++c) {
              double h = 0, g = 0;
              for (int k = 0; k < 6; ++k) { h += Jp[k * 6 + a] * wgt * Jp[k * 6 + c]; g += Jp[k * 6 + a] * wgt * Jl[k * 6 + c]; }
              Hpp[a * 6 + c] += h;
              hpl[a * 6 + c] += g;
            }
          }
        }
      }
    }
    if (it == 0) {  // computeLambdaInit: tau * max diagonal entry
      double maxDiag = 0;
      for (int a = 0; a < 6; ++a) maxDiag = std::max(fabs(Hpp[a * 6 + a]), maxDiag);
      for (int i = 0; i < np; ++i) for (int a = 0; a < 3; ++a) maxDiag = std::max(fabs(Hxx[9 * i + a * 3 + a]), maxDiag);
      for (int i = 0; i < n; ++i) for (int a = 0; a < 6; ++a) maxDiag = std::max(fabs(Hll[36 * i + a * 6 + a]), maxDiag);
      lambda = tau * maxDiag;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      // solve (H + lambda I) x = b by eliminating the landmark blocks (points, then lines)
      double S[36], rhs[6], dp[6];
      for (int k = 0; k < 36; ++k) S[k] = Hpp[k];
      for (int a = 0; a < 6; ++a) { S[a * 6 + a] += lambda; rhs[a] = bp[a]; }
      bool ok = true;
      for (int i = 0; i < np; ++i) {
        double M[9];
        for (int k = 0; k < 9; ++k) M[k] = Hxx[9 * i + k];
        for (int a = 0; a < 3; ++a) M[a * 3 + a] += lambda;
        if (!inv_lu<3>(M, &HxxInv[9 * i])) ok = false;
        const double* hi = &HxxInv[9 * i]; const double* hpx = &Hpx[18 * i];
        double T[18];  // Hpx * Hxx^-1 (6x3)
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += hpx[a * 3 + k] * hi[k * 3 + c]; T[a * 3 + c] = s; }
        for (int a = 0; a < 6; ++a) {
          double s = 0; for (int k = 0; k < 3; ++k) s += T[a * 3 + k] * bx[3 * i + k];
          rhs[a] -= s;
          for (int c = 0; c < 6; ++c) { double h = 0; for (int k = 0; k < 3; ++k) h += T[a * 3 + k] * hpx[c * 3 + k]; S[a * 6 + c] -= h; }
        }
      }
      for (int i = 0; i < n; ++i) {
        double M[36];
        for (int k = 0; k < 36; ++k) M[k] = Hll[36 * i + k];
        for (int a = 0; a < 6; ++a) M[a * 6 + a] += lambda;
        if (!inv_lu<6>(M, &HllInv[36 * i])) ok = false;
        const double* hi = &HllInv[36 * i]; const double* hpl = &Hpl[36 * i];
        double T[36];  // Hpl * Hll^-1
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += hpl[a * 6 + k] * hi[k * 6 + c]; T[a * 6 + c] = s; }
        for (int a = 0; a < 6; ++a) {
          double s = 0; for (int k = 0; k < 6; ++k) s += T[a * 6 + k] * bl[6 * i + k];
          rhs[a] -= s;
          for (int c = 0; c < 6; ++c) { double h = 0; for (int k = 0; k < 6; ++k) h += T[a * 6 + k] * hpl[c * 6 + k]; S[a * 6 + c] -= h; }
        }
      }
      double Si[36];
      if (!inv_lu<6>(S, Si)) ok = false;
      for (int a = 0; a < 6; ++a) { double s = 0; for (int k = 0; k < 6; ++k) s += Si[a * 6 + k] * rhs[k]; dp[a] = s; }
      double scale = 0;
      for (int a = 0; a < 6; ++a) scale += dp[a] * (lambda * dp[a] + bp[a]);
      for (int i = 0; i < np; ++i) {
        const double* hi = &HxxInv[9 * i]; const double* hpx = &Hpx[18 * i];
        double r[3];
        for (int a = 0; a < 3; ++a) { double s = 0; for (int k = 0; k < 6; ++k) s += hpx[k * 3 + a] * dp[k]; r[a] = bx[3 * i + a] - s; }
        for (int a = 0; a < 3; ++a) { double s = 0; for (int k = 0; k < 3; ++k) s += hi[a * 3 + k] * r[k]; dx[3 * i + a] = s; }
        for (int a = 0; a < 3; ++a) scale += dx[3 * i + a] * (lambda * dx[3 * i + a] + bx[3 * i + a]);
      }
      for (int i = 0; i < n; ++i) {
        const double* hi = &HllInv[36 * i]; const double* hpl = &Hpl[36 * i];
        double r[6];
        for (int a = 0; a < 6; ++a) { double s = 0; for (int k = 0; k < 6; ++k) s += hpl[k * 6 + a] * dp[k]; r[a] = bl[6 * i + a] - s; }
        for (int a = 0; a < 6; ++a) { double s = 0; for (int k = 0; k < 6; ++k) s += hi[a * 6 + k] * r[k]; dl[6 * i + a] = s; }
        for (int a = 0; a < 6; ++a) scale += dl[6 * i + a] * (lambda * dl[6 * i + a] + bl[6 * i + a]);
      }
      Iso camNew; iso_oplus(cam1, dp, camNew);
      std::vector<double> Lnew(L), Xnew(X);
      for (int k = 0; k < 3 * np; ++k) Xnew[k] += dx[k];
      for (int k = 0; k < 6 * n; ++k) Lnew[k] += dl[k];
      double tempChi = chi2_all(camNew, Xnew, Lnew);
      if (!ok) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
        alpha = std::min(alpha, upS);
        double scaleFactor = std::max(lowS, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        cam1 = camNew; L.swap(Lnew); X.swap(Xnew);
      } else {
        lambda *= ni;
        ni *= 2;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) break;
  }
  Iso out; iso_inv(cam1, out);  // estimate().inverse(), cast to float
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tf[r * 4 + c] = (float)out.R[r * 3 + c]; tf[r * 4 + 3] = (float)out.t[r]; }
  tf[12] = 0; tf[13] = 0; tf[14] = 0; tf[15] = 1;
}

void refine_pose_lines(const std::vector<Line>& train, const std::vector<Line>& query, const std::vector<Match>& ms,
                       float tf[16], int iterations, const Params& P) {
  Points none;
  refine_pose_hybrid(train, query, none, none, std::vector<Match>(), ms, tf, iterations, 525.0, 0.0, P);
}

// ------------------------------------------------------- point matching ----
void rootsift(float* desc, int n, int dim) {  // src/node.cpp:1823-1837
  for (int r = 0; r < n; ++r) {
    float* d = desc + (size_t)r * dim;
    for (int c = 0; c < dim; ++c) d[c] = fabsf(d[c]);
    // cv::reduce(CV_REDUCE_SUM, CV_32FC1) over a float row: sequential float accumulation
    float sum = 0.f;
    for (int c = 0; c < dim; ++c) sum += d[c];
    if (sum == 0) continue;
    for (int c = 0; c < dim; ++c) d[c] = sqrtf(d[c] / sum);
  }
}

void featureMatching(const Points& q, const Points& t, double nn_ratio, GlibcRand& rng, std::vector<Match>& out) {
  if (q.n == 0 || t.n < 2) return;  // knnMatch needs two neighbours (node.cpp:628-629 reads [i][1])
  std::vector<char> used(t.n, 0);
  for (int i = 0; i < q.n; ++i) {
    // batchDistance K = 2: keeps the two smallest, earlier train index first on ties (strict <)
    float d1 = FLT_MAX, d2 = FLT_MAX; int i1 = -1, i2 = -1;
    for (int j = 0; j < t.n; ++j) {
      float d = sqrtf(l2sqr_f(q.desc + (size_t)i * q.dim, t.desc + (size_t)j * t.dim, q.dim));
      if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
      else if (d < d2) { d2 = d; i2 = j; }
    }
    (void)i2;
    float dist_ratio_fac = d1 / d2;
    if (dist_ratio_fac < nn_ratio) {
      if (used[i1]) continue;
      used[i1] = 1;
      Match m; m.queryIdx = i; m.trainIdx = i1;
      m.distance = (float)(dist_ratio_fac + (float)rng.next() / (1000.0 * 2147483647));
      out.push_back(m);
    }
  }
}

// Node::featureMatching with feature_extractor_type ORB (src/node.cpp:606-641): "BruteForce-HammingLUT" — the distance of
// two binary rows is the number of differing bits (a lookup table per byte), knnMatch hands it back as a float.
void featureMatching_hamming(const uint8_t* qd, int nq, const uint8_t* td, int nt, int nbytes, double nn_ratio, GlibcRand& rng,
                             std::vector<Match>& out) {
  if (nq == 0 || nt < 2) return;
  static uint8_t lut[256];
  static bool init = false;
  if (!init) { for (int v = 0; v < 256; ++v) { int c = 0; for (int b = 0; b < 8; ++b) c += (v >> b) & 1; lut[v] = (uint8_t)c; } init = true; }
  std::vector<char> used(nt, 0);
  for (int i = 0; i < nq; ++i) {
    float d1 = FLT_MAX, d2 = FLT_MAX; int i1 = -1;
    for (int j = 0; j < nt; ++j) {
      int c = 0;
      for (int k = 0; k < nbytes; ++k) c += lut[qd[(size_t)i * nbytes + k] ^ td[(size_t)j * nbytes + k]];
      float d = (float)c;
      if (d < d1) { d2 = d1; d1 = d; i1 = j; }
      else if (d < d2) d2 = d;
    }
    float dist_ratio_fac = d1 / d2;
    if (dist_ratio_fac < nn_ratio) {
      if (used[i1]) continue;
      used[i1] = 1;
      Match m; m.queryIdx = i; m.trainIdx = i1;
      m.distance = (float)(dist_ratio_fac + (float)rng.next() / (1000.0 * 2147483647));
      out.push_back(m);
    }
  }
}

// ------------------------------------------ pose RANSAC (points + lines) ----
static void tf_to_double(const float tf[16], double tfd[16]) { for (int i = 0; i < 16; ++i) tfd[i] = (double)tf[i]; }

struct ScoreH { std::vector<int> pt, ln; double sse; };
static void score_hybrid(const std::vector<Line>& train, const std::vector<Line>& query, const Points& tp, const Points& qp,
                         const std::vector<Match>& pm, const std::vector<Match>& lm, const float tf[16], double thr,
                         double sigma_depth, ScoreH& sc, bool float_sse) {
  sc.pt.clear(); sc.ln.clear();
  float sse_f = 0; double sse_d = 0;
  double tfd[16]; tf_to_double(tf, tfd);
  for (size_t i = 0; i < pm.size(); ++i) {
    double d2 = error_function2(qp.xyz1 + 4 * pm[i].queryIdx, tp.xyz1 + 4 * pm[i].trainIdx, tfd, sigma_depth);
    if (d2 < thr * thr) { sc.pt.push_back((int)i); if (float_sse) sse_f += d2; else sse_d += d2; }
  }
  for (size_t i = 0; i < lm.size(); ++i) {
    const Line& q = query[lm[i].queryIdx];
    const Line& t = train[lm[i].trainIdx];
    float qa[4] = {(float)q.A[0], (float)q.A[1], (float)q.A[2], 1.f}, qb[4] = {(float)q.B[0], (float)q.B[1], (float)q.B[2], 1.f};
    double qA[3], qB[3];
    tf_apply_f(tf, qa, qA);
    tf_apply_f(tf, qb, qB);
    double da = mah_dist3d_pt_line(t.A, t.DU_A, qA, qB);
    double db = mah_dist3d_pt_line(t.B, t.DU_B, qA, qB);
    if (da < thr && db < thr) {
      sc.ln.push_back((int)i);
      if (float_sse) sse_f += da * da + db * db; else sse_d += da * da + db * db;
    }
  }
  sc.sse = float_sse ? (double)sse_f : sse_d;
}

// getTransform_Lns_Pts_pcl (motion.cpp:530-579)
static bool transform_lns_pts_pcl(const std::vector<Line>& train, const std::vector<Line>& query, const Points& tp, const Points& qp,
                                  const Match* pm, int npm, const Match* lm, int nlm, GlibcRand& rng, float tf[16]) {
  if (npm < 1 || npm + nlm < 3) return false;
  Tfc tfc; tfc_reset(&tfc);
  for (int i = 0; i < nlm; ++i) {
    int ptidx = rng.next() % npm;
    const float* tpt = tp.xyz1 + 4 * pm[ptidx].trainIdx;
    const float* qpt = qp.xyz1 + 4 * pm[ptidx].queryIdx;
    double train_pt[3] = {(double)tpt[0], (double)tpt[1], (double)tpt[2]}, query_pt[3] = {(double)qpt[0], (double)qpt[1], (double)qpt[2]};
    double train_prj[3], query_prj[3];
    project_pt_ln(train_pt, train[lm[i].trainIdx].A, train[lm[i].trainIdx].B, train_prj);
    project_pt_ln(query_pt, query[lm[i].queryIdx].A, query[lm[i].queryIdx].B, query_prj);
    float from[3] = {(float)query_prj[0], (float)query_prj[1], (float)query_prj[2]}, to[3] = {(float)train_prj[0], (float)train_prj[1], (float)train_prj[2]};
    if (from[2] != from[2] || to[2] != to[2]) continue;
    float weight = 1 / (fabsf(to[2]) + fabsf(from[2]));
    tfc_add(&tfc, from, to, weight);
  }
  for (int i = 0; i < npm; ++i) {
    const float* from = qp.xyz1 + 4 * pm[i].queryIdx;
    const float* to = tp.xyz1 + 4 * pm[i].trainIdx;
    if (from[2] != from[2] || to[2] != to[2]) continue;
    float weight = 1 / (fabsf(to[2]) + fabsf(from[2]));
    tfc_add(&tfc, from, to, weight);
  }
  bool valid = tfc.n >= 3;
  tfc_get(&tfc, tf);
  return valid;
}

void getTransform_PtsLines_ransac(const std::vector<Line>& train, const std::vector<Line>& query, const Points& train_pts,
                                  const Points& query_pts, int id_train, int id_query, const std::vector<Match>& all_pt,
                                  const std::vector<Match>& all_ln, GlibcRand& rng, double fx, double asynch_dt,
                                  const Params& P, PoseResult& out) {
  out = PoseResult();
  for (int i = 0; i < 16; ++i) out.tf[i] = out.tf_ransac[i] = (i % 5 == 0) ? 1.f : 0.f;
  int nPt = (int)all_pt.size(), nLn = (int)all_ln.size();
  int min_inlier_nmb = P.min_feature_matches, line_weight = P.line_match_number_weight;
  if (nPt + nLn * line_weight < min_inlier_nmb) { out.rmse = 1e9f; return; }
  if (min_inlier_nmb > 0.7 * (nPt + nLn * line_weight)) min_inlier_nmb = 0.7 * (nPt + nLn * line_weight);
  if (abs(id_train - id_query) > 50) min_inlier_nmb = P.min_matches_loopclose;
  std::vector<int> indexes(nPt + nLn);
  for (size_t i = 0; i < indexes.size(); ++i) indexes[i] = (int)i;
  int maxIter = P.ransac_iters_line_motion;
  double thr = P.max_mah_dist_for_inliers;
  float sum_squared_error = 1e9f;
  std::vector<int> best_pt, best_ln;
  float tf_best[16];
  for (int i = 0; i < 16; ++i) tf_best[i] = 0;
  int iter = 0;
  while (iter < maxIter) {
    ++iter;
    int left = (int)indexes.size();
    for (int k = 0; k < 3; ++k) { int r = rng.next() % left; std::swap(indexes[k], indexes[k + r]); --left; }
    Match ptM[3], lnM[3]; int npm = 0, nlm = 0;
    for (int k = 0; k < 3; ++k) {
      if (indexes[k] < nPt) ptM[npm++] = all_pt[indexes[k]];
      else lnM[nlm++] = all_ln[indexes[k] - nPt];
    }
    float tf[16];
    bool valid;
    if (nlm == 3) {
      double qA[9], qB[9], tA[9], tB[9];
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) {
          qA[3 * k + c] = query[lnM[k].queryIdx].A[c]; qB[3 * k + c] = query[lnM[k].queryIdx].B[c];
          tA[3 * k + c] = train[lnM[k].trainIdx].A[c]; tB[3 * k + c] = train[lnM[k].trainIdx].B[c];
        }
      double R[9], t[3];
      valid = relmotion_svd(qA, qB, tA, tB, 3, R, t);
      for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tf[r * 4 + c] = (float)R[r * 3 + c]; tf[r * 4 + 3] = (float)t[r]; }
      tf[12] = tf[13] = tf[14] = 0; tf[15] = 1;
    } else {
      valid = transform_lns_pts_pcl(train, query, train_pts, query_pts, ptM, npm, lnM, nlm, rng, tf);
    }
    if (!valid) continue;
    ScoreH sc;
    score_hybrid(train, query, train_pts, query_pts, all_pt, all_ln, tf, thr, P.sigma_depth, sc, true);
    if ((int)(sc.pt.size() + line_weight * sc.ln.size()) > (int)(best_pt.size() + line_weight * best_ln.size())) {
      best_pt = sc.pt; best_ln = sc.ln;
      memcpy(tf_best, tf, sizeof(tf));
      sum_squared_error = (float)sc.sse;
      out.best_iter = iter - 1;
    }
  }
  if (best_pt.size() + best_ln.size() < 3) return;
  for (int i : best_pt) out.pt_ransac_inliers.push_back(all_pt[i]);
  for (int i : best_ln) out.ransac_inliers.push_back(all_ln[i]);
  memcpy(out.tf_ransac, tf_best, sizeof(tf_best));
  float refined_tf[16];
  memcpy(refined_tf, tf_best, sizeof(tf_best));
  refine_pose_hybrid(train, query, train_pts, query_pts, out.pt_ransac_inliers, out.ransac_inliers, refined_tf, 25, fx, asynch_dt, P);
  // float / size_t -> float division, std::sqrt(float) (lineslam.h:35 `using namespace std`), motion.cpp:731
  double refined_rmse = (double)sqrtf(sum_squared_error / (float)(best_pt.size() + best_ln.size()));
  std::vector<Match> refined_pt, refined_ln;
  for (int it = 0; it < 20; ++it) {
    ScoreH sc;
    score_hybrid(train, query, train_pts, query_pts, all_pt, all_ln, refined_tf, thr, P.sigma_depth, sc, false);
    if (sc.pt.size() + sc.ln.size() * P.line_match_number_weight > refined_pt.size() + refined_ln.size() * P.line_match_number_weight) {
      refined_pt.clear(); refined_ln.clear();
      for (int i : sc.pt) refined_pt.push_back(all_pt[i]);
      for (int i : sc.ln) refined_ln.push_back(all_ln[i]);
      refined_rmse = sqrt(sc.sse / (double)(sc.pt.size() + sc.ln.size()));
      refine_pose_hybrid(train, query, train_pts, query_pts, refined_pt, refined_ln, refined_tf, 20, fx, asynch_dt, P);
    } else break;
  }
  out.pt_inliers = refined_pt;
  out.inliers = refined_ln;
  out.rmse = (float)refined_rmse;
  memcpy(out.tf, refined_tf, sizeof(refined_tf));
  out.found = (int)(refined_pt.size() + line_weight * refined_ln.size()) >= min_inlier_nmb;
}

// getTransform_PtsLines_ransac with line-only input (nPt = 0)
void getTransform_Lines_ransac(const std::vector<Line>& train, const std::vector<Line>& query, int id_train, int id_query,
                               const std::vector<Match>& all_ln, uint32_t seed, const Params& P, PoseResult& out) {
  Points none;
  GlibcRand rng; rng.seed(seed);
  getTransform_PtsLines_ransac(train, query, none, none, id_train, id_query, std::vector<Match>(), all_ln, rng, 525.0, 0.0, P, out);
}

// ------------------------------------ line-only RANSAC + levmar refinement ----
// dist3d_pt_line (utils.cpp:626-636)
static double dist3d_pt_line(const double X[3], const double A[3], const double B[3]) {
  double AB[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
  double nab = sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
  if (nab < 1e-10) return -1;
  double XA[3] = {X[0] - A[0], X[1] - A[1], X[2] - A[2]};
  double ax = sqrt(XA[0] * XA[0] + XA[1] * XA[1] + XA[2] * XA[2]);
  double inv = 1 / nab;
  double nv[3] = {(B[0] - A[0]) * inv, (B[1] - A[1]) * inv, (B[2] - A[2]) * inv};
  double d = XA[0] * nv[0] + XA[1] * nv[1] + XA[2] * nv[2];
  return sqrt(fabs(ax * ax - d * d));
}
// r2q (utils.cpp:1696-1707)
static void r2q(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  double r = sqrt(1 + t);
  double s = 0.5 / r;
  q[0] = 0.5 * r; q[1] = (R[7] - R[5]) * s; q[2] = (R[2] - R[6]) * s; q[3] = (R[3] - R[1]) * s;
}
struct RelData { const std::vector<Line>* a; const std::vector<Line>* b; };
// costFun_optimizeRelmotion (motion.cpp:60-96), OPT_USE_MAHDIST: cv::Mat products R*x + t and R^T*(x - t)
static void relmotion_cost(double* p, double* error, int, int, void* adata) {
  const RelData* d = (const RelData*)adata;
  double R[9];
  q2r(p, R);
  const double* t = p + 4;
  for (size_t i = 0; i < d->a->size(); ++i) {
    const Line& a = (*d->a)[i];
    const Line& b = (*d->b)[i];
    double aA[3], aB[3], bA[3], bB[3];
    for (int r = 0; r < 3; ++r) {
      aA[r] = (R[r * 3] * a.A[0] + R[r * 3 + 1] * a.A[1] + R[r * 3 + 2] * a.A[2]) + t[r];
      aB[r] = (R[r * 3] * a.B[0] + R[r * 3 + 1] * a.B[1] + R[r * 3 + 2] * a.B[2]) + t[r];
    }
    double dA[3] = {b.A[0] - t[0], b.A[1] - t[1], b.A[2] - t[2]}, dB[3] = {b.B[0] - t[0], b.B[1] - t[1], b.B[2] - t[2]};
    for (int r = 0; r < 3; ++r) {
      bA[r] = R[r] * dA[0] + R[3 + r] * dA[1] + R[6 + r] * dA[2];
      bB[r] = R[r] * dB[0] + R[3 + r] * dB[1] + R[6 + r] * dB[2];
    }
    error[i] = 0.25 * (mah_dist3d_pt_line(b.A, b.DU_A, aA, aB) + mah_dist3d_pt_line(b.B, b.DU_B, aA, aB) +
                       mah_dist3d_pt_line(a.A, a.DU_A, bA, bB) + mah_dist3d_pt_line(a.B, a.DU_B, bA, bB));
  }
}
void optimizeRelmotion(const std::vector<Line>& a, const std::vector<Line>& b, double R[9], double t[3]) {
  double q[4];
  r2q(R, q);
  double opts[5] = {1E-03, 1E-10, 1E-20, 1E-20, 1E-06}, info[10];
  double para[7] = {q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  std::vector<double> meas(a.size(), 0.0);
  RelData data{&a, &b};
  dlevmar_dif_restated(relmotion_cost, para, meas.data(), 7, (int)a.size(), 50, opts, info, &data);
  q2r(para, R);
  t[0] = para[4]; t[1] = para[5]; t[2] = para[6];
}
static void relmotion_consensus(const std::vector<Line>& a, const std::vector<Line>& b, const double R[9], const double t[3],
                                double distThresh, double angThresh, std::vector<int>& inl) {
  const double PI_T = 3.14159265;
  inl.clear();
  for (size_t i = 0; i < a.size(); ++i) {
    double aA[3], aB[3], RaAB[3];
    double aAB[3] = {a[i].A[0] - a[i].B[0], a[i].A[1] - a[i].B[1], a[i].A[2] - a[i].B[2]};
    double bAB[3] = {b[i].A[0] - b[i].B[0], b[i].A[1] - b[i].B[1], b[i].A[2] - b[i].B[2]};
    for (int r = 0; r < 3; ++r) {  // Eigen Matrix3d * Vector3d + Vector3d
      aA[r] = ((R[r * 3] * a[i].A[0] + R[r * 3 + 1] * a[i].A[1]) + R[r * 3 + 2] * a[i].A[2]) + t[r];
      aB[r] = ((R[r * 3] * a[i].B[0] + R[r * 3 + 1] * a[i].B[1]) + R[r * 3 + 2] * a[i].B[2]) + t[r];
      RaAB[r] = (R[r * 3] * aAB[0] + R[r * 3 + 1] * aAB[1]) + R[r * 3 + 2] * aAB[2];
    }
    double dist = 0.5 * dist3d_pt_line(aA, b[i].A, b[i].B) + 0.5 * dist3d_pt_line(aB, b[i].A, b[i].B);
    double dt = (RaAB[0] * bAB[0] + RaAB[1] * bAB[1]) + RaAB[2] * bAB[2];
    double na = sqrt(aAB[0] * aAB[0] + aAB[1] * aAB[1] + aAB[2] * aAB[2]);
    double nb = sqrt(bAB[0] * bAB[0] + bAB[1] * bAB[1] + bAB[2] * bAB[2]);
    double angle = 180 * lsl_acos(fabs(dt / na / nb)) / PI_T;
    if (dist < distThresh && angle < angThresh) inl.push_back((int)i);
  }
}
void computeRelativeMotion_Ransac(const std::vector<Line>& a, const std::vector<Line>& b, uint32_t seed, const Params& P,
                                  RelMotion& out) {
  out = RelMotion();
  const int n = (int)a.size();
  if (n < 3 || (int)b.size() < 3) return;
  const double PI_T = 3.14159265;
  std::vector<double> au(3 * n);
  for (int i = 0; i < n; ++i) {
    double l[3] = {a[i].B[0] - a[i].A[0], a[i].B[1] - a[i].A[1], a[i].B[2] - a[i].A[2]};
    double inv = 1 / sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; ++k) au[3 * i + k] = l[k] * inv;
  }
  const int maxIters = P.ransac_iters_line_motion;
  const double distThresh = P.pt2line3d_dist_relmotion, angThresh = P.line3d_angle_relmotion;
  const double cosDeg = lsl_cos(5 * PI_T / 180);
  std::vector<int> indexes(n);
  for (int i = 0; i < n; ++i) indexes[i] = i;
  GlibcRand rng; rng.seed(seed);
  std::vector<int> maxConSet;
  double bR[9], bt[3];
  int iter = 0;
  while (iter < maxIters) {
    iter++;
    int left = n;
    for (int k = 0; k < 3; ++k) { int r = rng.next() % left; std::swap(indexes[k], indexes[k + r]); --left; }
    bool degenerate = true;
    for (int i = 0; i < 3 && degenerate; ++i)
      for (int j = i + 1; j < 3; ++j) {
        const double* ui = &au[3 * indexes[i]]; const double* uj = &au[3 * indexes[j]];
        if (fabs(ui[0] * uj[0] + ui[1] * uj[1] + ui[2] * uj[2]) < cosDeg) { degenerate = false; break; }
      }
    if (degenerate) continue;
    double qA[9], qB[9], tA[9], tB[9];
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 3; ++c) {
        qA[3 * k + c] = a[indexes[k]].A[c]; qB[3 * k + c] = a[indexes[k]].B[c];
        tA[3 * k + c] = b[indexes[k]].A[c]; tB[3 * k + c] = b[indexes[k]].B[c];
      }
    double R[9], t[3];
    relmotion_svd(qA, qB, tA, tB, 3, R, t);
    std::vector<int> inlier;
    relmotion_consensus(a, b, R, t, distThresh, angThresh, inlier);
    if (inlier.size() > maxConSet.size()) { maxConSet = inlier; memcpy(bR, R, sizeof(bR)); memcpy(bt, t, sizeof(bt)); }
    if (maxConSet.size() >= (size_t)n * 1) break;
  }
  out.conset = maxConSet;
  if (maxConSet.size() < 1) return;
  memcpy(out.R, bR, sizeof(bR)); memcpy(out.t, bt, sizeof(bt)); out.have = true;
  if (maxConSet.size() < 4) return;
  std::vector<Line> ina, inb;
  for (int i : maxConSet) { ina.push_back(a[i]); inb.push_back(b[i]); }
  optimizeRelmotion(ina, inb, out.R, out.t); out.lm_calls++;
  double R[9], t[3];
  memcpy(R, out.R, sizeof(R)); memcpy(t, out.t, sizeof(t));
  std::vector<int> prevConSet;
  while (1) {
    std::vector<int> conset;
    relmotion_consensus(a, b, R, t, distThresh, angThresh, conset);
    if (conset.size() <= prevConSet.size()) break;
    prevConSet = conset;
    memcpy(out.R, R, sizeof(R)); memcpy(out.t, t, sizeof(t));
    ina.clear(); inb.clear();
    for (int i : prevConSet) { ina.push_back(a[i]); inb.push_back(b[i]); }
    optimizeRelmotion(ina, inb, R, t); out.lm_calls++;
  }
  out.conset = prevConSet;
}

}  // namespace orc
