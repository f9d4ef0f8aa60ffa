// lsl_points.h — point-feature arithmetic of the pair stage, shared by the CPU oracle and the sm_100a
// kernels so both execute the same IEEE operation sequence (float where the reference is float):
//   cv::BFMatcher L2 distance (OpenCV 2.4 normL2Sqr_)          -> l2sqr_f
//   errorFunction2 (src/misc.cpp:699-786) + depth_covariance (src/misc2.h:21-36) -> error_function2
//   Eigen LDLT 3x3 solve (misc.cpp:779), Eigen 3x3 inverse                    -> ldlt3_solve, inv3_eigen
//   pcl::TransformationFromCorrespondences (PCL 1.7)          -> Tfc, tfc_add, tfc_get
//   compPt3dCov(Eigen::Vector3f ...) (src/line/utils.cpp:724-742)             -> pt_info_f
//   projectPt3d2Ln3d_2 (src/line/utils.cpp:506-512)                           -> project_pt_ln
// OpenCV, Eigen and PCL are not in /root/reference: these are restatements of the published
// algorithms (parity with the real libraries UNPINNED, Tier-T; oracle and GPU agree bit for bit).
#pragma once
#include <float.h>
#include "lsl_math.h"

namespace lslm {

// normL2Sqr_(const float*, const float*, int) of OpenCV 2.4 modules/core/src/stat.cpp, SSE2 branch
// (the reference is an x86-64 build): two 4-lane accumulators over blocks of 8, lanes added
// (d0 + d1) then buf[0] + buf[1] + buf[2] + buf[3]; scalar tail.
LSL_HD float l2sqr_f(const float* a, const float* b, int n) {
  float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
  int j = 0;
  for (; j <= n - 8; j += 8) {
    for (int l = 0; l < 4; ++l) {
      float t0 = a[j + l] - b[j + l], t1 = a[j + 4 + l] - b[j + 4 + l];
      d0[l] = d0[l] + t0 * t0;
      d1[l] = d1[l] + t1 * t1;
    }
  }
  float s0 = d0[0] + d1[0], s1 = d0[1] + d1[1], s2 = d0[2] + d1[2], s3 = d0[3] + d1[3];
  float d = s0 + s1 + s2 + s3;
  for (; j < n; ++j) { float t = a[j] - b[j]; d += t * t; }
  return d;
}

// Eigen compute_inverse_size3 (cofactors of column 0 for the determinant). Row-major 3x3.
LSL_HD void inv3_eigen(const double* m, double* r) {
#define LSL_COF(i, j) (m[((i + 1) % 3) * 3 + (j + 1) % 3] * m[((i + 2) % 3) * 3 + (j + 2) % 3] - \
                       m[((i + 1) % 3) * 3 + (j + 2) % 3] * m[((i + 2) % 3) * 3 + (j + 1) % 3])
  double c0 = LSL_COF(0, 0), c1 = LSL_COF(1, 0), c2 = LSL_COF(2, 0);
  double det = c0 * m[0] + c1 * m[3] + c2 * m[6];
  double invdet = 1.0 / det;
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  r[3] = LSL_COF(0, 1) * invdet; r[4] = LSL_COF(1, 1) * invdet; r[5] = LSL_COF(2, 1) * invdet;
  r[6] = LSL_COF(0, 2) * invdet; r[7] = LSL_COF(1, 2) * invdet; r[8] = LSL_COF(2, 2) * invdet;
#undef LSL_COF
}

// x = A^-1 b through Eigen's LDLT (unblocked lower, symmetric pivoting on the largest diagonal,
// Eigen 3.2 ldlt_inplace<Lower>::unblocked + solve). A symmetric 3x3 row-major (copied).
LSL_HD void ldlt3_solve(const double* Ain, const double* b, double* x) {
  double A[9];
  for (int i = 0; i < 9; ++i) A[i] = Ain[i];
  int tr[3] = {0, 1, 2};
  double cutoff = 0.0;
  for (int k = 0; k < 3; ++k) {
    int big = k;
    double bv = fabs(A[k * 3 + k]);
    for (int i = k + 1; i < 3; ++i) if (fabs(A[i * 3 + i]) > bv) { bv = fabs(A[i * 3 + i]); big = i; }
    if (k == 0) cutoff = fabs(DBL_EPSILON * bv);
    if (bv < cutoff) { for (int i = k; i < 3; ++i) { A[i * 3 + i] = 0.0; tr[i] = i; } break; }
    tr[k] = big;
    if (big != k) {  // symmetric swap of rows/cols k and big in the lower triangle
      int s = 3 - big - 1;
      for (int j = 0; j < k; ++j) { double t = A[k * 3 + j]; A[k * 3 + j] = A[big * 3 + j]; A[big * 3 + j] = t; }
      for (int j = 0; j < s; ++j) { double t = A[(big + 1 + j) * 3 + k]; A[(big + 1 + j) * 3 + k] = A[(big + 1 + j) * 3 + big]; A[(big + 1 + j) * 3 + big] = t; }
      { double t = A[k * 3 + k]; A[k * 3 + k] = A[big * 3 + big]; A[big * 3 + big] = t; }
      for (int i = k + 1; i < big; ++i) { double t = A[i * 3 + k]; A[i * 3 + k] = A[big * 3 + i]; A[big * 3 + i] = t; }
    }
    int rs = 3 - k - 1;
    if (k > 0) {
      double temp[2];
      for (int j = 0; j < k; ++j) temp[j] = A[j * 3 + j] * A[k * 3 + j];
      double s = 0.0;
      for (int j = 0; j < k; ++j) s += A[k * 3 + j] * temp[j];
      A[k * 3 + k] -= s;
      for (int i = 0; i < rs; ++i) {
        double t = 0.0;
        for (int j = 0; j < k; ++j) t += A[(k + 1 + i) * 3 + j] * temp[j];
        A[(k + 1 + i) * 3 + k] -= t;
      }
    }
    if (rs > 0 && fabs(A[k * 3 + k]) > cutoff)
      for (int i = 0; i < rs; ++i) A[(k + 1 + i) * 3 + k] /= A[k * 3 + k];
  }
  double y[3] = {b[0], b[1], b[2]};
  for (int k = 0; k < 3; ++k) if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }   // P b
  for (int i = 1; i < 3; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * 3 + j] * y[j];              // L^-1
  double dmax = fmax(fmax(fabs(A[0]), fabs(A[4])), fabs(A[8]));
  double tol = fmax(dmax * DBL_EPSILON, 1.0 / DBL_MAX);
  for (int i = 0; i < 3; ++i) { if (fabs(A[i * 3 + i]) > tol) y[i] /= A[i * 3 + i]; else y[i] = 0.0; }  // D^-1
  for (int i = 1; i >= 0; --i) for (int j = i + 1; j < 3; ++j) y[i] -= A[j * 3 + i] * y[j];         // L^-T
  for (int k = 2; k >= 0; --k) if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }  // P^T
  x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
}

// errorFunction2 (src/misc.cpp:699-786): squared Mahalanobis distance of x1 (query point) mapped by tf
// against x2 (train point). tf = the float pose cast to double, row-major 4x4 (only rows 0..2 used).
// raster_cov_x/y = (3 tan(58deg/640))^2, (3 tan(45deg/480))^2 (glibc tan, misc.cpp:704-711).
LSL_HD double error_function2(const float* x1, const float* x2, const double* tf, double sigma_depth) {
  const double raster_cov_x = 0x1.79c2199b5183dp-16, raster_cov_y = 0x1.94427cddf0ce5p-16;
  if (x1[2] != x1[2] || x2[2] != x2[2]) return DBL_MAX;
  double a[4] = {(double)x1[0], (double)x1[1], (double)x1[2], (double)x1[3]};
  double mu2[3] = {(double)x2[0], (double)x2[1], (double)x2[2]};
  double m12[3];
  for (int r = 0; r < 3; ++r) m12[r] = ((tf[r * 4] * a[0] + tf[r * 4 + 1] * a[1]) + tf[r * 4 + 2] * a[2]) + tf[r * 4 + 3] * a[3];
  double dmu[3] = {m12[0] - mu2[0], m12[1] - mu2[1], m12[2] - mu2[2]};
  double sd1 = sigma_depth * a[2] * a[2], sd2 = sigma_depth * mu2[2] * mu2[2];
  double dc1 = sd1 * sd1, dc2 = sd2 * sd2;
  {
    double dsq = (dmu[0] * dmu[0] + dmu[1] * dmu[1]) + dmu[2] * dmu[2];
    double s1 = raster_cov_x < dc1 ? dc1 : raster_cov_x;   // std::max(a, b): a unless a < b
    double s2 = raster_cov_x < dc2 ? dc2 : raster_cov_x;
    if (dsq > 2.0 * (s1 + s2)) return DBL_MAX;
  }
  double c1[3] = {1 * raster_cov_x * a[2], 1 * raster_cov_y * a[2], dc1};
  double c2[3] = {1 * raster_cov_x * mu2[2], 1 * raster_cov_y * mu2[2], dc2};
  // cov1_in_frame_2 = (R^T * cov1) * R, dense 3x3 products with the zero terms (exact) omitted
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double m0 = tf[0 * 4 + i] * c1[0], m1 = tf[1 * 4 + i] * c1[1], m2 = tf[2 * 4 + i] * c1[2];
      S[i * 3 + j] = (m0 * tf[0 * 4 + j] + m1 * tf[1 * 4 + j]) + m2 * tf[2 * 4 + j];
    }
  if (dmu[2] != dmu[2]) dmu[2] = 0.0;
  S[0] += c2[0]; S[4] += c2[1]; S[8] += c2[2];
  double x[3];
  ldlt3_solve(S, dmu, x);
  double d2 = (dmu[0] * x[0] + dmu[1] * x[1]) + dmu[2] * x[2];
  if (!(d2 >= 0.0)) return DBL_MAX;
  return d2;
}

// projectPt3d2Ln3d_2 (src/line/utils.cpp:506-512)
LSL_HD void project_pt_ln(const double* P, const double* A, const double* B, double* out) {
  double AB[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, AP[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
  double s = (AB[0] * AP[0] + AB[1] * AP[1] + AB[2] * AP[2]) / (AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
  for (int k = 0; k < 3; ++k) out[k] = A[k] + s * AB[k];
}

// pcl::TransformationFromCorrespondences (all float): incremental weighted means and covariance.
struct Tfc { int n; float W, m1[3], m2[3], cov[9]; };
LSL_HD void tfc_reset(Tfc* t) {
  t->n = 0; t->W = 0.f;
  for (int i = 0; i < 3; ++i) t->m1[i] = t->m2[i] = 0.f;
  for (int i = 0; i < 9; ++i) t->cov[i] = 0.f;
}
LSL_HD void tfc_add(Tfc* t, const float* p /*from*/, const float* c /*to*/, float w) {
  if (w == 0.0f) return;
  ++t->n;
  t->W += w;
  float alpha = w / t->W;
  float d1[3] = {p[0] - t->m1[0], p[1] - t->m1[1], p[2] - t->m1[2]};
  float d2[3] = {c[0] - t->m2[0], c[1] - t->m2[1], c[2] - t->m2[2]};
  float om = 1.0f - alpha;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t->cov[i * 3 + j] = om * (t->cov[i * 3 + j] + alpha * (d2[i] * d1[j]));
  for (int i = 0; i < 3; ++i) { t->m1[i] += alpha * d1[i]; t->m2[i] += alpha * d2[i]; }
}
// 3x3 float SVD A = U diag(s) V^T by one-sided (Hestenes) Jacobi, singular values descending; the third
// left vector is completed as u0 x u1 (a minimal 3-sample covariance has rank <= 2). Stand-in for
// Eigen::JacobiSVD<Matrix3f>: U, V differ from Eigen's by signs/rounding; R = U diag(1,1,d) V^T does not
// depend on the sign choices.
LSL_HD void svd3_f(const float* Ain, float* U, float* s, float* V) {
  float A[9];
  for (int i = 0; i < 9; ++i) { A[i] = Ain[i]; V[i] = (i % 4 == 0) ? 1.f : 0.f; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        float al = 0.f, be = 0.f, ga = 0.f;
        for (int k = 0; k < 3; ++k) { al += A[k * 3 + p] * A[k * 3 + p]; be += A[k * 3 + q] * A[k * 3 + q]; ga += A[k * 3 + p] * A[k * 3 + q]; }
        if (ga == 0.f || fabsf(ga) <= 1e-7f * sqrtf(al * be)) continue;
        rotated = 1;
        float zeta = (be - al) / (2.0f * ga);
        float t = 1.0f / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
        if (zeta < 0.f) t = -t;
        float c = 1.0f / sqrtf(1.0f + t * t), sn = c * t;
        for (int k = 0; k < 3; ++k) {
          float ap = A[k * 3 + p], aq = A[k * 3 + q];
          A[k * 3 + p] = c * ap - sn * aq; A[k * 3 + q] = sn * ap + c * aq;
          float vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - sn * vq; V[k * 3 + q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  float nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrtf(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i) {  // selection sort, descending, stable
    int m = i;
    for (int j = i + 1; j < 3; ++j) if (nrm[ord[j]] > nrm[ord[m]]) m = j;
    if (m != i) { int t = ord[i]; ord[i] = ord[m]; ord[m] = t; }
  }
  float Vs[9];
  for (int j = 0; j < 3; ++j) {
    s[j] = nrm[ord[j]];
    for (int k = 0; k < 3; ++k) Vs[k * 3 + j] = V[k * 3 + ord[j]];
  }
  for (int j = 0; j < 2; ++j) {
    float inv = 1.0f / s[j];
    for (int k = 0; k < 3; ++k) U[k * 3 + j] = A[k * 3 + ord[j]] * inv;
  }
  U[2] = U[3] * U[7] - U[6] * U[4];     // u2 = u0 x u1
  U[5] = U[6] * U[1] - U[0] * U[7];
  U[8] = U[0] * U[4] - U[3] * U[1];
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}
LSL_HD float det3_f(const float* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
// getTransformation(): R = U s V^T with s = diag(1,1,sign), t = mean2 - R mean1; tf row-major 4x4 float.
LSL_HD void tfc_get(const Tfc* t, float* tf) {
  float U[9], s[3], V[9];
  svd3_f(t->cov, U, s, V);
  float sg = (det3_f(U) * det3_f(V) < 0.f) ? -1.f : 1.f;
  float R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[i * 3 + j] = (U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1]) + (U[i * 3 + 2] * sg) * V[j * 3 + 2];
  for (int i = 0; i < 3; ++i) {
    float rm = (R[i * 3] * t->m1[0] + R[i * 3 + 1] * t->m1[1]) + R[i * 3 + 2] * t->m1[2];
    for (int j = 0; j < 3; ++j) tf[i * 4 + j] = R[i * 3 + j];
    tf[i * 4 + 3] = t->m2[i] - rm;
  }
  tf[12] = 0.f; tf[13] = 0.f; tf[14] = 0.f; tf[15] = 1.f;
}

// information matrix of a point edge (transformation_estimation.cpp:262,278):
// compPt3dCov(Eigen::Vector3f, f, cu, cv, dt) (utils.cpp:724-742; double arithmetic on the float
// coordinates, result rounded to Matrix3f) .cast<double>().inverse()
LSL_HD void pt_info_f(const float* pt, double f, double sigma_impt, double c1, double c2, double c3, double dt, double* info) {
  double c2e = c2 + (dt - 0.005 > 0.0 ? dt - 0.005 : 0.0) * 0.5;
  double x = (double)pt[0], y = (double)pt[1], z = (double)pt[2];
  double sz = c1 * z * z + c2e * z + c3;
  double J[9] = {z / f, 0, x / z, 0, z / f, y / z, 0, 0, 1};
  double S[3] = {sigma_impt * sigma_impt, sigma_impt * sigma_impt, sz * sz};
  double cf[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += (J[i * 3 + k] * S[k]) * J[j * 3 + k];
      cf[i * 3 + j] = (double)(float)s;
    }
  inv3_eigen(cf, info);
}

}  // namespace lslm
